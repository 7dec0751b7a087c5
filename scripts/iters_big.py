import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import neilpy_b200 as nb
from bench import make_cloud, PARAMS
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
pts = torch.cat([torch.from_numpy(make_cloud(n, r, 2)) for r in range(2)]).cuda()
st = {}
nb.smrf(pts, return_stages=st, **PARAMS)
print('single GPU, whole 2-rank cloud:', tuple(st['Zpro'].shape), st['inpaint1'], st['inpaint2'])
