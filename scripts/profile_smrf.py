"""Small driver for ncu / stage timing: one smrf() on a synthetic cloud resident in HBM."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import neilpy_b200 as nb
from bench import make_cloud, PARAMS

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pts = torch.from_numpy(make_cloud(n, 0)).cuda()
torch.cuda.synchronize()
for _ in range(reps):
    t0 = time.perf_counter()
    st = {}
    Z, t, oc, op = nb.smrf(pts, return_stages=st, **PARAMS)
    torch.cuda.synchronize()
    print('smrf %.1f ms  grid %s  inpaint iters %s %s  objects %.3f' % (
        (time.perf_counter() - t0) * 1e3, tuple(Z.shape), st['inpaint1'], st['inpaint2'], float(op.float().mean())))
