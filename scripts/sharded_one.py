import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29533')
import torch, torch.distributed as dist
from bench import make_cloud, PARAMS
torch.cuda.set_device(0)
dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', 0))
import neilpy_b200 as nb
from neilpy_b200.distributed import smrf_sharded
pts = torch.from_numpy(make_cloud(20_000_000, 0, 1)).cuda()
r = smrf_sharded(pts, **PARAMS)
print('sharded world=1:', r['info']['inpaint1'], r['info']['inpaint2'])
st = {}
nb.smrf(pts, return_stages=st, **PARAMS)
print('single         :', st['inpaint1'], st['inpaint2'])
dist.destroy_process_group()
