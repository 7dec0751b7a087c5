"""Stage timings of the row-band sharded smrf (run under torchrun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['SMRF_TIMING'] = '1'
import torch, torch.distributed as dist
from bench import make_cloud, PARAMS
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from neilpy_b200.distributed import smrf_sharded
pts = torch.from_numpy(make_cloud(int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000, rank, world)).cuda()
for i in range(3):
    r = smrf_sharded(pts, **PARAMS)
    if rank == 0 and i == 2:
        print(r['info'])
dist.destroy_process_group()
