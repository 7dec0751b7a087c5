"""Per-stage CUDA-event timings of smrf() on the bench workload (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neilpy_b200 as nb
from neilpy_b200 import _lib, api
from bench import make_cloud, PARAMS

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
pts = torch.from_numpy(make_cloud(n, 0)).cuda()
st = {}
nb.smrf(pts, return_stages=st, **PARAMS)
lib = _lib.load()

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

print('create_dem   %.2f ms' % timed(lambda: nb.create_dem(pts, None, None, 1, 'min')))
g1 = st['Zmin_binned']; g2 = st['Zpro_punched']
print('inpaint #1   %.2f ms' % timed(lambda: nb.inpaint_nans_by_springs(g1)))
print('inpaint #2   %.2f ms' % timed(lambda: nb.inpaint_nans_by_springs(g2)))
print('progressive  %.2f ms' % timed(lambda: nb.progressive_filter(st['Zmin_filtered'], np.arange(18) + 1, 1, .15)))
print('smrf         %.2f ms' % timed(lambda: nb.smrf(pts, **PARAMS)))
os.environ['SMRF_INPAINT_PRECOND'] = 'jacobi'
print('inpaint #2 jacobi %.2f ms' % timed(lambda: nb.inpaint_nans_by_springs(g2), 1))
