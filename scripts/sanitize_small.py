"""Small end-to-end + per-kernel run for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neilpy_b200 as nb
from neilpy_b200.synth import synth_cloud

x, y, z, _ = synth_cloud(30000, 120.0, 90.0, seed=1)
xyzw = torch.as_tensor(np.stack([x, y, z, np.zeros_like(x)], 1).astype(np.float32)).cuda()
Z, t, oc, op = nb.smrf(xyzw, cellsize=1, windows=18, return_extras=False)
print('smrf ok', tuple(Z.shape), int(op.sum()))
Z64, t, oc, op = nb.smrf(x, y, z, cellsize=1, windows=5)
print('smrf f64 ok', Z64.shape)
rng = np.random.default_rng(0)
for shape in [(37, 41), (64, 513), (130, 953), (5, 300)]:
    A = rng.normal(size=shape).astype(np.float32)
    for w in (1, 3, 9, 18, 25):
        m = nb.progressive_filter(A, np.array([w]), 1, .15)
    m = nb.progressive_filter(A, np.arange(1, 7), 1, .15, return_when_dropped=True)
print('openings ok')
B = rng.normal(size=(90, 70)); B[rng.random(B.shape) < .4] = np.nan
print('inpaint ok', float(np.abs(nb.inpaint_nans_by_springs(B)).max()))
