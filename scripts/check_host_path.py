"""numpy in -> numpy out equals CUDA in -> CUDA out (the host staging path of api._to_host), and its time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neilpy_b200 as nb
from neilpy_b200.synth import synth_cloud
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
side = (n / 2.0) ** 0.5
x, y, z, _ = synth_cloud(n, side, side, seed=5)
xyzw = np.stack([x, y, z, np.zeros_like(x)], 1).astype(np.float32)
kw = dict(cellsize=1, windows=18, slope_threshold=.15, elevation_threshold=.5, elevation_scaler=1.25)
Zd, t, ocd, opd = nb.smrf(torch.as_tensor(xyzw).cuda(), **kw)
for rep in range(3):
    t0 = time.perf_counter()
    Zh, th, och, oph = nb.smrf(xyzw, **kw)
    dt = time.perf_counter() - t0
    assert isinstance(Zh, np.ndarray) and och.dtype == np.bool_ and oph.dtype == np.bool_ and Zh.dtype == np.float32
    dz = float(np.abs(Zh - Zd.cpu().numpy()).max())
    flips = int((och != ocd.cpu().numpy()).sum()), int((oph != opd.cpu().numpy()).sum())
    print('host path %.1f ms, max |dZ| %.3g, flips %s' % (dt * 1e3, dz, flips))
    assert dz < 1e-3 and max(flips) <= 3
print('ok')
