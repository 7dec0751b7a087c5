"""Deterministic synthetic SMRF workloads: terrain + flat-roof
buildings + vegetation patches + low outliers.  numpy only; shared by the tests, the
oracle-side baselines and bench.py (there is no network for real tiles, and the
reference's DK22 sample is absent from its checkout)."""
import numpy as np


def terrain(x, y):
    tp = 2.0 * np.pi
    return (30.0 * np.sin(tp * x / 2000.0 + .3) * np.cos(tp * y / 1700.0 + 1.1)
            + 8.0 * np.sin(tp * x / 400.0 + 2.0) * np.sin(tp * y / 370.0 + .7)
            + 2.0 * np.sin(tp * x / 80.0 + .5) * np.cos(tp * y / 90.0 + .2)
            + 0.02 * x + 100.0)


def _lattice_rand(ix, iy, salt):
    """Stateless per-lattice-cell uniform [0,1) (so any sub-region of the plane
    generates the same buildings regardless of how points are sliced)."""
    h = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
         ^ (iy.astype(np.uint64) + np.uint64(salt)) * np.uint64(0xC2B2AE3D27D4EB4F))
    h ^= h >> np.uint64(29)
    h *= np.uint64(0xBF58476D1CE4E5B9)
    h ^= h >> np.uint64(32)
    h *= np.uint64(0x94D049BB133111EB)
    h ^= h >> np.uint64(29)
    return (h >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def synth_cloud(n, ex, ey, seed=0, x0=0.0, y0=0.0, dtype=np.float32):
    """terrain + one flat-roof building per 120 m lattice cell + vegetation
    patches + 1e-4 low outliers; x, y, z rounded to `dtype` (float32 by default so
    that a float4 stream and the float64 reference see identical numbers).
    Returns x, y, z (float64 arrays holding dtype-representable values) and the
    generator's own object label (1 = building/vegetation/outlier)."""
    rng = np.random.default_rng(seed)
    x = rng.random(n) * ex + x0
    y = rng.random(n) * ey + y0
    z = terrain(x, y) + rng.normal(0.0, 0.03, n)
    with np.errstate(over='ignore'):
        ix = np.floor(x / 120.0).astype(np.int64)
        iy = np.floor(y / 120.0).astype(np.int64)
        bx0 = ix * 120.0 + 10.0 + 50.0 * _lattice_rand(ix, iy, 1)
        by0 = iy * 120.0 + 10.0 + 50.0 * _lattice_rand(ix, iy, 2)
        bw = 8.0 + 52.0 * _lattice_rand(ix, iy, 3)
        bd = 8.0 + 52.0 * _lattice_rand(ix, iy, 4)
        bh = 3.0 + 27.0 * _lattice_rand(ix, iy, 5)
    inb = (x >= bx0) & (x < bx0 + bw) & (y >= by0) & (y < by0 + bd)
    roof = terrain(bx0 + .5 * bw, by0 + .5 * bd) + bh
    z = np.where(inb, roof, z)
    tp = 2.0 * np.pi
    veg = (np.sin(tp * x / 310.0 + 1.0) * np.sin(tp * y / 270.0 + 2.0) > 0.35) & ~inb
    lifted = veg & (rng.random(n) < 0.6)
    z = z + np.where(lifted, 0.3 + 24.7 * rng.random(n), 0.0)
    low = rng.random(n) < 1e-4
    z = z - np.where(low, 5.0 + 45.0 * rng.random(n), 0.0)
    label = (inb | lifted | low).astype(np.uint8)
    x = x.astype(dtype).astype(np.float64)
    y = y.astype(dtype).astype(np.float64)
    z = z.astype(dtype).astype(np.float64)
    return x, y, z, label


def synth_dem(ny, nx, seed=3, nan_frac=0.3, dtype=np.float32):
    """config 3: terrain + buildings sampled on a unit lattice as `dtype`, with
    `nan_frac` of the cells set NaN."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(ny, dtype=np.float64), np.arange(nx, dtype=np.float64), indexing='ij')
    z = terrain(xx, yy)
    with np.errstate(over='ignore'):
        ix = np.floor(xx / 120.0).astype(np.int64)
        iy = np.floor(yy / 120.0).astype(np.int64)
        bx0 = ix * 120.0 + 10.0 + 50.0 * _lattice_rand(ix, iy, 1)
        by0 = iy * 120.0 + 10.0 + 50.0 * _lattice_rand(ix, iy, 2)
        bw = 8.0 + 52.0 * _lattice_rand(ix, iy, 3)
        bd = 8.0 + 52.0 * _lattice_rand(ix, iy, 4)
        bh = 3.0 + 27.0 * _lattice_rand(ix, iy, 5)
    inb = (xx >= bx0) & (xx < bx0 + bw) & (yy >= by0) & (yy < by0 + bd)
    z = np.where(inb, terrain(bx0 + .5 * bw, by0 + .5 * bd) + bh, z)
    z = z + rng.normal(0.0, 0.03, z.shape)
    z = z.astype(dtype)
    if nan_frac > 0:
        z[rng.random(z.shape) < nan_frac] = np.nan
    return z
