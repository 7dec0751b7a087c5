"""Device-side generator for grid-only benchmark surfaces (BASELINE.json configs[2]/[4]):
the terrain field of neilpy_b200.synth plus one flat-roof building per 120-cell lattice
cell and 3 cm noise, produced in row chunks so that a 32768 x 32768 float32 surface never
needs more than its own 4.3 GB.  It is a workload generator only (no oracle comparison is
made at that size; parity there rests on size-independent properties)."""
import math


def dem_on_device(torch, ny, nx, dev, seed=3, chunk=1024):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    tp = 2.0 * math.pi
    xx = torch.arange(nx, dtype=torch.float32, device=dev)[None, :]

    def terrain(x, y):
        return (30.0 * torch.sin(tp * x / 2000.0 + .3) * torch.cos(tp * y / 1700.0 + 1.1)
                + 8.0 * torch.sin(tp * x / 400.0 + 2.0) * torch.sin(tp * y / 370.0 + .7)
                + 2.0 * torch.sin(tp * x / 80.0 + .5) * torch.cos(tp * y / 90.0 + .2) + 0.02 * x + 100.0)

    def lat(ix, iy, salt):
        v = torch.sin(ix * 12.9898 + iy * 78.233 + salt * 37.719) * 43758.5453
        return v - torch.floor(v)

    for r0 in range(0, ny, chunk):
        r1 = min(ny, r0 + chunk)
        yy = torch.arange(r0, r1, dtype=torch.float32, device=dev)[:, None]
        z = terrain(xx, yy)
        ix, iy = torch.floor(xx / 120.0), torch.floor(yy / 120.0)
        bx0 = ix * 120.0 + 10.0 + 50.0 * lat(ix, iy, 1.0)
        by0 = iy * 120.0 + 10.0 + 50.0 * lat(ix, iy, 2.0)
        bw, bd = 8.0 + 52.0 * lat(ix, iy, 3.0), 8.0 + 52.0 * lat(ix, iy, 4.0)
        bh = 3.0 + 27.0 * lat(ix, iy, 5.0)
        inb = (xx >= bx0) & (xx < bx0 + bw) & (yy >= by0) & (yy < by0 + bd)
        z = torch.where(inb, terrain(bx0 + .5 * bw, by0 + .5 * bd) + bh, z)
        z += 0.03 * torch.randn(z.shape, generator=g, device=dev, dtype=torch.float32)
        out[r0:r1] = z
    return out
