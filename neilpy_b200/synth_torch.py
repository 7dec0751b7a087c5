"""Device-side generator for grid-only benchmark surfaces (BASELINE.json configs[2]/[4]):
the terrain field of neilpy_b200.synth plus one flat-roof building per 120-cell lattice
cell and 3 cm noise, produced in row chunks so that a 32768 x 32768 float32 surface never
needs more than its own 4.3 GB.  It is a workload generator only (no oracle comparison is
made at that size; parity there rests on size-independent properties)."""
import math


def dem_on_device(torch, ny, nx, dev, seed=3, chunk=1024, row0=0, scale=1.0):
    """rows [row0, row0 + ny) of the surface; `scale` = cell size in metres (the field is defined in metres)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 7919 * row0)
    out = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    tp = 2.0 * math.pi
    xx = torch.arange(nx, dtype=torch.float32, device=dev)[None, :] * scale

    def terrain(x, y):
        return (30.0 * torch.sin(tp * x / 2000.0 + .3) * torch.cos(tp * y / 1700.0 + 1.1)
                + 8.0 * torch.sin(tp * x / 400.0 + 2.0) * torch.sin(tp * y / 370.0 + .7)
                + 2.0 * torch.sin(tp * x / 80.0 + .5) * torch.cos(tp * y / 90.0 + .2) + 0.02 * x + 100.0)

    def lat(ix, iy, salt):
        v = torch.sin(ix * 12.9898 + iy * 78.233 + salt * 37.719) * 43758.5453
        return v - torch.floor(v)

    for r0 in range(0, ny, chunk):
        r1 = min(ny, r0 + chunk)
        yy = torch.arange(row0 + r0, row0 + r1, dtype=torch.float32, device=dev)[:, None] * scale
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


def cloud_on_device(torch, n, ex, ey, dev, seed=0, chunk=1 << 24):
    """The point-cloud generator of neilpy_b200.synth.synth_cloud evaluated on the device (BASELINE.json
    configs[3]: 125 M points per GPU would take minutes in numpy): same terrain, building lattice, vegetation
    patches, noise and low outliers, float32 x, y, z, as an (n, 4) float4 stream.  The random streams differ
    from numpy's, so it is a workload generator, not a fixture."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((n, 4), dtype=torch.float32, device=dev)
    tp = 2.0 * math.pi

    def terrain(x, y):
        return (30.0 * torch.sin(tp * x / 2000.0 + .3) * torch.cos(tp * y / 1700.0 + 1.1)
                + 8.0 * torch.sin(tp * x / 400.0 + 2.0) * torch.sin(tp * y / 370.0 + .7)
                + 2.0 * torch.sin(tp * x / 80.0 + .5) * torch.cos(tp * y / 90.0 + .2) + 0.02 * x + 100.0)

    def lat(ix, iy, salt):
        v = torch.sin(ix * 12.9898 + iy * 78.233 + salt * 37.719) * 43758.5453
        return v - torch.floor(v)

    for a in range(0, n, chunk):
        m = min(chunk, n - a)
        r = lambda: torch.rand(m, generator=g, device=dev, dtype=torch.float64)
        x, y = r() * ex, r() * ey
        z = terrain(x, y) + 0.03 * torch.randn(m, generator=g, device=dev, dtype=torch.float64)
        ix, iy = torch.floor(x / 120.0), torch.floor(y / 120.0)
        bx0 = ix * 120.0 + 10.0 + 50.0 * lat(ix, iy, 1.0)
        by0 = iy * 120.0 + 10.0 + 50.0 * lat(ix, iy, 2.0)
        bw, bd = 8.0 + 52.0 * lat(ix, iy, 3.0), 8.0 + 52.0 * lat(ix, iy, 4.0)
        bh = 3.0 + 27.0 * lat(ix, iy, 5.0)
        inb = (x >= bx0) & (x < bx0 + bw) & (y >= by0) & (y < by0 + bd)
        z = torch.where(inb, terrain(bx0 + .5 * bw, by0 + .5 * bd) + bh, z)
        veg = (torch.sin(tp * x / 310.0 + 1.0) * torch.sin(tp * y / 270.0 + 2.0) > 0.35) & ~inb
        lifted = veg & (r() < 0.6)
        z = z + torch.where(lifted, 0.3 + 24.7 * r(), torch.zeros_like(z))
        low = r() < 1e-4
        z = z - torch.where(low, 5.0 + 45.0 * r(), torch.zeros_like(z))
        out[a:a + m, 0], out[a:a + m, 1], out[a:a + m, 2], out[a:a + m, 3] = x.float(), y.float(), z.float(), 0.0
    return out
