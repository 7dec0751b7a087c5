"""The row-band sharded path equals the unsharded path -- on ONE GPU.

neilpy_b200/distributed.py runs over a small Comm interface (neilpy_b200/comm.py); ThreadComm
carries K virtual ranks as K threads of this process on one device, so the band code that NCCL
carries on the 8-GPU box (halo exchanges, the ghost-extended global V-cycle, grouped windows,
all-reduced dot products, gathered spline coefficients) is executed unchanged here and compared
with `neilpy_b200.smrf` on the whole cloud.  (tests/test_gpu_distributed.py repeats it over NCCL
when the box has two GPUs.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KW = dict(cellsize=1, windows=18, slope_threshold=.15, elevation_threshold=.5, elevation_scaler=1.25)


def _sharded(xyzw, world, kw, slices=None):
    import torch
    from neilpy_b200.comm import run_virtual_ranks
    from neilpy_b200.distributed import smrf_sharded
    dev = torch.device('cuda', torch.cuda.current_device())
    if slices is None:
        slices = [xyzw[r::world] for r in range(world)]
    parts = [torch.as_tensor(np.ascontiguousarray(s)).to(dev) for s in slices]
    return run_virtual_ranks(world, lambda comm: smrf_sharded(parts[comm.rank], gather=True, comm=comm, **kw), dev), slices


# (bands, extent): two and four bands; eight bands on a grid whose last band (1417 - 7 * 192 = 73 rows) cannot serve
# the 128 ghost rows of the preferred multigrid split, so that the (3, 64) plan runs; eight bands with the (4, 128) plan
@pytest.mark.parametrize('world,ex,ey', [(2, 500.0, 700.0), (4, 500.0, 700.0), (8, 300.0, 1416.0), (8, 250.0, 2100.0)])
def test_virtual_bands_equal_unsharded(world, ex, ey):
    import torch
    import neilpy_b200 as nb
    from neilpy_b200.distributed import mg_plan, MG_PLANS
    from neilpy_b200.synth import synth_cloud
    x, y, z, _ = synth_cloud(int(ex * ey * 2), ex, ey, seed=11)
    assert mg_plan(int(ey) + 1, world) == (MG_PLANS[1] if ey == 1416.0 else MG_PLANS[0])
    xyzw = np.stack([x, y, z, np.zeros_like(x)], 1).astype(np.float32)
    st = {}
    Z1, t1, oc1, op1 = nb.smrf(torch.as_tensor(xyzw).cuda(), return_stages=st, **KW)
    res, _ = _sharded(xyzw, world, KW)
    op = torch.empty_like(op1)
    for r in range(world):
        assert tuple(res[r]['t'])[:6] == tuple(t1)[:6] and res[r]['shape'] == tuple(Z1.shape)
        op[r::world] = res[r]['is_object_point']
    dz = float((res[0]['Zpro'] - Z1).abs().max())
    cell_flips = int((res[0]['object_cells'] != oc1).sum())
    point_flips = int((op != op1).sum())
    its = (res[0]['info']['inpaint1']['iterations'], res[0]['info']['inpaint2']['iterations'])
    print('world %d: max|dZ| %.3g m, cell flips %d, point flips %d, CG iterations %s vs %s'
          % (world, dz, cell_flips, point_flips, its, (st['inpaint1']['iterations'], st['inpaint2']['iterations'])))
    # every rank holds the same gathered grids
    for r in range(1, world):
        assert torch.equal(res[r]['Zpro'], res[0]['Zpro']) and torch.equal(res[r]['object_cells'], res[0]['object_cells'])
    assert cell_flips == 0 and point_flips == 0
    assert dz <= 1e-4                       # float32 grid at ~150 m: one ulp is 1.5e-5 m
    assert its == (st['inpaint1']['iterations'], st['inpaint2']['iterations'])


def test_virtual_bands_with_an_empty_rank_and_float64_points():
    """A rank without points takes part in every collective; float64 SoA input."""
    import torch
    import neilpy_b200 as nb
    from neilpy_b200.synth import synth_cloud
    from neilpy_b200.comm import run_virtual_ranks
    from neilpy_b200.distributed import smrf_sharded
    x, y, z, _ = synth_cloud(300000, 400.0, 420.0, seed=4, dtype=np.float64)
    x, y = x + 500000.0, y + 5400000.0
    kw = dict(KW, windows=8)
    Z1, t1, oc1, op1 = nb.smrf(torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda(), torch.as_tensor(z).cuda(), **kw)
    cuts = [0, 200000, 200000, 300000]                 # rank 1 holds nothing
    parts = [tuple(torch.as_tensor(v[cuts[r]:cuts[r + 1]].copy()).cuda() for v in (x, y, z)) for r in range(3)]
    dev = torch.device('cuda', torch.cuda.current_device())
    res = run_virtual_ranks(3, lambda comm: smrf_sharded(parts[comm.rank], gather=True, comm=comm, **kw), dev)
    op = torch.cat([res[r]['is_object_point'] for r in range(3)])
    assert res[1]['is_object_point'].numel() == 0
    assert int((res[0]['object_cells'] != oc1).sum()) == 0 and int((op != op1).sum()) == 0
    assert float((res[0]['Zpro'] - Z1).abs().max()) <= 1e-6


def test_out_of_range_and_non_finite_points_raise_on_every_rank():
    import torch
    from neilpy_b200.comm import run_virtual_ranks
    from neilpy_b200.distributed import smrf_sharded
    from neilpy_b200.synth import synth_cloud
    x, y, z, _ = synth_cloud(100000, 300.0, 320.0, seed=5)
    xyzw = np.stack([x, y, z, np.zeros_like(x)], 1).astype(np.float32)
    xyzw[7, 0] = np.nan
    dev = torch.device('cuda', torch.cuda.current_device())
    parts = [torch.as_tensor(xyzw[r::2].copy()).cuda() for r in range(2)]
    with pytest.raises(ValueError):
        run_virtual_ranks(2, lambda comm: smrf_sharded(parts[comm.rank], comm=comm, **KW), dev)
