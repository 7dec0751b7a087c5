"""Size-independent properties at (near) BASELINE.json sizes, where the CPU oracle cannot go.

  binning        : equals an independent scatter-min of the same float64 cell indices (50 M points)
  opening        : idempotent, anti-extensive, increasing; marching kernel == direct kernel (8192^2)
  progressive    : the mask is monotone in the slope threshold
  inpaint        : known cells untouched, discrete Laplacian of the fill <= tolerance (16 M unknowns)
  smrf           : invariant under a permutation of the points (20 M points)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env():
    import torch
    import neilpy_b200 as nb
    from neilpy_b200.synth_torch import dem_on_device
    return torch, nb, dem_on_device


def test_binning_50m_points_equals_scatter_min(env):
    torch, nb, _ = env
    from bench import make_cloud
    pts = torch.from_numpy(make_cloud(50_000_000, 3)).cuda()
    Z, t = nb.create_dem(pts, None, None, 1, 'min')
    ny, nx = Z.shape
    inv = ~nb.Affine(*tuple(t)[:6])
    x, y, z = pts[:, 0].double(), pts[:, 1].double(), pts[:, 2]
    c = torch.floor(x * inv.a + y * inv.b + inv.c).long()
    r = torch.floor(x * inv.d + y * inv.e + inv.f).long()
    assert int(c.min()) >= 0 and int(c.max()) < nx and int(r.min()) >= 0 and int(r.max()) < ny
    ref = torch.full((ny * nx,), float('inf'), dtype=torch.float32, device='cuda')
    ref.scatter_reduce_(0, r * nx + c, z, 'amin')
    ref = ref.view(ny, nx)
    empty = torch.isinf(ref)
    assert torch.equal(torch.isnan(Z), empty)
    assert torch.equal(Z[~empty], ref[~empty])
    assert 0.10 < float(empty.float().mean()) < 0.17          # Poisson(2): e^-2 = 13.5 % empty cells


@pytest.mark.parametrize('w', [1, 4, 9, 18, 30])
def test_opening_properties_8192(env, w):
    torch, nb, dem = env
    from test_gpu_stages import open_window
    from neilpy_b200 import _lib
    from neilpy_b200.api import _ptr, _stream, _code
    Z = dem(torch, 8192, 8192, torch.device('cuda'), seed=w)
    lib = _lib.load()

    def opening(A, thr=0.0):
        out, tmp = torch.empty_like(A), torch.empty_like(A)
        mask = torch.zeros(A.shape, dtype=torch.uint8, device='cuda')
        _lib.check(lib.smrf_open_window(_ptr(A), _ptr(out), _ptr(tmp), _ptr(mask), None, A.shape[0], A.shape[1],
                                        A.shape[1], _code(A.dtype), w, float(thr), 0, 0, 0, A.shape[0], _stream()), 'open')
        return out, mask

    O1, m1 = opening(Z, 0.15 * w)
    assert bool((O1 <= Z).all())                                   # anti-extensive
    O2, _ = opening(O1)
    assert torch.equal(O1, O2)                                     # idempotent
    O3, _ = opening(Z + 1.0)                                       # increasing + translation-compatible
    assert bool((O3 >= O1).all())
    assert torch.equal(m1.bool(), (Z.double() - O1.double()) > 0.15 * w)
    # every opened value is a value of the input (min/max never invent numbers)
    sample = O1[::257, ::263].flatten()
    zs = torch.sort(Z.flatten()).values
    idx = torch.searchsorted(zs, sample).clamp(max=zs.numel() - 1)
    assert torch.equal(zs[idx], sample)


def test_march_equals_direct_kernel_on_a_large_grid(env, monkeypatch):
    torch, nb, dem = env
    Z = dem(torch, 4096, 6144, torch.device('cuda'), seed=9)
    a = nb.progressive_filter(Z, np.array([1, 2, 3, 5]), 1, .15)
    monkeypatch.setenv('SMRF_OPEN_IMPL', 'generic')
    b = nb.progressive_filter(Z, np.array([1, 2, 3, 5]), 1, .15)
    assert torch.equal(a, b)


def test_progressive_mask_is_monotone_in_the_threshold(env):
    torch, nb, dem = env
    Z = dem(torch, 8192, 8192, torch.device('cuda'), seed=2)
    w = np.arange(18) + 1
    loose, tight = nb.progressive_filter(Z, w, 1, .30), nb.progressive_filter(Z, w, 1, .15)
    assert bool((tight | ~loose).all()) and int(tight.sum()) > int(loose.sum()) > 0


def test_inpaint_large_grid_is_harmonic(env):
    torch, nb, dem = env
    dev = torch.device('cuda')
    Z = dem(torch, 8192, 8192, dev, seed=4).double()
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    Z[torch.rand(Z.shape, generator=g, device=dev) < 0.2] = float('nan')
    Z[1000:1200, 3000:3300] = float('nan')
    Z[:150, :90] = float('nan')
    unk = torch.isnan(Z)
    F, info = nb.inpaint_nans_by_springs(Z, return_info=True)
    assert not bool(torch.isnan(F).any()) and torch.equal(F[~unk], Z[~unk])
    P = torch.nn.functional.pad(F[None, None], (1, 1, 1, 1), mode='replicate')[0, 0]     # natural boundary
    lap = P[:-2, 1:-1] + P[2:, 1:-1] + P[1:-1, :-2] + P[1:-1, 2:] - 4 * F
    assert float(lap[unk].abs().max()) <= 2e-9, info
    assert info['iterations'] < 60


def test_smrf_is_invariant_under_point_order(env):
    torch, nb, _ = env
    from bench import make_cloud, PARAMS
    pts = torch.from_numpy(make_cloud(20_000_000, 5)).cuda()
    perm = torch.randperm(pts.shape[0], device='cuda')
    Z1, t1, oc1, op1 = nb.smrf(pts, **PARAMS)
    Z2, t2, oc2, op2 = nb.smrf(pts[perm].contiguous(), **PARAMS)
    assert tuple(t1) == tuple(t2)
    # binning and the opening are order-independent bit for bit; the solver's reductions are
    # atomic, so the DTM agrees to rounding and a threshold-marginal point may flip
    assert int((oc1 != oc2).sum()) == 0
    assert float((Z1 - Z2).abs().max()) <= 1e-4
    assert int((op1[perm] != op2).sum()) <= 20


@pytest.mark.parametrize('w', [1, 4, 18])
def test_opening_2048_against_the_oracle(w):
    """One full-width grid against the oracle itself (not a property): 2048 x 2048 cells, five column strips and
    several row segments of the marching kernel."""
    import torch
    from neilpy_b200 import _lib
    from neilpy_b200.api import _ptr, _stream, _code
    from neilpy_b200.synth import synth_dem
    from oracle import smrf_oracle as O
    lib = _lib.load()
    Z = synth_dem(2048, 2048, seed=3, nan_frac=0.0).astype(np.float32)
    Z += np.random.default_rng(1).normal(0, 0.2, Z.shape).astype(np.float32)
    ref = O.opening(Z.astype(np.float64), O.disk(w))
    zin = torch.as_tensor(Z).cuda()
    out, tmp = torch.empty_like(zin), torch.empty_like(zin)
    mask = torch.zeros(zin.shape, dtype=torch.uint8, device='cuda')
    _lib.check(lib.smrf_open_window(_ptr(zin), _ptr(out), _ptr(tmp), _ptr(mask), None, 2048, 2048, 2048, _code(zin.dtype), w,
                                    0.15 * w, 0, 0, 0, 2048, _stream()), 'smrf_open_window')
    assert np.array_equal(out.cpu().numpy().astype(np.float64), ref)
    assert np.array_equal(mask.cpu().numpy().astype(bool), (Z.astype(np.float64) - ref) > 0.15 * w)
