"""Stage-wise parity of the CUDA path against the oracle, through the C ABI (needs a B200).

Every stage is fed the oracle's upstream output, so a difference is that stage's own.
Bars: binning / opening / masks bit-exact; inpaint <= 1e-6 m of the exact harmonic fill and
<= 2e-2 m of the reference's (inexact) LSQR; slope / spline / interpolation <= 1e-9 m in
float64 and <= 2e-4 m with float32 grids.
"""
import numpy as np
import pytest
import scipy.ndimage as ndi
from scipy import interpolate

from conftest import load_isprs
from oracle import smrf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nb():
    import neilpy_b200
    return neilpy_b200


def eq_nan(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


# ------------------------------------------------------------------ create_dem
@pytest.mark.parametrize('name', ['samp12', 'samp53'])
@pytest.mark.parametrize('bin_type', ['min', 'max'])
def test_create_dem_isprs_bit_exact(nb, name, bin_type):
    x, y, z, _ = load_isprs(name)
    I0, t0 = O.create_dem(x, y, z, cellsize=1, bin_type=bin_type)
    I1, t1 = nb.create_dem(x, y, z, cellsize=1, bin_type=bin_type)
    assert I1.dtype == np.float64 and tuple(t1)[:6] == t0.coeffs
    assert eq_nan(I0, I1)


@pytest.mark.parametrize('cs', [1, 0.5, 0.1, 2.0])
def test_create_dem_cellsizes_and_utm_offsets(nb, cs):
    x, y, z, _ = O.synth_cloud(200000, 400.0, 300.0, seed=5, dtype=np.float64)
    x, y = x + 500000.0, y + 5400000.0          # float64-only coordinates
    I0, t0 = O.create_dem(x, y, z, cellsize=cs, bin_type='min')
    I1, t1 = nb.create_dem(x, y, z, cellsize=cs, bin_type='min')
    assert tuple(t1)[:6] == t0.coeffs and eq_nan(I0, I1)


def test_create_dem_float32_streams(nb):
    import torch
    x, y, z, _ = O.synth_cloud(300000, 500.0, 350.0, seed=6)       # float32-representable values
    I0, t0 = O.create_dem(x, y, z, cellsize=0.5, bin_type='min')
    xyzw = np.stack([x, y, z, np.zeros_like(x)], 1).astype(np.float32)
    I1, t1 = nb.create_dem(xyzw, None, None, cellsize=0.5, bin_type='min')
    assert I1.dtype == np.float32 and eq_nan(I0, I1.astype(np.float64)) and tuple(t1)[:6] == t0.coeffs
    I2, _ = nb.create_dem(x.astype(np.float32), y.astype(np.float32), z.astype(np.float32), cellsize=0.5, bin_type='min')
    assert eq_nan(I0, I2.astype(np.float64))
    dev = torch.as_tensor(xyzw).cuda()
    I3, _ = nb.create_dem(dev, None, None, cellsize=0.5, bin_type='min')
    assert I3.is_cuda and eq_nan(I0, I3.cpu().numpy().astype(np.float64))


def test_create_dem_edge_rule_nan_z_and_errors(nb):
    x = np.array([0.0, 0.5, 3.0, 3.0]); y = np.array([0.0, 0.5, 2.0, 2.0]); z = np.array([1.0, 2.0, 3.0, np.nan])
    I0, _ = O.create_dem(x, y, z, cellsize=1, bin_type='min')
    I1, _ = nb.create_dem(x, y, z, cellsize=1, bin_type='min')
    assert eq_nan(I0, I1)
    with pytest.raises(ValueError, match='This type not supported.'):
        nb.create_dem(x, y, z, bin_type='median')
    with pytest.raises(ValueError):
        nb.create_dem(np.array([0.0, np.nan]), np.array([0.0, 1.0]), np.array([0.0, 1.0]))
    Ii, _ = nb.create_dem(x, y, z, cellsize=1, bin_type='min', inpaint=True)
    assert not np.isnan(Ii).any()


def test_create_dem_with_explicit_edges(nb):
    """edges=(xedges, yedges): points outside are dropped first (neilpy.py:1125-1132), cellsize comes from the edges"""
    x, y, z, _ = O.synth_cloud(100000, 300.0, 200.0, seed=8, dtype=np.float64)
    xedges = np.arange(40.0, 261.0, 2.0)
    yedges = np.arange(170.0, 29.0, -2.0)
    for bin_type in ('min', 'max'):
        I0, t0 = O.create_dem(x, y, z, cellsize=1, bin_type=bin_type, edges=(xedges, yedges))
        I1, t1 = nb.create_dem(x, y, z, cellsize=1, bin_type=bin_type, edges=(xedges, yedges))
        assert tuple(t1)[:6] == t0.coeffs and eq_nan(I0, I1) and I1.shape == (len(yedges) - 1, len(xedges) - 1)


# ------------------------------------------------------------------ progressive_filter
def surface(ny, nx, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(ny, dtype=np.float64), np.arange(nx, dtype=np.float64), indexing='ij')
    z = O.terrain(xx * 3, yy * 3) + rng.normal(0, 0.3, (ny, nx))
    for _ in range(max(2, ny * nx // 1500)):
        r0, c0 = rng.integers(0, ny), rng.integers(0, nx)
        h, w = rng.integers(2, 25), rng.integers(2, 25)
        z[r0:r0 + h, c0:c0 + w] += rng.uniform(2, 20)
    return z.astype(dtype)


def open_window(nb, Z, w, thr=0.0, negate=0, rows=None):
    """one smrf_open_window call through the C ABI; returns (this, mask)"""
    import torch
    from neilpy_b200 import _lib
    from neilpy_b200.api import _ptr, _stream, _code
    lib = _lib.load()
    zin = torch.as_tensor(np.ascontiguousarray(Z)).cuda()
    out = torch.full_like(zin, float('nan'))
    tmp = torch.empty_like(zin)
    mask = torch.zeros(zin.shape, dtype=torch.uint8, device='cuda')
    ny, nx = zin.shape
    lo, hi = rows if rows else (0, ny)
    _lib.check(lib.smrf_open_window(_ptr(zin), _ptr(out), _ptr(tmp), _ptr(mask), None, ny, nx, nx, _code(zin.dtype), w,
                                    float(thr), 0, negate, lo, hi, _stream()), 'smrf_open_window')
    torch.cuda.synchronize()
    return out.cpu().numpy(), mask.cpu().numpy().astype(bool)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('w', [1, 2, 3, 4, 5, 6, 7, 8, 11, 13, 16, 17, 18, 19, 20, 21, 25, 33, 36, 40, 41, 44, 48, 56, 64, 72])
def test_single_opening_bit_exact(nb, dtype, w):
    # 1000 columns: two column strips of the marching kernel; 300 rows: several row segments
    Z = surface(300, 1000, w, dtype)
    ref = O.opening(Z.astype(np.float64), O.disk(w))
    got, mask = open_window(nb, Z, w, thr=0.15 * w)
    assert got.dtype == dtype and np.array_equal(got.astype(np.float64), ref)
    assert np.array_equal(mask, (Z.astype(np.float64) - ref) > 0.15 * w)


@pytest.mark.parametrize('shape', [(37, 41), (64, 513), (90, 477), (130, 953), (200, 1431), (75, 2), (3, 300), (1, 1)])
@pytest.mark.parametrize('w', [1, 3, 9, 18])
def test_opening_odd_shapes_against_the_border_rule(nb, shape, w):
    """widths that are not multiples of 4, narrower than a strip, or smaller than the disk.
    Checked against the ignore-out-of-image rule itself (scipy's reflect is only equivalent
    to it while the grid is larger than the radius -- SURVEY F5)."""
    Z = surface(shape[0], shape[1], sum(shape) + w, np.float32)
    Zd = Z.astype(np.float64)
    pad = np.pad(Zd, w, constant_values=np.inf)
    er = ndi.grey_erosion(pad, footprint=O.disk(w), mode='constant', cval=np.inf)[w:-w, w:-w]
    padm = np.pad(er, w, constant_values=-np.inf)
    ref = ndi.grey_dilation(padm, footprint=O.disk(w), mode='constant', cval=-np.inf)[w:-w, w:-w]
    got, _ = open_window(nb, Z, w)
    assert np.array_equal(got.astype(np.float64), ref)
    if min(shape) > 2 * w:
        assert np.array_equal(ref, O.opening(Zd, O.disk(w)))


def test_march_and_direct_kernels_agree(nb, monkeypatch):
    Z = surface(257, 1203, 99, np.float32)
    a = {w: open_window(nb, Z, w, thr=0.1 * w) for w in (1, 5, 12, 18, 23, 37)}
    monkeypatch.setenv('SMRF_OPEN_IMPL', 'generic')
    for w, (s, m) in a.items():
        s2, m2 = open_window(nb, Z, w, thr=0.1 * w)
        assert np.array_equal(s, s2) and np.array_equal(m, m2)


def test_tma_and_cp_async_loaders_agree(nb, monkeypatch):
    """rows that the TMA unit can describe (16-byte aligned) take the TMA loader; SMRF_OPEN_NO_TMA=1 forces
    the cp.async loader of the same kernel.  Identical results, including at every image border."""
    Z = surface(211, 1500, 5, np.float32)
    a = {w: open_window(nb, Z, w, thr=0.1 * w) for w in (1, 2, 4, 9, 18, 30, 47, 72)}
    monkeypatch.setenv('SMRF_OPEN_NO_TMA', '1')
    for w, (s, m) in a.items():
        s2, m2 = open_window(nb, Z, w, thr=0.1 * w)
        assert np.array_equal(s, s2) and np.array_equal(m, m2), w


def test_float64_rank_space_equals_the_direct_kernels(nb, monkeypatch):
    """float64 surfaces are opened in rank space (csrc/rank.cu): bit-identical to the brute-force float64
    kernels, with ties, negative values, signed zeros and NaN cells in the surface."""
    Z = surface(190, 700, 21, np.float64)
    Z[40:60, 100:160] = Z[40, 100]            # ties
    Z[100:110, 300:320] *= -1.0
    Z[5, 7] = 0.0
    Z[5, 8] = -0.0
    Z[150, 650] = np.nan
    windows = np.array([1, 2, 3, 5, 9, 18])
    m1, w1 = nb.progressive_filter(Z, windows, 1, .15, return_when_dropped=True)
    monkeypatch.setenv('SMRF_OPEN_IMPL', 'generic')
    m2, w2 = nb.progressive_filter(Z, windows, 1, .15, return_when_dropped=True)
    assert np.array_equal(m1, m2) and np.array_equal(w1, w2)


def test_open_window_row_band(nb):
    """row-band sharding: only rows [lo, hi) are written, from halo rows that are inputs only"""
    Z = surface(400, 640, 7, np.float32)
    w = 9
    full, fm = open_window(nb, Z, w, thr=0.5)
    lo, hi = 120, 250
    part, pm = open_window(nb, Z, w, thr=0.5, rows=(lo, hi))
    assert np.array_equal(part[lo:hi], full[lo:hi]) and np.array_equal(pm[lo:hi], fm[lo:hi])
    assert np.isnan(part[:lo]).all() and np.isnan(part[hi:]).all() and not pm[:lo].any() and not pm[hi:].any()
    # a band cut out with 2w halo rows gives the same interior
    b0, b1 = lo - 2 * w, hi + 2 * w
    band, bm = open_window(nb, Z[b0:b1], w, thr=0.5, rows=(2 * w, 2 * w + hi - lo))
    assert np.array_equal(band[2 * w:2 * w + hi - lo], full[lo:hi]) and np.array_equal(bm[2 * w:2 * w + hi - lo], fm[lo:hi])


@pytest.mark.parametrize('w', [1, 2, 7, 13, 18, 30, 45])
def test_open_window_with_padded_rows(nb, w):
    """row stride > nx (what smrf_progressive_open uses internally): same result as contiguous rows"""
    import torch
    from neilpy_b200 import _lib
    from neilpy_b200.api import _ptr, _stream, _code
    lib = _lib.load()
    Z = surface(150, 1001, w + 50, np.float32)                  # nx = 1001: not a multiple of 4
    ref, rmask = open_window(nb, Z, w, thr=0.1 * w)
    ny, nx, pitch = 150, 1001, 1004
    zin = torch.full((ny, pitch), 1e30, dtype=torch.float32, device='cuda')    # poison in the padding
    zin[:, :nx] = torch.as_tensor(Z).cuda()
    out = torch.full((ny, pitch), -7.0, dtype=torch.float32, device='cuda')
    tmp = torch.empty_like(out)
    mask = torch.zeros((ny, nx), dtype=torch.uint8, device='cuda')
    _lib.check(lib.smrf_open_window(_ptr(zin), _ptr(out), _ptr(tmp), _ptr(mask), None, ny, nx, pitch, _code(zin.dtype),
                                    w, 0.1 * w, 0, 0, 0, ny, _stream()), 'smrf_open_window')
    torch.cuda.synchronize()
    assert np.array_equal(out[:, :nx].cpu().numpy(), ref) and np.array_equal(mask.cpu().numpy().astype(bool), rmask)
    assert bool((out[:, nx:] == -7.0).all())                    # the padding is never written


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_progressive_filter_matches_oracle(nb, dtype):
    Z = surface(333, 700, 3, dtype)
    windows = np.arange(18) + 1
    m0, w0 = O.progressive_filter(Z.astype(np.float64), windows, 1, .15, return_when_dropped=True)
    m1, w1 = nb.progressive_filter(Z, windows, 1, .15, return_when_dropped=True)
    assert m1.dtype == np.bool_ and w1.dtype == np.uint8
    assert np.array_equal(m0, m1) and np.array_equal(w0, w1)
    # arbitrary radii, non-unit cellsize, and the low-outlier form (negated surface, one window)
    wl = np.array([2, 5, 9])
    assert np.array_equal(O.progressive_filter(Z.astype(np.float64), wl, 0.5, .3), nb.progressive_filter(Z, wl, 0.5, .3))
    one = np.array([1])
    Zp = Z.copy()
    Zp[50, 60] -= 30; Zp[200, 333] -= 8; Zp[0, 0] -= 12
    lo0 = O.progressive_filter(-Zp.astype(np.float64), one, 1, 5)
    import torch
    from neilpy_b200 import _lib
    from neilpy_b200.api import _progressive
    zt = torch.as_tensor(Zp).cuda()
    mk = torch.zeros(zt.shape, dtype=torch.uint8, device='cuda')
    _progressive(_lib.load(), zt, one, 5 * (one * 1), mk, None, None, negate=1)
    assert lo0.sum() >= 3 and np.array_equal(lo0, mk.cpu().numpy().astype(bool))
    assert np.array_equal(zt.cpu().numpy(), Zp)           # the input surface is never written


def test_progressive_filter_isprs_grid(nb):
    x, y, z, _ = load_isprs('samp12')
    st = {}
    O.smrf(x, y, z, 1, 18, .15, .5, 1.25, stages=st)
    got = nb.progressive_filter(st['Zmin_filtered'], np.arange(18) + 1, 1, .15)
    assert np.array_equal(got, st['progressive_cells'])


# ------------------------------------------------------------------ inpaint
def test_inpaint_fda_against_the_exact_least_squares_fill(nb):
    """inpaint_nans_by_fda (neilpy.py:1171-1216): the CGLS fill equals the exact least-squares solution of the
    reference's own system (<= 1e-6 m) and the reference's LSQR answer within LSQR's own tolerance."""
    rng = np.random.default_rng(4)
    yy, xx = np.meshgrid(np.arange(70.0), np.arange(95.0), indexing='ij')
    Z = O.terrain(xx * 3, yy * 3) + rng.normal(0, 0.1, (70, 95))
    A = Z.copy()
    A[rng.random(A.shape) < 0.2] = np.nan
    A[10:19, 20:33] = np.nan          # a block
    A[0, 0] = np.nan                  # a corner (no equation of its own)
    A[69, 5:9] = np.nan               # on the last row
    A[30:34, 0] = np.nan              # on the first column
    exact = O.fda_fill_exact(A)
    ref = O.inpaint_nans_by_fda(A)
    got, info = nb.inpaint_nans_by_fda(A, return_info=True)
    assert info['converged'] and not np.isnan(got).any()
    assert np.array_equal(got[~np.isnan(A)], A[~np.isnan(A)])
    assert np.abs(got - exact).max() <= 1e-6, float(np.abs(got - exact).max())
    assert np.abs(got - ref).max() <= 1e-3
    # float32 grid, in place, no NaN, all NaN
    got32 = nb.inpaint_nans_by_fda(A.astype(np.float32))
    assert got32.dtype == np.float32 and np.abs(got32 - exact).max() <= 5e-5
    B = A.copy()
    assert nb.inpaint_nans_by_fda(B, inplace=True) is None and np.abs(B - exact).max() <= 1e-6
    assert np.array_equal(nb.inpaint_nans_by_fda(Z), Z)
    assert np.array_equal(nb.inpaint_nans_by_fda(np.full((9, 7), np.nan)), np.zeros((9, 7)))


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_inpaint_against_exact_harmonic_fill(nb, dtype):
    rng = np.random.default_rng(11)
    A = surface(180, 230, 12, np.float64)
    A[rng.random(A.shape) < 0.3] = np.nan
    A[40:95, 60:130] = np.nan                      # a building-sized hole
    A[:20, :15] = np.nan                           # a hole on the grid corner (natural boundary)
    A = A.astype(dtype)
    exact = O.harmonic_fill_exact(A.astype(np.float64))
    got, info = nb.inpaint_nans_by_springs(A, return_info=True)
    assert got.dtype == dtype and not np.isnan(got).any()
    known = ~np.isnan(A)
    assert np.array_equal(got[known], A[known])
    tol = 1e-6 if dtype == np.float64 else 2e-5      # float32 storage rounds to ~8e-6 m at 100 m
    assert np.abs(got.astype(np.float64) - exact).max() <= tol, info
    ref = O.inpaint_nans_by_springs(A.astype(np.float64))
    assert np.abs(got.astype(np.float64) - ref).max() <= 2e-2


def test_inpaint_edge_cases(nb):
    A = surface(33, 47, 1)
    assert np.array_equal(nb.inpaint_nans_by_springs(A), A)                      # no NaN: unchanged
    assert np.array_equal(nb.inpaint_nans_by_springs(np.full((6, 9), np.nan)), np.zeros((6, 9)))   # all NaN: zeros
    B = A.copy(); B[5:9, 7:30] = np.nan
    keep = B.copy()
    out = nb.inpaint_nans_by_springs(B)
    assert eq_nan(B, keep) and not np.isnan(out).any()                           # input not mutated
    # (dot products are reduced with atomics: two solves agree to rounding, not bit for bit)
    assert nb.inpaint_nans_by_springs(B, inplace=True) is None and np.allclose(B, out, rtol=0, atol=1e-8)
    one = np.array([[1.0, np.nan, 3.0]])
    assert np.allclose(nb.inpaint_nans_by_springs(one), [[1.0, 2.0, 3.0]], atol=1e-9)


def test_inpaint_reports_a_non_finite_residual(nb):
    """an inf elevation next to a hole must not read as a converged solve (the residual max keeps
    NaN / inf), and the caller is told."""
    from neilpy_b200.api import InpaintWarning
    A = surface(40, 52, 3)
    A[10:14, 20:30] = np.nan
    A[9, 22] = np.inf
    with pytest.warns(InpaintWarning, match='non-finite'):
        _, info = nb.inpaint_nans_by_springs(A, return_info=True)
    assert not info['converged'] and not np.isfinite(info['residual'])


def test_inpaint_isprs_sparse_sample(nb):
    x, y, z, _ = load_isprs('samp53')                 # 82 % empty cells, LSQR needs 333 iterations
    I, _ = O.create_dem(x, y, z, 1, 'min')
    exact = O.harmonic_fill_exact(I)
    got, info = nb.inpaint_nans_by_springs(I, return_info=True)
    assert np.abs(got - exact).max() <= 1e-6, info
    assert np.abs(got - O.inpaint_nans_by_springs(I)).max() <= 2e-2


# ------------------------------------------------------------------ slope, spline, classify
def run_tail(nb, Zpro, x, y, z, t6, cs, dtype, et=.5, es=1.25):
    """slope -> prefilter x2 -> classify through the C ABI on a given provisional surface"""
    import ctypes as C
    import torch
    from neilpy_b200 import _lib
    from neilpy_b200.affine import Affine
    from neilpy_b200.api import _ptr, _stream, _code, _factors, _inverse6
    lib = _lib.load()
    dev = torch.device('cuda')
    Zt = torch.as_tensor(np.ascontiguousarray(Zpro.astype(dtype))).cuda()
    ny, nx = Zt.shape
    code = _code(Zt.dtype)
    S = torch.empty_like(Zt)
    _lib.check(lib.smrf_slope(_ptr(Zt), _ptr(S), ny, nx, code, float(cs), _stream()), 'slope')
    Sraw = S.clone()
    ws = torch.empty(lib.smrf_spline_workspace_bytes(ny, nx), dtype=torch.uint8, device=dev)
    cz = torch.empty_like(Zt)
    rf, cf = _factors(ny, dev), _factors(nx, dev)
    _lib.check(lib.smrf_spline_prefilter(_ptr(Zt), _ptr(cz), 1, 0, ny, nx, code, _ptr(rf), _ptr(cf), _ptr(ws), ws.numel(), _stream()), 'pre')
    _lib.check(lib.smrf_spline_prefilter(_ptr(S), _ptr(S), 1, 0, ny, nx, code, _ptr(rf), _ptr(cf), _ptr(ws), ws.numel(), _stream()), 'pre')
    xt, yt, zt = [torch.as_tensor(v).cuda() for v in (x, y, z)]
    n = xt.numel()
    obj = torch.empty(n, dtype=torch.uint8, device=dev)
    ev = torch.empty(n, dtype=torch.float64, device=dev)
    sv = torch.empty(n, dtype=torch.float64, device=dev)
    inv6 = _inverse6(Affine(*t6))
    _lib.check(lib.smrf_classify(_ptr(xt), _ptr(yt), _ptr(zt), n, _lib.PTS_SOA_F64, inv6, _ptr(cz), _ptr(S), ny, nx, code,
                                 et, es, _ptr(obj), _ptr(ev), _ptr(sv), None, None, _stream()), 'classify')
    torch.cuda.synchronize()
    return Sraw.cpu().numpy(), cz.cpu().numpy(), ev.cpu().numpy(), sv.cpu().numpy(), obj.cpu().numpy().astype(bool)


@pytest.mark.parametrize('dtype,tol', [(np.float64, 1e-9), (np.float32, 2e-4)])
def test_slope_spline_and_classification_given_the_oracle_surface(nb, dtype, tol):
    x, y, z, _ = load_isprs('samp12')
    st = {}
    O.smrf(x, y, z, 1, 18, .15, .5, 1.25, stages=st)
    Zpro = st['Zpro'].astype(dtype).astype(np.float64)
    ny, nx = Zpro.shape
    # the oracle's tail on the (possibly float32-rounded) surface
    gy, gx = np.gradient(Zpro, 1)
    S = np.sqrt(gy ** 2 + gx ** 2)
    rc, cc = np.arange(0.5, ny + .5), np.arange(0.5, nx + .5)
    f1, f2 = interpolate.RectBivariateSpline(rc, cc, Zpro), interpolate.RectBivariateSpline(rc, cc, S)
    ev0, sv0 = f1.ev(st['r'], st['c']), f2.ev(st['r'], st['c'])
    req0 = .5 + 1.25 * sv0
    obj0 = np.abs(ev0 - z) > req0
    Sg, cz, ev, sv, obj = run_tail(nb, Zpro, x, y, z, st['t'], 1, dtype)
    if dtype == np.float64:
        assert np.array_equal(Sg, S)                                   # same float64 operations in the same order
        assert np.abs(cz - np.asarray(f1.tck[2]).reshape(ny, nx)).max() <= tol
    else:
        assert np.abs(Sg - S).max() <= 1e-6
    assert np.abs(ev - ev0).max() <= tol and np.abs(sv - sv0).max() <= tol
    flips = obj != obj0
    margin = np.abs(np.abs(ev0 - z) - req0)
    assert np.all(margin[flips] <= 3 * tol), (int(flips.sum()), float(margin[flips].max()) if flips.any() else 0)
    if dtype == np.float64:
        assert flips.sum() <= 2


def test_points_outside_the_centre_range_are_clamped_like_bispeu(nb):
    rng = np.random.default_rng(4)
    Z = surface(12, 9, 5)
    t6 = (1.0, 0.0, -0.5, 0.0, -1.0, 11.5)
    x = rng.uniform(-0.5, 8.5, 400); y = rng.uniform(-0.5, 11.5, 400); z = rng.uniform(90, 130, 400)
    r = (11.5 - y); c = (x + 0.5)
    f1 = interpolate.RectBivariateSpline(np.arange(0.5, 12.5), np.arange(0.5, 9.5), Z)
    _, _, ev, _, _ = run_tail(nb, Z, x, y, z, t6, 1, np.float64)
    assert np.abs(ev - f1.ev(r, c)).max() <= 1e-9
