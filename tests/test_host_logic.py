"""Host-side logic of neilpy_b200 (no GPU, no library call)."""
import numpy as np
import pytest
from scipy import interpolate

from neilpy_b200 import spline as S
from neilpy_b200.affine import Affine
from neilpy_b200 import api
from oracle import smrf_oracle as O


def test_affine_matches_oracle_restatement():
    rng = np.random.default_rng(0)
    for cs in (1, 0.5, 0.25, 0.1, 5, 1.0 / 3):
        west, north = float(rng.uniform(-1e6, 1e6)), float(rng.uniform(-1e7, 1e7))
        a, o = Affine.from_origin(west, north, cs, cs), O.Affine6.from_origin(west, north, cs, cs)
        assert a.coeffs == o.coeffs and (~a).coeffs == (~o).coeffs
        x, y = rng.uniform(-1e6, 1e6, 100), rng.uniform(-1e7, 1e7, 100)
        ca, ra = ~a * (x, y)
        co, ro = ~o * (x, y)
        assert np.array_equal(ca, co) and np.array_equal(ra, ro)
        assert a[0] == cs and a[4] == -cs and a[8] == 1.0 and len(a) == 9
        xy = a * (3, 4)
        assert xy == (west + 3 * cs, north - 4 * cs)


@pytest.mark.parametrize('cs', [1, 0.5, 0.1, 2.5])
def test_edges_match_reference_geometry(cs):
    rng = np.random.default_rng(3)
    x, y, z = rng.uniform(100, 163.7, 500), rng.uniform(-20, 31.3, 500), rng.normal(size=500)
    I, t = O.create_dem(x, y, z, cellsize=cs, bin_type='min')
    xe, ye = api._edges(x.min(), x.max(), y.min(), y.max(), cs)
    assert (len(ye) - 1, len(xe) - 1) == I.shape
    assert (float(xe[0]), float(ye[0])) == (t.coeffs[2], t.coeffs[5])


def test_windows_and_thresholds_follow_numpy_order():
    w = api._windows(18)
    assert w.tolist() == list(range(1, 19))
    thr = .15 * (w * 1)
    assert thr[2] == 0.44999999999999996        # 0.15*3 in float64, not 0.45
    assert api._windows(np.array([1, 3, 9])).tolist() == [1, 3, 9]
    with pytest.raises(ValueError):
        api._windows(np.array([[1, 2]]))


def chord_half(w, dy):
    h = 0
    while (h + 1) * (h + 1) <= w * w - dy * dy:
        h += 1
    return h


@pytest.mark.parametrize('w', [1, 2, 3, 7, 18, 36, 72])
def test_chord_table_is_the_disk(w):
    """The kernels use h(dy) = floor(sqrt(w^2 - dy^2)); row dy of disk(w) must be |dx| <= h(dy)."""
    d = O.disk(w)
    for dy in range(-w, w + 1):
        h = chord_half(w, dy)
        row = np.zeros(2 * w + 1, dtype=np.uint8)
        row[w - h:w + h + 1] = 1
        assert np.array_equal(d[dy + w], row)
    assert chord_half(w, w) == 0 and all(chord_half(w, dy) >= 1 for dy in range(w))


@pytest.mark.parametrize('shape', [(4, 4), (5, 9), (4, 30), (37, 41), (120, 77)])
def test_spline_factors_and_evaluator_match_fitpack(shape):
    rng = np.random.default_rng(sum(shape))
    ny, nx = shape
    Z = rng.normal(size=shape) * 10 + 100
    f = interpolate.RectBivariateSpline(np.arange(0.5, ny + .5), np.arange(0.5, nx + .5), Z)
    tx, ty, c = f.tck
    assert np.array_equal(tx, [S.knot(j, ny) for j in range(ny + 4)])
    assert np.array_equal(ty, [S.knot(j, nx) for j in range(nx + 4)])
    coef = S.prefilter(Z)
    assert np.abs(coef - np.asarray(c).reshape(shape)).max() < 1e-10
    r, cq = rng.uniform(-2, ny + 2, 300), rng.uniform(-2, nx + 2, 300)     # outside -> clamped, as bispeu
    r[:6] = [0.5, ny - 0.5, 2.5, ny - 2.5, 0.0, ny]
    cq[:6] = [0.5, nx - 0.5, nx - 2.5, 2.5, nx, 0.0]
    assert np.abs(S.evaluate(coef, r, cq) - f.ev(r, cq)).max() < 1e-10


def test_spline_needs_four_samples():
    with pytest.raises(ValueError):
        S.notaknot_factors(3)


def test_spline_factor_decay_bounds_the_chunk_warmup():
    """The CUDA prefilter restarts every 256-cell chunk 40 cells early from a zero state;
    the recurrences must contract by ~0.268 per cell for that to be exact to rounding."""
    fac = S.notaknot_factors(4096)
    l1, l2, dinv, u1, u2 = fac
    assert np.all(np.abs(l1[8:-8]) < 0.27) and np.all(l2[8:-8] == 0)
    assert np.all(np.abs(u1[8:-8] * dinv[8:-8]) < 0.27) and np.all(u2[8:-8] == 0)
    assert 0.27 ** 40 < 1e-22


def test_synthetic_cloud_is_deterministic_and_f32_exact():
    x, y, z, lab = O.synth_cloud(20000, 300.0, 200.0, seed=0)
    x2, y2, z2, _ = O.synth_cloud(20000, 300.0, 200.0, seed=0)
    assert np.array_equal(x, x2) and np.array_equal(z, z2)
    assert np.array_equal(x.astype(np.float32).astype(np.float64), x)
    assert np.array_equal(z.astype(np.float32).astype(np.float64), z)
    assert 0.05 < lab.mean() < 0.6


def test_threaded_host_copy(monkeypatch):
    """The staging -> result copy of api._to_host: exact for every size and dtype, and it picks its
    own thread count (torchrun exports OMP_NUM_THREADS=1, which would serialise torch's copy)."""
    from neilpy_b200 import api
    rng = np.random.default_rng(0)
    for dt in (np.float32, np.float64, np.uint8, np.bool_):
        for n in (0, 1, 4097, (4 << 20) + 3):
            src = rng.integers(0, 2, n).astype(dt)
            dst = np.empty(n, dt)
            api._threaded_copy(dst, src)
            assert np.array_equal(dst, src)
    monkeypatch.setenv('LOCAL_WORLD_SIZE', '1')
    one = api._copy_threads()
    monkeypatch.setenv('LOCAL_WORLD_SIZE', '64')
    assert 1 <= api._copy_threads() <= one <= 4


def test_inpaint_convergence_is_reported():
    """a solve that ends above its tolerance, or with a non-finite residual, warns and says so in `info`."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('error')                      # the converged cases must not warn
        assert api._converged({'iterations': 5, 'residual': 1e-9, 'unknown': 10}, 1e-6)['converged']
        assert api._converged({'iterations': 0, 'residual': float('nan'), 'unknown': 0}, 1e-6)['converged']   # nothing to solve
    with pytest.warns(api.InpaintWarning, match='tolerance'):
        assert not api._converged({'iterations': 4000, 'residual': 1e-3, 'unknown': 10}, 1e-6)['converged']
    for bad in (float('inf'), float('nan')):
        with pytest.warns(api.InpaintWarning, match='non-finite'):
            assert not api._converged({'iterations': 1, 'residual': bad, 'unknown': 10}, 1e-6)['converged']
