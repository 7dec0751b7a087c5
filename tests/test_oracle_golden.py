"""The oracle against the reference's own pinned results (CPU only).

The only reference-authored known answer for this path is the samp12 print-out of
examples/smrf/The Simple Morphological Filter (SMRF) for Point Cloud Processing.ipynb:902-905.
"""
import hashlib

import numpy as np
import pytest
import scipy.ndimage as ndi

from conftest import load_isprs
from oracle import smrf_oracle as O


def sha(a):
    return hashlib.sha256(np.packbits(np.asarray(a, dtype=bool))).hexdigest()[:12]


def test_samp12_reproduces_notebook_printout(expected):
    from sklearn.metrics import cohen_kappa_score
    x, y, z, g = load_isprs('samp12')
    p = expected['params']
    Z, t, oc, op = O.smrf(x, y, z, p['cellsize'], p['windows'], p['slope_threshold'], p['elevation_threshold'],
                          p['elevation_scaler'])
    # the notebook's own (idiosyncratic) definitions, cell at :1711-1720
    total = 1 - np.sum(op == g) / len(g)
    t1 = np.sum((g == 0) & (op == 1)) / np.sum(g == 1)
    t2 = np.sum((g == 1) & (op == 0)) / np.sum(g == 0)
    kappa = cohen_kappa_score(g, op)
    nb = expected['notebook_samp12']
    # the notebook prints 12 significant digits
    assert '%.12g' % (100 * t1) == '%.12g' % nb['type_I']
    assert '%.12g' % (100 * t2) == '%.12g' % nb['type_II']
    assert '%.12g' % (100 * total) == '%.12g' % nb['total']
    assert '%.12g' % (100 * kappa) == '%.12g' % nb['kappa']
    e = expected['samples']['samp12']
    assert list(Z.shape) == e['shape']
    assert list(t.coeffs) == e['t']
    assert sha(op) == e['sha_point_mask'] and sha(oc) == e['sha_cell_mask']


@pytest.mark.parametrize('name', ['samp21', 'samp24', 'samp31', 'samp54'])
def test_isprs_regression_pins(expected, name):
    x, y, z, g = load_isprs(name)
    p = expected['params']
    st = {}
    Z, t, oc, op = O.smrf(x, y, z, p['cellsize'], p['windows'], p['slope_threshold'], p['elevation_threshold'],
                          p['elevation_scaler'], stages=st)
    e = expected['samples'][name]
    assert list(Z.shape) == e['shape'] and list(t.coeffs) == e['t']
    assert int(np.isnan(st['Zmin_binned']).sum()) == e['empty_cells']
    assert sha(np.isnan(st['Zmin_binned'])) == e['sha_empty']
    assert float(np.nansum(st['Zmin_binned'])) == e['zmin_nansum']
    assert int(st['low_outliers'].sum()) == e['low_outlier_cells']
    assert sha(st['progressive_cells']) == e['sha_progressive_cells']
    assert int(oc.sum()) == e['object_cells'] and int(op.sum()) == e['object_points']
    assert sha(oc) == e['sha_cell_mask'] and sha(op) == e['sha_point_mask']


def test_disk_is_skimage_definition():
    assert O.disk(1).tolist() == [[0, 1, 0], [1, 1, 1], [0, 1, 0]]
    assert [int(O.disk(w).sum()) for w in (1, 2, 3, 4, 5, 18, 36)] == [5, 13, 29, 49, 81, 1009, 4053]
    d = O.disk(18)
    assert np.array_equal(d, d.T) and np.array_equal(d, d[::-1]) and d[0, 18] == 1 and d[0, 17] == 0


def ignore_oob_morph(img, w, op):
    """The border rule the CUDA kernels implement: samples outside the image do not exist."""
    ny, nx = img.shape
    out = np.empty_like(img)
    fp = O.disk(w).astype(bool)
    for y in range(ny):
        for x in range(nx):
            y0, y1 = max(0, y - w), min(ny, y + w + 1)
            x0, x1 = max(0, x - w), min(nx, x + w + 1)
            sub = img[y0:y1, x0:x1]
            m = fp[y0 - y + w:y1 - y + w, x0 - x + w:x1 - x + w]
            out[y, x] = op(sub[m])
    return out


@pytest.mark.parametrize('w', [1, 2, 3, 5, 8])
def test_reflect_border_equals_ignore_out_of_bounds(w):
    """for a disk footprint scipy's mode='reflect' == ignoring out-of-image samples
    (grid dimensions >= w)."""
    rng = np.random.default_rng(w)
    img = rng.normal(size=(23, 31))
    er = ndi.grey_erosion(img, footprint=O.disk(w))
    assert np.array_equal(er, ignore_oob_morph(img, w, np.min))
    op = ndi.grey_dilation(er, footprint=O.disk(w))
    assert np.array_equal(op, ignore_oob_morph(er, w, np.max))


def test_progressive_chain_is_not_reducible():
    """each window must open the previous window's output."""
    rng = np.random.default_rng(0)
    Z = np.cumsum(np.cumsum(rng.normal(size=(64, 64)), 0), 1) * 0.05
    chained = Z.copy()
    for w in range(1, 7):
        chained = O.opening(chained, O.disk(w))
    assert np.abs(chained - O.opening(Z, O.disk(6))).max() > 0


def test_inpaint_properties():
    rng = np.random.default_rng(1)
    A = rng.normal(size=(20, 17)) + 50
    assert np.array_equal(O.inpaint_nans_by_springs(A), A)          # no NaN -> unchanged
    B = A.copy()
    B[rng.random(B.shape) < 0.4] = np.nan
    filled = O.inpaint_nans_by_springs(B)
    exact = O.harmonic_fill_exact(B)
    assert not np.isnan(filled).any()
    assert np.array_equal(filled[~np.isnan(B)], B[~np.isnan(B)])
    assert np.abs(filled - exact).max() < 1e-2                      # LSQR is inexact (F7)
    # the exact fill is discrete-harmonic on the NaN cells
    pad = np.pad(exact, 1, mode='edge')
    lap = (pad[:-2, 1:-1] + pad[2:, 1:-1] + pad[1:-1, :-2] + pad[1:-1, 2:]) - 4 * exact
    assert np.abs(lap[np.isnan(B)]).max() < 1e-9
    assert np.array_equal(O.inpaint_nans_by_springs(np.full((5, 6), np.nan)), np.zeros((5, 6)))


def test_create_dem_geometry_and_edge_rule():
    # x on a vertical edge goes right, y on a horizontal edge goes down (south)
    x = np.array([0.0, 0.5, 3.0]); y = np.array([0.0, 0.5, 2.0]); z = np.array([1.0, 2.0, 3.0])
    I, t = O.create_dem(x, y, z, cellsize=1, bin_type='min')
    assert I.shape == (3, 4) and t.coeffs == (1.0, 0.0, -0.5, 0.0, -1.0, 2.5)
    assert I[2, 0] == 1.0 and I[2, 1] == 2.0 and I[0, 3] == 3.0     # (0.5,0.5) sits on two edges -> col 1, row 2
    assert np.isnan(I).sum() == 9
    with pytest.raises(ValueError):
        O.create_dem(x, y, z, bin_type='median')


# ---------------------------------------------------------------------------------------------
# The reference's own source, executed (tests/golden/make_reference_exec_golden.py): unique_rows,
# inpaint_nans_by_springs, create_dem, progressive_filter and smrf cut out of neilpy/neilpy.py and
# run unmodified, with only rasterio's from_origin and skimage's disk / opening supplied.  The
# oracle's restatement must reproduce every output bit for bit.
def _reference_exec():
    import importlib.util
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    spec = importlib.util.spec_from_file_location('make_reference_exec_golden', os.path.join(here, 'make_reference_exec_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(here, 'reference_exec.json')) as f:
        return mod, json.load(f)


@pytest.mark.parametrize('case', ['samp11', 'samp12', 'synth_w6', 'synth_cs2_fill'])
def test_oracle_equals_the_executed_reference_source(case):
    mod, want = _reference_exec()
    x, y, z, kw = mod.cases()[case]
    got = mod.run(mod.oracle_functions(), x, y, z, kw)
    assert set(got) == set(want[case])
    for k in sorted(got):
        assert got[k] == want[case][k], k
