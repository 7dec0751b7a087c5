"""CPU oracle for the SMRF hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, in numpy/scipy, the four reference functions on the SMRF
path of thomaspingel/neilpy so that the CUDA path can be checked against them:

    create_dem                 neilpy/neilpy.py:1110-1166
    unique_rows                neilpy/neilpy.py:1221-1224
    inpaint_nans_by_springs    neilpy/neilpy.py:1227-1271
    inpaint_nans_by_fda        neilpy/neilpy.py:1171-1216
    progressive_filter         neilpy/neilpy.py:1659-1680
    smrf                       neilpy/neilpy.py:1685-1808

The reference itself cannot be imported in this image (matplotlib, rasterio,
affine, skimage, ... are absent -- SURVEY.md F1), so three third-party pieces are
restated from their published definitions:

    skimage.morphology.disk(w)          -> `disk`            (x^2 + y^2 <= w^2)
    skimage.morphology.opening(img,fp)  -> grey_erosion then grey_dilation of
                                           scipy.ndimage with the same footprint
                                           (skimage >=0.19 dispatches to exactly these)
    rasterio.transform.from_origin, ~t, t*(x,y)
                                        -> `Affine6` (the 6-float formulas of the
                                           `affine` package: __invert__, __mul__)

Everything else (pandas groupby-min, scipy.sparse + lsqr, RectBivariateSpline,
np.gradient) is called exactly as the reference calls it.

Parity pins (tests/test_oracle_golden.py):
 1. `smrf` on sample_data/samp12.txt with the notebook parameters reproduces the
    reference notebook's printed output (Type I 2.00566304861 %, Type II
    4.12498595032 %, total 3.09100328095 %, kappa 93.8109576375 %) digit for digit.
 2. The reference's OWN source for the five functions above, cut out of
    neilpy/neilpy.py and executed unmodified in this container with only the three
    third-party names above supplied (tests/golden/make_reference_exec_golden.py),
    returns bit for bit what this restatement returns -- DTM, transform, cell mask,
    point mask, extras, create_dem min / max+inpaint, inpaint, progressive_filter mask
    and when_dropped -- on samp11, samp12 and two synthetic clouds (digests in
    tests/golden/reference_exec.json).
What stays unpinned is only the three third-party restatements themselves (skimage's
disk / opening, rasterio's from_origin): the reference's tests hold no vector for
them and the packages are absent; they follow the packages' published definitions,
and pin 1 passes through all three.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (neilpy_b200) never does.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.ndimage as ndi
from scipy import interpolate, sparse
from scipy.sparse import linalg as splinalg


# --------------------------------------------------------------------------
# third-party restatements
# --------------------------------------------------------------------------
class Affine6:
    """The six coefficients (a, b, c, d, e, f) of affine.Affine, with the
    arithmetic of its __invert__ and __mul__ (neilpy.py:1141-1142, :1772)."""

    def __init__(self, a, b, c, d, e, f):
        self.coeffs = (float(a), float(b), float(c), float(d), float(e), float(f))

    @classmethod
    def from_origin(cls, west, north, xsize, ysize):
        # rasterio.transform.from_origin = translation(west, north) * scale(xsize, -ysize)
        return cls(xsize, 0.0, west, 0.0, -ysize, north)

    def __getitem__(self, i):
        return (self.coeffs + (0.0, 0.0, 1.0))[i]

    def __invert__(self):
        a, b, c, d, e, f = self.coeffs
        idet = 1.0 / (a * e - b * d)
        ra = e * idet
        rb = -b * idet
        rd = -d * idet
        re = a * idet
        return Affine6(ra, rb, -c * ra - f * rb, rd, re, -c * rd - f * re)

    def __mul__(self, other):
        sa, sb, sc, sd, se, sf = self.coeffs
        vx, vy = other
        return (vx * sa + vy * sb + sc, vx * sd + vy * se + sf)

    def __repr__(self):
        return "Affine6(%r, %r, %r,\n        %r, %r, %r)" % self.coeffs


def disk(radius):
    """skimage.morphology.disk: (2r+1)^2 uint8, 1 where dx^2+dy^2 <= r^2."""
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=np.uint8)


def opening(image, footprint):
    """skimage.morphology.opening for an ndarray footprint."""
    eroded = ndi.grey_erosion(image, footprint=footprint)
    return ndi.grey_dilation(eroded, footprint=footprint)


# --------------------------------------------------------------------------
# neilpy.py:1110-1166
# --------------------------------------------------------------------------
def create_dem(x, y, z, cellsize=1, bin_type='max', inpaint=False, edges=None):
    floor2 = lambda x, v: v * np.floor(x / v)
    ceil2 = lambda x, v: v * np.ceil(x / v)

    if edges is None:
        xedges = np.arange(floor2(np.min(x), cellsize) - .5 * cellsize,
                           ceil2(np.max(x), cellsize) + 1.5 * cellsize, cellsize)
        yedges = np.arange(ceil2(np.max(y), cellsize) + .5 * cellsize,
                           floor2(np.min(y), cellsize) - 1.5 * cellsize, -cellsize)
    else:
        xedges = edges[0]
        yedges = edges[1]
        out_of_range = (x < xedges[0]) | (x > xedges[-1]) | (y > yedges[0]) | (y < yedges[-1])
        x = x[~out_of_range]
        y = y[~out_of_range]
        z = z[~out_of_range]
        cellsize = np.abs(xedges[1] - xedges[0])

    nx, ny = len(xedges) - 1, len(yedges) - 1

    I = np.empty(nx * ny)
    I[:] = np.nan

    t = Affine6.from_origin(xedges[0], yedges[0], cellsize, cellsize)
    c, r = ~t * (x, y)
    c, r = np.floor(c).astype(np.int64), np.floor(r).astype(np.int64)

    mx = pd.DataFrame({'i': np.ravel_multi_index((r, c), (ny, nx)), 'z': z}).groupby('i')
    del c, r
    if bin_type == 'max':
        mx = mx.max()
    elif bin_type == 'min':
        mx = mx.min()
    else:
        raise ValueError('This type not supported.')

    I.flat[mx.index.values] = mx.values
    I = I.reshape((ny, nx))

    if inpaint == True:
        I = inpaint_nans_by_springs(I)

    return I, t


# --------------------------------------------------------------------------
# neilpy.py:1221-1271
# --------------------------------------------------------------------------
def unique_rows(a):
    a = np.ascontiguousarray(a)
    unique_a = np.unique(a.view([('', a.dtype)] * a.shape[1]))
    return unique_a.view(a.dtype).reshape((unique_a.shape[0], a.shape[1]))


def inpaint_nans_by_springs(A, inplace=False, neighbors=4, return_info=False):
    m, n = np.shape(A)
    nanmat = np.isnan(A)

    nan_list = np.flatnonzero(nanmat)
    known_list = np.flatnonzero(~nanmat)

    r, c = np.unravel_index(nan_list, (m, n))

    num_neighbors = neighbors
    neighbors = np.array([[0, 1], [0, -1], [-1, 0], [1, 0]])
    neighbors = np.vstack([np.vstack((r + i[0], c + i[1])).T for i in neighbors])
    del r, c

    springs = np.tile(nan_list, num_neighbors)
    good_rows = (np.all(neighbors >= 0, 1)) & (neighbors[:, 0] < m) & (neighbors[:, 1] < n)

    neighbors = np.ravel_multi_index((neighbors[good_rows, 0], neighbors[good_rows, 1]), (m, n))
    springs = springs[good_rows]

    springs = np.vstack((springs, neighbors)).T
    del neighbors, good_rows

    springs = np.sort(springs, axis=1)
    springs = unique_rows(springs)

    n_springs = np.shape(springs)[0]

    i = np.tile(np.arange(n_springs), 2)
    springs = springs.T.ravel()
    data = np.hstack((np.ones(n_springs, dtype=np.int8), -1 * np.ones(n_springs, dtype=np.int8)))
    springs = sparse.coo_matrix((data, (i, springs)), (n_springs, m * n), dtype=np.int8).tocsr()
    del i, data

    rhs = -springs[:, known_list] * A[np.unravel_index(known_list, (m, n))]
    sol = splinalg.lsqr(springs[:, nan_list], rhs)
    results = sol[0]

    if inplace:
        A[np.unravel_index(nan_list, (m, n))] = results
        return None
    B = A.copy()
    B[np.unravel_index(nan_list, (m, n))] = results
    if return_info:
        return B, {'lsqr_iterations': int(sol[2]), 'lsqr_istop': int(sol[1])}
    return B


def harmonic_fill_exact(A):
    """Exact (sparse-direct) solution of the spring system that
    inpaint_nans_by_springs hands to LSQR: for every NaN cell i,
    deg(i)*u_i - sum_{NaN nbrs} u_j = sum_{known nbrs} a_k, deg = number of
    in-grid 4-neighbours.  Used to state how far LSQR (atol=btol=1e-6) and the
    CUDA solver each sit from the true minimiser."""
    A = np.asarray(A, dtype=np.float64)
    m, n = A.shape
    nan = np.isnan(A)
    if not nan.any():
        return A.copy()
    idx = -np.ones(m * n, dtype=np.int64)
    nan_list = np.flatnonzero(nan)
    idx[nan_list] = np.arange(nan_list.size)
    r, c = np.unravel_index(nan_list, (m, n))
    rows, cols, vals = [], [], []
    deg = np.zeros(nan_list.size)
    rhs = np.zeros(nan_list.size)
    flatA = A.ravel()
    for dr, dc in ((0, 1), (0, -1), (-1, 0), (1, 0)):
        rr, cc = r + dr, c + dc
        ok = (rr >= 0) & (rr < m) & (cc >= 0) & (cc < n)
        deg += ok
        nb = np.where(ok, rr * n + cc, 0)
        nb_nan = ok & nan.ravel()[nb]
        nb_known = ok & ~nan.ravel()[nb]
        rows.append(np.flatnonzero(nb_nan))
        cols.append(idx[nb[nb_nan]])
        vals.append(-np.ones(int(nb_nan.sum())))
        rhs += np.where(nb_known, flatA[nb], 0.0)
    rows.append(np.arange(nan_list.size))
    cols.append(np.arange(nan_list.size))
    vals.append(deg)
    L = sparse.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(nan_list.size, nan_list.size))
    # components with no known neighbour anywhere are singular; LSQR returns the
    # minimum-norm solution (0 for an all-NaN grid).  Regularise those only.
    ncomp, lab = sparse.csgraph.connected_components(L, directed=False)
    has_known = np.zeros(ncomp, dtype=bool)
    np.logical_or.at(has_known, lab, rhs != 0)
    touches = np.zeros(ncomp, dtype=bool)
    known_nb = np.zeros(nan_list.size, dtype=bool)
    for dr, dc in ((0, 1), (0, -1), (-1, 0), (1, 0)):
        rr, cc = r + dr, c + dc
        ok = (rr >= 0) & (rr < m) & (cc >= 0) & (cc < n)
        nb = np.where(ok, rr * n + cc, 0)
        known_nb |= ok & ~nan.ravel()[nb]
    np.logical_or.at(touches, lab, known_nb)
    free = ~touches[lab]
    if free.any():
        L = L + sparse.diags(free.astype(np.float64))
    u = splinalg.spsolve(L.tocsc(), rhs)
    B = A.copy()
    B.ravel()[nan_list] = u
    return B


# --------------------------------------------------------------------------
# neilpy.py:1171-1216  (the step after the path)
# --------------------------------------------------------------------------
def _fda_system(A, fast=True):
    """The sparse system of inpaint_nans_by_fda exactly as the reference assembles it: returns (a, rhs_k, nan_list)."""
    m, n = np.shape(A)
    nanmat = np.isnan(A)
    nan_list = np.flatnonzero(nanmat)
    known_list = np.flatnonzero(~nanmat)
    index = np.arange(m * n, dtype=np.int64).reshape((m, n))
    i = np.hstack((np.tile(index[1:-1, :].ravel(), 3),
                   np.tile(index[:, 1:-1].ravel(), 3)))
    j = np.hstack((index[0:-2, :].ravel(),
                   index[2:, :].ravel(),
                   index[1:-1, :].ravel(),
                   index[:, 0:-2].ravel(),
                   index[:, 2:].ravel(),
                   index[:, 1:-1].ravel()))
    data = np.hstack((np.ones(2 * n * (m - 2), dtype=np.int64),
                      -2 * np.ones(n * (m - 2), dtype=np.int64),
                      np.ones(2 * m * (n - 2), dtype=np.int64),
                      -2 * np.ones(m * (n - 2), dtype=np.int64)))
    if fast == True:  # noqa: E712
        goodrows = np.isin(i, index[ndi.binary_dilation(nanmat)])      # np.in1d in the reference (removed in numpy 2)
        i = i[goodrows]
        j = j[goodrows]
        data = data[goodrows]
    fda = sparse.coo_matrix((data, (i, j)), (m * n, m * n), dtype=np.int8).tocsr()
    rhs = -fda[:, known_list] * A[np.unravel_index(known_list, (m, n))]
    k = fda[:, np.unique(nan_list)]
    k = k.nonzero()[0]
    a = fda[k][:, nan_list]
    return a, rhs[k], nan_list


def inpaint_nans_by_fda(A, fast=True, inplace=False):
    m, n = np.shape(A)
    a, rhs, nan_list = _fda_system(A, fast)
    results = splinalg.lsqr(a, rhs)[0]
    if inplace:
        A[np.unravel_index(nan_list, (m, n))] = results
    else:
        B = A.copy()
        B[np.unravel_index(nan_list, (m, n))] = results
        return B


def fda_fill_exact(A):
    """The exact least-squares fill the reference's LSQR approximates (tight LSQR on the same system)."""
    m, n = np.shape(A)
    a, rhs, nan_list = _fda_system(A, True)
    B = A.copy()
    if len(nan_list):
        B[np.unravel_index(nan_list, (m, n))] = splinalg.lsqr(a.astype(np.float64), rhs, atol=1e-15, btol=1e-15,
                                                              conlim=1e16, iter_lim=200000)[0]
    return B


# --------------------------------------------------------------------------
# neilpy.py:1659-1680
# --------------------------------------------------------------------------
def progressive_filter(Z, windows, cellsize=1, slope_threshold=.15, return_when_dropped=False,
                       stage_hook=None):
    last_surface = Z.copy()
    elevation_thresholds = slope_threshold * (windows * cellsize)
    is_object_cell = np.zeros(np.shape(Z), dtype=bool)
    if return_when_dropped:
        when_dropped = np.zeros(np.shape(Z), dtype=np.uint8)
    for i, window in enumerate(windows):
        elevation_threshold = elevation_thresholds[i]
        # neilpy.py:1667-1669 builds a 3x3 square for window==1 but :1670 then
        # ignores it and calls disk(window): radius 1 is the 5-pixel cross.
        this_surface = opening(last_surface, disk(window))
        new_obj = last_surface - this_surface > elevation_threshold
        is_object_cell = (is_object_cell) | (new_obj)
        if return_when_dropped:
            when_dropped[new_obj] = i
        if stage_hook is not None:
            stage_hook(i, window, this_surface, new_obj)
        if i < len(windows) and len(windows) > 1:
            last_surface = this_surface.copy()
    if return_when_dropped:
        return is_object_cell, when_dropped
    else:
        return is_object_cell


# --------------------------------------------------------------------------
# neilpy.py:1685-1808
# --------------------------------------------------------------------------
def smrf(x, y, z, cellsize=1, windows=5, slope_threshold=.15, elevation_threshold=.5,
         elevation_scaler=1.25, low_filter_slope=5, low_outlier_fill=False,
         return_extras=False, stages=None):
    """`stages`, if a dict, receives every intermediate of the run (the
    per-stage golden vectors the CUDA parity tests consume)."""
    if np.isscalar(windows):
        windows = np.arange(windows) + 1

    Zmin, t = create_dem(x, y, z, cellsize=cellsize, bin_type='min')
    is_empty_cell = np.isnan(Zmin)
    if stages is not None:
        stages['Zmin_binned'] = Zmin.copy()
    Zmin = inpaint_nans_by_springs(Zmin)
    if stages is not None:
        stages['Zmin_inpainted'] = Zmin.copy()
    low_outliers = progressive_filter(-Zmin, np.array([1]), cellsize, slope_threshold=low_filter_slope)

    if low_outlier_fill:
        Zmin[low_outliers] = np.nan
        Zmin = inpaint_nans_by_springs(Zmin)
    if stages is not None:
        stages['low_outliers'] = low_outliers.copy()
        stages['Zmin_filtered'] = Zmin.copy()

    if return_extras:
        object_cells, drop_raster = progressive_filter(Zmin, windows, cellsize, slope_threshold,
                                                       return_when_dropped=True)
    else:
        object_cells = progressive_filter(Zmin, windows, cellsize, slope_threshold)
    if stages is not None:
        stages['progressive_cells'] = object_cells.copy()

    Zpro = Zmin
    del Zmin
    object_cells = is_empty_cell | low_outliers | object_cells
    Zpro[object_cells] = np.nan
    if stages is not None:
        stages['Zpro_punched'] = Zpro.copy()
    Zpro = inpaint_nans_by_springs(Zpro)

    col_centers = np.arange(0.5, Zpro.shape[1] + .5)
    row_centers = np.arange(0.5, Zpro.shape[0] + .5)

    c, r = ~t * (x, y)
    f1 = interpolate.RectBivariateSpline(row_centers, col_centers, Zpro)
    elevation_values = f1.ev(r, c)

    if return_extras:
        when_dropped = drop_raster[np.round(r).astype(int), np.round(c).astype(int)]

    gy, gx = np.gradient(Zpro, cellsize)
    S = np.sqrt(gy ** 2 + gx ** 2)
    del gy, gx
    f2 = interpolate.RectBivariateSpline(row_centers, col_centers, S)
    slope_values = f2.ev(r, c)

    required_value = elevation_threshold + (elevation_scaler * slope_values)
    is_object_point = np.abs(elevation_values - z) > required_value

    if stages is not None:
        stages.update(Zpro=Zpro.copy(), object_cells=object_cells.copy(), S=S.copy(),
                      r=np.asarray(r), c=np.asarray(c),
                      elevation_values=np.asarray(elevation_values),
                      slope_values=np.asarray(slope_values),
                      is_object_point=np.asarray(is_object_point), t=t.coeffs)
    del S

    if return_extras == True:
        extras = {}
        extras['above_ground_height'] = z - elevation_values
        extras['drop_raster'] = drop_raster
        extras['when_dropped'] = when_dropped

    if return_extras == False:
        return Zpro, t, object_cells, is_object_point
    else:
        return Zpro, t, object_cells, is_object_point, extras


# --------------------------------------------------------------------------
# deterministic synthetic inputs live in neilpy_b200/synth.py so that
# bench.py's GPU arm can build its workload without importing the oracle; re-exported here
# for the tests.
# --------------------------------------------------------------------------
from neilpy_b200.synth import synth_cloud, synth_dem, terrain  # noqa: E402,F401
