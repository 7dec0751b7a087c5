"""Host-side pieces of the interpolating bicubic spline the reference builds with
``scipy.interpolate.RectBivariateSpline(row_centers, col_centers, Z)`` (kx=ky=3, s=0)
at neilpy/neilpy.py:1768-1774 and :1788-1790.

FITPACK's interpolating spline on sites x_i = i + 0.5 (i = 0..n-1) uses the knot vector
    t = [x_0]*4 + [x_2, ..., x_{n-3}] + [x_{n-1}]*4            (not-a-knot)
and solves the collocation system  sum_j B_j(x_i) c_j = f_i  per axis.  That system
depends only on n, so it is factored here once per length (banded LU without pivoting;
B-spline collocation matrices are totally positive, so this is stable) and the CUDA
prefilter only does the forward / backward substitutions along rows and columns.

Nothing in this module touches the oracle; the numpy evaluator at the bottom mirrors the
CUDA gather kernel and exists so the host logic can be tested against scipy on CPU.
"""
from __future__ import annotations

import functools

import numpy as np


def knot(j, n):
    """t[j] of the length-(n+4) not-a-knot vector for sites 0.5, 1.5, ..., n-0.5."""
    if j <= 3:
        return 0.5
    if j >= n:
        return n - 0.5
    return j - 1.5


def find_interval(x, n):
    """l with t[l] <= x < t[l+1], 3 <= l <= n-1 (x already clamped to [0.5, n-0.5])."""
    if x < 2.5:
        return 3
    return min(int(np.floor(x + 1.5)), n - 1)


def basis(x, l, n):
    """The four cubic B-splines that are non-zero on [t[l], t[l+1]] at x
    (de Boor-Cox recurrence, the arithmetic of FITPACK's fpbspl)."""
    h = [1.0, 0.0, 0.0, 0.0]
    for j in range(1, 4):
        hh = h[:j]
        h[0] = 0.0
        for i in range(1, j + 1):
            tli = knot(l + i, n)
            tlj = knot(l + i - j, n)
            f = hh[i - 1] / (tli - tlj)
            h[i - 1] = h[i - 1] + f * (tli - x)
            h[i] = f * (x - tlj)
    return h


@functools.lru_cache(maxsize=32)
def notaknot_factors(n):
    """LU factors of the n x n collocation matrix A[i][j] = B_j(x_i), x_i = i + 0.5.

    Returns a float64 array of shape (5, n): rows l1, l2, dinv, u1, u2 such that
        forward :  y_i = f_i - l1_i*y_{i-1} - l2_i*y_{i-2}
        backward:  c_i = (y_i - u1_i*c_{i+1} - u2_i*c_{i+2}) * dinv_i
    A has at most two sub- and two super-diagonals (row 1 and row n-2 carry four entries,
    every other row at most three)."""
    if n < 4:
        raise ValueError('the interpolating cubic spline needs at least 4 grid rows and columns')
    # band storage: band[i][d], d = j - i + 2, j in [i-2, i+2]
    band = [[0.0] * 5 for _ in range(n)]
    for i in range(n):
        if 6 <= i <= n - 7:
            band[i][1], band[i][2], band[i][3] = 1.0 / 6.0, 4.0 / 6.0, 1.0 / 6.0
            continue
        x = i + 0.5
        l = find_interval(x, n)
        h = basis(x, l, n)
        for a in range(4):
            j = l - 3 + a
            if h[a] != 0.0:
                d = j - i + 2
                assert 0 <= d <= 4, (n, i, j)
                band[i][d] = h[a]
    l1 = [0.0] * n
    l2 = [0.0] * n
    for i in range(n):
        # eliminate band[i][0] (col i-2) with row i-2, then band[i][1] (col i-1) with row i-1
        if i >= 2 and band[i][0] != 0.0:
            m = band[i][0] / band[i - 2][2]
            l2[i] = m
            band[i][1] -= m * band[i - 2][3]
            band[i][2] -= m * band[i - 2][4]
            band[i][0] = 0.0
        if i >= 1 and band[i][1] != 0.0:
            m = band[i][1] / band[i - 1][2]
            l1[i] = m
            band[i][2] -= m * band[i - 1][3]
            band[i][3] -= m * band[i - 1][4]
            band[i][1] = 0.0
    # note: eliminating col i-2 first with row i-2 uses U row i-2 (already final); the col i-1
    # multiplier must then act on the updated entry, which the order above guarantees.
    out = np.zeros((5, n), dtype=np.float64)
    out[0] = l1
    out[1] = l2
    out[2] = [1.0 / band[i][2] for i in range(n)]
    out[3] = [band[i][3] for i in range(n)]
    out[4] = [band[i][4] for i in range(n)]
    return out


def solve_axis0(f, fac):
    """numpy mirror of the CUDA substitution along axis 0 (sequential, exact)."""
    l1, l2, dinv, u1, u2 = fac
    n = f.shape[0]
    y = np.array(f, dtype=np.float64, copy=True)
    for i in range(1, n):
        y[i] -= l1[i] * y[i - 1]
        if i >= 2:
            y[i] -= l2[i] * y[i - 2]
    c = y
    c[n - 1] = y[n - 1] * dinv[n - 1]
    for i in range(n - 2, -1, -1):
        v = y[i] - u1[i] * c[i + 1]
        if i + 2 < n:
            v = v - u2[i] * c[i + 2]
        c[i] = v * dinv[i]
    return c


def prefilter(Z):
    """B-spline coefficients of the interpolating bicubic spline through Z (numpy mirror)."""
    Z = np.asarray(Z, dtype=np.float64)
    ny, nx = Z.shape
    c = solve_axis0(Z, notaknot_factors(ny))
    c = solve_axis0(c.T, notaknot_factors(nx)).T
    return np.ascontiguousarray(c)


def evaluate(coef, r, c):
    """numpy mirror of the CUDA gather: spline value at fractional (row, col) = (r, c)
    in centre coordinates, arguments clamped to [0.5, n-0.5] as FITPACK's bispeu does."""
    ny, nx = coef.shape
    r = np.atleast_1d(np.asarray(r, dtype=np.float64))
    c = np.atleast_1d(np.asarray(c, dtype=np.float64))
    out = np.empty(r.shape, dtype=np.float64)
    for k in range(r.size):
        rr = min(max(r[k], 0.5), ny - 0.5)
        cc = min(max(c[k], 0.5), nx - 0.5)
        lr = find_interval(rr, ny)
        lc = find_interval(cc, nx)
        hr = basis(rr, lr, ny)
        hc = basis(cc, lc, nx)
        sp = 0.0
        for a in range(4):
            for b in range(4):
                sp = sp + coef[lr - 3 + a, lc - 3 + b] * hr[a] * hc[b]
        out[k] = sp
    return out
