"""CPU oracle for the raster products computed from the DTM right after the SMRF path
 -- TEST INFRASTRUCTURE ONLY.

Restates, in numpy:

    slope      neilpy/neilpy.py:456-467    np.gradient(Z, cellsize/z_factor), sqrt(gx^2+gy^2),
                                           arctan / rad2deg as asked
    aspect     neilpy/neilpy.py:471-484    arctan2(gy, -gx) of the unit-spacing gradient, turned
                                           to a compass bearing, flat cells = flat_as
    hillshade  neilpy/neilpy.py:814-824    cos(zen)cos(S) + sin(zen)sin(S)cos(az - A), clipped at 0,
                                           x255, rounded, uint8
    pssm       neilpy/neilpy.py:846-867    round(255 * rad2deg(arctan(ve*S)) / 90) as uint8, then
                                           matplotlib's bone_r (or bone) colour map

`bone_lut` restates matplotlib's 'bone' colour map from its published segment data
(matplotlib/_cm.py `_bone_data`) and LinearSegmentedColormap's 256-entry table
(matplotlib.colors._create_lookup_table, `reversed()` for bone_r); matplotlib is absent here.

Parity pins (tests/golden/make_terrain_golden.py, tests/test_terrain_oracle.py):
  * slope, aspect, hillshade and pssm are the reference's OWN functions, executed from the
    source text of /root/reference/neilpy/neilpy.py at generation time (they need numpy only;
    pssm gets this module's colour map in place of plt.cm), on small synthetic DEMs -- their
    outputs are stored in tests/golden/terrain_golden.npz and this restatement must equal them
    bit for bit.
  * the colour table: the reference ships examples/dk22_smrfed_bonemap.png, a pssm() image
    saved with plt.imsave.  All of its 256 distinct colours are exactly floor(255 * bone_r
    table) of this restatement (stored in the golden file as `bonemap_png_colours`).

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

_BONE = {'red': ((0., 0., 0.), (0.746032, 0.652778, 0.652778), (1.0, 1.0, 1.0)),
         'green': ((0., 0., 0.), (0.365079, 0.319444, 0.319444), (0.746032, 0.777778, 0.777778), (1.0, 1.0, 1.0)),
         'blue': ((0., 0., 0.), (0.365079, 0.444444, 0.444444), (1.0, 1.0, 1.0))}


def _table(data, n=256):
    a = np.array(data)
    x, y0, y1 = a[:, 0] * (n - 1), a[:, 1], a[:, 2]
    xi = (n - 1) * np.linspace(0, 1, n)
    k = np.searchsorted(x, xi)[1:-1]
    w = (xi[1:-1] - x[k - 1]) / (x[k] - x[k - 1])
    return np.clip(np.concatenate([[y1[0]], w * (y0[k] - y1[k - 1]) + y1[k - 1], [y0[-1]]]), 0, 1)


def bone_lut(reverse=False):
    """256 x 4 float64 RGBA table of matplotlib's `bone` (reverse=False) or `bone_r`."""
    chan = []
    for c in ('red', 'green', 'blue'):
        d = _BONE[c]
        if reverse:
            d = [(1.0 - x, b, a) for x, a, b in reversed(d)]
        chan.append(_table(d))
    chan.append(np.ones(256))
    return np.stack(chan, 1)


class _Cmap:
    def __init__(self, lut):
        self.lut = lut

    def __call__(self, idx):
        return self.lut[np.asarray(idx).astype(np.intp)]


class plt_stub:                       # what pssm needs of matplotlib.pyplot
    class cm:
        bone = _Cmap(bone_lut(False))
        bone_r = _Cmap(bone_lut(True))


def slope(Z, cellsize=1, z_factor=1, return_as='degrees'):
    gy, gx = np.gradient(Z, cellsize / z_factor)
    S = np.sqrt(gx ** 2 + gy ** 2)
    if return_as in ('degrees', 'radians'):
        S = np.arctan(S)
        if return_as == 'degrees':
            S = np.rad2deg(S)
    return S


def aspect(Z, return_as='degrees', flat_as='nan'):
    gy, gx = np.gradient(Z)
    A = np.pi / 2 - np.arctan2(gy, -gx)
    A[A < 0] = A[A < 0] + 2 * np.pi
    if return_as == 'degrees':
        A = np.rad2deg(A)
    A[(gx == 0) & (gy == 0)] = np.nan if flat_as == 'nan' else flat_as
    return A


def hillshade(Z, cellsize=1, z_factor=1, zenith=45, azimuth=315, return_uint8=True):
    zenith, azimuth = np.deg2rad((zenith, azimuth))
    S = slope(Z, cellsize=cellsize, z_factor=z_factor, return_as='radians')
    A = aspect(Z, return_as='radians', flat_as=0)
    H = (np.cos(zenith) * np.cos(S)) + (np.sin(zenith) * np.sin(S) * np.cos(azimuth - A))
    H[H < 0] = 0
    if return_uint8:
        H = np.round(255 * H).astype(np.uint8)
    return H


def pssm(Z, cellsize=1, ve=2.3, reverse=False, apply_colormap=True):
    gy, gx = np.gradient(Z, cellsize)
    S = np.sqrt(gx ** 2 + gy ** 2)
    P = np.round(255 * (np.rad2deg(np.arctan(ve * S)) / 90)).astype(np.uint8)
    if not apply_colormap:
        return P
    return bone_lut(reverse=not reverse)[P]          # reverse=False -> bone_r (neilpy.py:861-864)


def synth_dem(ny, nx, seed=0, flat=True):
    """Small float64 test DEM: smooth hills + noise, a tilted plane patch, and an exactly flat patch."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:ny, 0:nx].astype(np.float64)
    Z = 40 * np.sin(x / 9.0 + .3) * np.cos(y / 7.0) + 0.8 * x - 0.3 * y + 300 + rng.normal(0, .4, (ny, nx))
    Z[: ny // 4, : nx // 4] = (0.5 * x - 0.25 * y)[: ny // 4, : nx // 4]
    if flat:
        Z[-(ny // 4):, -(nx // 4):] = 123.25
    return np.round(Z * 1024) / 1024
