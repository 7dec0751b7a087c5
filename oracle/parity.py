"""End-to-end parity rule of the SMRF path -- TEST INFRASTRUCTURE ONLY (tests/, smoke()).

BASELINE.json north_star: "mask disagreements are only allowed where the point's residual lies
within that tolerance of the threshold".  The reference's harmonic fill is an inexact LSQR solve
(neilpy.py:1264, 2e-4 .. 1.1e-2 m from the exact fill) and the CUDA solver converges tighter, so a
decision that sits within TOL of its threshold may come out the other way.  This module computes,
from the ORACLE's own run, the margin of every decision and explains every disagreement:

  cell (progressive filter, neilpy.py:1671)   min_i | last_i - this_i - thr_i |
  cell (low-outlier pass, neilpy.py:1744)     | (-Z) - open(-Z) - low_filter_slope * cellsize |
  point (neilpy.py:1794-1795)                 | |elev - z| - (elevation_threshold + scaler * slope) |

A flipped cell is explained iff its margin <= TOL.  A flipped point is explained iff its margin
<= TOL, or it lies within NEAR cells of an explained flipped cell (such a cell is punched and
re-filled, which moves the DTM under the neighbouring points by decimetres; the bicubic spline's
influence decays by 0.268 per cell).
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage as ndi

from . import smrf_oracle as O

TOL = 2e-2     # metres: the reference LSQR's own distance from the exact harmonic fill
NEAR = 8       # cells


def progressive_margin(Z, windows, cellsize, slope_threshold):
    """min over the windows of |last - this - thr| at every cell, for the oracle's own surfaces."""
    windows = np.asarray(windows)
    thr = slope_threshold * (windows * cellsize)
    state = {'last': Z.copy(), 'margin': np.full(Z.shape, np.inf)}

    def hook(i, window, this_surface, new_obj):
        state['margin'] = np.minimum(state['margin'], np.abs(state['last'] - this_surface - thr[i]))
        if len(windows) > 1:
            state['last'] = this_surface.copy()

    O.progressive_filter(Z, windows, cellsize, slope_threshold, stage_hook=hook)
    return state['margin']


def explain(stages0, z, params, object_cells1, is_object_point1, tol=TOL, near=NEAR):
    """stages0: the `stages` dict of oracle.smrf; z: the point elevations; params: the smrf keyword
    arguments; object_cells1 / is_object_point1: the result under test.  Returns a dict of counts;
    parity holds iff unexplained_cells == unexplained_points == 0."""
    cellsize = params.get('cellsize', 1)
    windows = params.get('windows', 5)
    if np.isscalar(windows):
        windows = np.arange(windows) + 1
    oc0, op0 = stages0['object_cells'], np.asarray(stages0['is_object_point'])
    oc1, op1 = np.asarray(object_cells1, dtype=bool), np.asarray(is_object_point1, dtype=bool)
    cell_flip = oc0 != oc1
    out = {'cell_flips': int(cell_flip.sum()), 'point_flips': int((op0 != op1).sum()),
           'unexplained_cells': 0, 'unexplained_points': 0, 'max_cell_margin': 0.0, 'max_point_margin': 0.0}
    explained_cells = np.zeros_like(cell_flip)
    if cell_flip.any():
        m_prog = progressive_margin(stages0['Zmin_filtered'], windows, cellsize, params.get('slope_threshold', .15))
        m_low = progressive_margin(-stages0['Zmin_inpainted'], np.array([1]), cellsize, params.get('low_filter_slope', 5))
        margin = np.minimum(m_prog, m_low)
        out['max_cell_margin'] = float(margin[cell_flip].max())
        explained_cells = cell_flip & (margin <= tol)
        out['unexplained_cells'] = int((cell_flip & ~explained_cells).sum())
    flips = op0 != op1
    if flips.any():
        ev, sv = stages0['elevation_values'], stages0['slope_values']
        required = params.get('elevation_threshold', .5) + params.get('elevation_scaler', 1.25) * sv
        pm = np.abs(np.abs(ev - z) - required)
        near_cell = np.zeros_like(flips)
        if explained_cells.any():
            zone = ndi.binary_dilation(explained_cells, structure=np.ones((3, 3), bool), iterations=near)
            r = np.clip(np.floor(stages0['r']).astype(np.int64), 0, zone.shape[0] - 1)
            c = np.clip(np.floor(stages0['c']).astype(np.int64), 0, zone.shape[1] - 1)
            near_cell = zone[r, c]
        out['unexplained_points'] = int((flips & ~near_cell & (pm > tol)).sum())
        far = flips & ~near_cell
        out['max_point_margin'] = float(pm[far].max()) if far.any() else 0.0
    return out
