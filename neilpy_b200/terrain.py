"""Raster products of the DTM, the step after the SMRF path in the reference's notebooks
(SURVEY.md 8f rank 4): drop-ins for

    slope(Z, cellsize=1, z_factor=1, return_as='degrees')                       neilpy.py:456-467
    aspect(Z, return_as='degrees', flat_as='nan')                               neilpy.py:471-484
    hillshade(Z, cellsize=1, z_factor=1, zenith=45, azimuth=315, return_uint8=True)      :814-824
    pssm(Z, cellsize=1, ve=2.3, reverse=False, apply_colormap=True)                      :846-867

All four are one launch of smrf_terrain (csrc/terrain.cu); there is no CPU fallback.  numpy
in -> numpy out, CUDA tensor in -> CUDA tensor out (so `pssm(smrf(...)[0])` can stay on the
device).  float32 grids are widened to float64 per cell (the reference would differentiate a
float32 array in float32; its own DTMs are float64).

Differences from the reference, on purpose: an unsupported `return_as` raises ValueError (the
reference prints a message and then fails with UnboundLocalError); NaN cells give 0 in the
uint8 products (the reference's NaN -> uint8 cast is platform-defined).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

SLOPE, ASPECT, HILLSHADE, PSSM = 0, 1, 2, 3

# matplotlib's 'bone' colour map: (x, y) break points per channel (matplotlib/_cm.py, _bone_data;
# every segment is continuous, so one y per break point suffices)
_BONE_BREAKS = {'r': ((0., 0.), (0.746032, 0.652778), (1., 1.)),
                'g': ((0., 0.), (0.365079, 0.319444), (0.746032, 0.777778), (1., 1.)),
                'b': ((0., 0.), (0.365079, 0.444444), (1., 1.))}
_tables = {}


def bone_table(reverse=False, n=256):
    """matplotlib.cm.bone (reverse=False) / bone_r (reverse=True) as a 256 x 4 float64 RGBA table,
    built as LinearSegmentedColormap builds it: entry i interpolates the break points at
    i/(n-1), end entries are the end values; the reversed map mirrors the break points first."""
    key = (bool(reverse), n)
    if key in _tables:
        return _tables[key]
    cols = []
    for ch in 'rgb':
        pts = _BONE_BREAKS[ch]
        if reverse:
            pts = tuple((1.0 - x, y) for x, y in pts[::-1])
        bx = np.array([p[0] for p in pts]) * (n - 1)
        by = np.array([p[1] for p in pts])
        pos = (n - 1) * np.linspace(0, 1, n)
        col = np.empty(n)
        col[0], col[-1] = by[0], by[-1]
        hi = np.searchsorted(bx, pos[1:-1])
        frac = (pos[1:-1] - bx[hi - 1]) / (bx[hi] - bx[hi - 1])
        col[1:-1] = frac * (by[hi] - by[hi - 1]) + by[hi - 1]
        cols.append(np.clip(col, 0, 1))
    cols.append(np.ones(n))
    _tables[key] = np.stack(cols, 1)
    return _tables[key]


def _grid(Z):
    from .api import _device
    dev = _device()
    if isinstance(Z, torch.Tensor):
        t, on_device = Z, Z.is_cuda
    else:
        t, on_device = torch.from_numpy(np.ascontiguousarray(np.asarray(Z))), False
    if t.dim() != 2:
        raise ValueError('Z must be a 2-D grid')
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    return t.to(dev).contiguous(), on_device, dev


def _run(Z, mode, *, return_as=0, spacing=1.0, flat_is_nan=0, flat_value=0.0, cos_zen=0.0, sin_zen=0.0, azimuth=0.0,
         ve=1.0, want='f64', lut=None):
    from .api import _code, _ptr, _stream, _to_host
    lib = _lib.load()
    grid, on_device, dev = _grid(Z)
    ny, nx = grid.shape
    out_f64 = torch.empty((ny, nx), dtype=torch.float64, device=dev) if want == 'f64' else None
    out_u8 = torch.empty((ny, nx), dtype=torch.uint8, device=dev) if want == 'u8' else None
    rgba = torch.empty((ny, nx, 4), dtype=torch.float64, device=dev) if want == 'rgba' else None
    lut_dev = torch.from_numpy(lut).to(dev) if lut is not None else None
    _lib.check(lib.smrf_terrain(_ptr(grid), ny, nx, _code(grid.dtype), mode, return_as, float(spacing), int(flat_is_nan),
                                float(flat_value), float(cos_zen), float(sin_zen), float(azimuth), float(ve),
                                _ptr(out_f64), _ptr(out_u8), _ptr(rgba), _ptr(lut_dev), _stream()), 'smrf_terrain')
    out = out_f64 if want == 'f64' else (out_u8 if want == 'u8' else rgba)
    if on_device:
        return out
    return _to_host(out)


_RETURN_AS = {'percent': 0, 'radians': 1, 'degrees': 2}


def slope(Z, cellsize=1, z_factor=1, return_as='degrees'):
    if return_as not in _RETURN_AS:
        raise ValueError('return_as %r is not supported.' % (return_as,))
    return _run(Z, SLOPE, return_as=_RETURN_AS[return_as], spacing=cellsize / z_factor)


def aspect(Z, return_as='degrees', flat_as='nan'):
    if return_as not in ('degrees', 'radians'):
        raise ValueError('return_as %r is not supported.' % (return_as,))
    nan = isinstance(flat_as, str) and flat_as == 'nan'
    return _run(Z, ASPECT, return_as=_RETURN_AS[return_as], flat_is_nan=int(nan), flat_value=0.0 if nan else float(flat_as))


def hillshade(Z, cellsize=1, z_factor=1, zenith=45, azimuth=315, return_uint8=True):
    zen, az = np.deg2rad((zenith, azimuth))                 # as the reference converts them (:815)
    return _run(Z, HILLSHADE, spacing=cellsize / z_factor, cos_zen=np.cos(zen), sin_zen=np.sin(zen), azimuth=az,
                want='u8' if return_uint8 else 'f64')


def pssm(Z, cellsize=1, ve=2.3, reverse=False, apply_colormap=True):
    if not apply_colormap:
        return _run(Z, PSSM, spacing=cellsize, ve=ve, want='u8')
    # reverse=False is the *reversed* bone map, white for flat ground (:861-864)
    return _run(Z, PSSM, spacing=cellsize, ve=ve, want='rgba', lut=bone_table(reverse=not reverse))
