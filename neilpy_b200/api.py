"""The reference call surface of the SMRF path, served by libsmrf_b200.so on a B200.

    smrf                     <- neilpy/neilpy.py:1685-1808
    create_dem               <- neilpy/neilpy.py:1110-1166
    progressive_filter       <- neilpy/neilpy.py:1659-1680
    inpaint_nans_by_springs  <- neilpy/neilpy.py:1227-1271

Same names, positional order, defaults, return shapes and exceptions as the reference.
Host code here only does what the reference does with host scalars (grid geometry via
np.arange, the affine transform, the per-window thresholds, the spline's banded
factorisation) and moves buffers; every per-point and per-cell operation runs in the CUDA
library.  There is no CPU fallback.

Array-likes in, numpy out (H2D / D2H copies included) -- or CUDA torch tensors in, CUDA
torch tensors out (nothing leaves the device).  Two extra keyword arguments exist:
`dtype` (grid element type: float64 = the reference's arithmetic, float32 = throughput
mode; default follows the input) and, on `smrf`, `inpaint_tol` (residual max-norm in
metres at which the harmonic solver stops).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import warnings

import numpy as np
import torch

from . import _lib
from . import spline as _spline
from .affine import Affine

__all__ = ['smrf', 'create_dem', 'progressive_filter', 'inpaint_nans_by_springs', 'inpaint_nans_by_fda', 'Affine',
           'InpaintWarning']

INPAINT_TOL = 1e-9       # metres, max-norm of the residual deg*u - sum(nbrs): <= 1e-6 m from the exact fill
SMRF_INPAINT_TOL = 1e-6  # inside smrf(): residual max-norm; the fill ends <= ~1e-3 m from the exact one, 10x tighter than the reference's own LSQR (1.5e-2 m)
INPAINT_MAX_ITER = 1 << 15


# --------------------------------------------------------------------------- helpers
def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('neilpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _grid_dtype(dtype, default):
    if dtype is None:
        return default
    if dtype in (torch.float32, np.float32, 'float32', 'f32', 'f4'):
        return torch.float32
    if dtype in (torch.float64, np.float64, 'float64', 'f64', 'f8', float):
        return torch.float64
    try:
        nd = np.dtype(dtype)
        if nd == np.float32:
            return torch.float32
        if nd == np.float64:
            return torch.float64
    except TypeError:
        pass
    raise ValueError('dtype must be float32 or float64')


def _code(tdtype):
    return _lib.F32 if tdtype == torch.float32 else _lib.F64


def _make_transform(west, north, cellsize):
    """rasterio.transform.from_origin(west, north, cellsize, cellsize) (neilpy.py:1141)."""
    try:
        import affine  # the real thing, if the environment has it
        return affine.Affine(float(cellsize), 0.0, float(west), 0.0, -float(cellsize), float(north))
    except ImportError:
        return Affine.from_origin(west, north, cellsize, cellsize)


def _inverse6(t):
    """~t, with the arithmetic of affine.Affine.__invert__."""
    a, b, c, d, e, f = [float(v) for v in tuple(t)[:6]]
    inv = ~Affine(a, b, c, d, e, f)
    return (C.c_double * 6)(*inv[:6])


class _Points:
    """x, y, z on the device in one of the library's three stream layouts."""

    def __init__(self, x, y, z, dev):
        self.index = None          # pandas index of z, if z was a Series (the reference returns a Series then)
        self.on_device = False
        if y is None and z is None:
            a = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
            if a.dim() != 2 or a.shape[1] != 4 or a.dtype != torch.float32:
                raise ValueError('a single point argument must be an (N, 4) float32 array (x, y, z, unused)')
            self.on_device = a.is_cuda
            self.a = a.to(dev, non_blocking=True).contiguous()
            self.fmt, self.n = _lib.PTS_XYZW_F32, int(a.shape[0])
            self.ptrs = (_ptr(self.a), C.c_void_p(0), C.c_void_p(0))
            self.default_dtype = torch.float32
            return
        if hasattr(z, 'index') and hasattr(z, 'values') and not isinstance(z, torch.Tensor):
            self.index = z.index
        ts = []
        for v in (x, y, z):
            if isinstance(v, torch.Tensor):
                t = v
            else:
                arr = np.asarray(v.values if hasattr(v, 'values') else v)
                if arr.dtype != np.float32:
                    arr = arr.astype(np.float64, copy=False)
                with warnings.catch_warnings():           # read-only views (pandas copy-on-write) are only read here
                    warnings.simplefilter('ignore', UserWarning)
                    t = torch.from_numpy(np.ascontiguousarray(arr))
            ts.append(t.reshape(-1))
        self.on_device = all(t.is_cuda for t in ts)
        if all(t.dtype == torch.float32 for t in ts):
            want, self.fmt, self.default_dtype = torch.float32, _lib.PTS_SOA_F32, torch.float32
        else:
            want, self.fmt, self.default_dtype = torch.float64, _lib.PTS_SOA_F64, torch.float64
        self.x, self.y, self.z = [t.to(device=dev, dtype=want, non_blocking=True).contiguous() for t in ts]
        if not (self.x.numel() == self.y.numel() == self.z.numel()):
            raise ValueError('x, y and z must have the same length')
        self.n = int(self.x.numel())
        self.ptrs = (_ptr(self.x), _ptr(self.y), _ptr(self.z))


def _extent(lib, pts, dev):
    out4 = torch.empty(4, dtype=torch.float64, device=dev)
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    scratch = torch.empty(4, dtype=torch.int64, device=dev)
    _lib.check(lib.smrf_extent(pts.ptrs[0], pts.ptrs[1], pts.n, pts.fmt, _ptr(out4), _ptr(bad), _ptr(scratch),
                               _stream()), 'smrf_extent')
    e = out4.cpu().numpy()
    if int(bad.item()) or pts.n == 0:
        raise ValueError('x and y must be finite and non-empty (np.arange / np.ravel_multi_index fail in the reference)')
    return float(e[0]), float(e[1]), float(e[2]), float(e[3])


def _edges(xmin, xmax, ymin, ymax, cellsize):
    """neilpy.py:1113-1124, on host scalars, with numpy so that len(np.arange(...)) is the reference's."""
    floor2 = lambda x, v: v * np.floor(x / v)
    ceil2 = lambda x, v: v * np.ceil(x / v)
    xedges = np.arange(floor2(xmin, cellsize) - .5 * cellsize, ceil2(xmax, cellsize) + 1.5 * cellsize, cellsize)
    yedges = np.arange(ceil2(ymax, cellsize) + .5 * cellsize, floor2(ymin, cellsize) - 1.5 * cellsize, -cellsize)
    return xedges, yedges


def _bin(lib, pts, dev, tdtype, cellsize, bin_type, edges):
    """create_dem without the optional inpaint: returns grid, empty mask, transform, inverse, cellsize."""
    if bin_type == 'max':
        bt = _lib.BIN_MAX
    elif bin_type == 'min':
        bt = _lib.BIN_MIN
    else:
        raise ValueError('This type not supported.')
    if edges is None:
        xmin, xmax, ymin, ymax = _extent(lib, pts, dev)
        xedges, yedges = _edges(xmin, xmax, ymin, ymax, cellsize)
        strict = True
    else:
        xedges, yedges = np.asarray(edges[0]), np.asarray(edges[1])
        cellsize = np.abs(xedges[1] - xedges[0])
        strict = False   # the reference drops out-of-range points first (neilpy.py:1128-1131)
    nx, ny = len(xedges) - 1, len(yedges) - 1
    if nx <= 0 or ny <= 0:
        raise ValueError('empty grid')
    t = _make_transform(xedges[0], yedges[0], cellsize)
    inv6 = _inverse6(t)
    grid = torch.empty((ny, nx), dtype=tdtype, device=dev)
    empty = torch.empty((ny, nx), dtype=torch.uint8, device=dev)
    oor = torch.zeros(1, dtype=torch.int64, device=dev)
    code, st = _code(tdtype), _stream()
    _lib.check(lib.smrf_bin_init(_ptr(grid), ny, nx, code, bt, st), 'smrf_bin_init')
    _lib.check(lib.smrf_bin_accumulate(pts.ptrs[0], pts.ptrs[1], pts.ptrs[2], pts.n, pts.fmt, inv6, _ptr(grid), ny, nx,
                                       code, bt, _ptr(oor), st), 'smrf_bin_accumulate')
    _lib.check(lib.smrf_bin_finalize(_ptr(grid), _ptr(empty), ny, nx, code, bt, st), 'smrf_bin_finalize')
    if strict and int(oor.item()):
        raise ValueError('invalid entry in coordinates array')   # np.ravel_multi_index's message
    return grid, empty, t, inv6, cellsize


_pinned = {}
_copy_pool = None
_side_streams = {}


def _side_stream(dev):
    s = _side_streams.get(dev.index)
    if s is None:
        s = _side_streams[dev.index] = torch.cuda.Stream(device=dev)
    return s


def _copy_threads():
    """Host threads for the staging -> result copy.  Explicit, because torchrun exports
    OMP_NUM_THREADS=1 and torch's own host copy then runs on one core (175 MB: 110 ms instead of 20)."""
    cpus = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    ranks = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1))
    return max(1, min(4, cpus // ranks))


def _threaded_copy(dst, src):
    """dst[...] = src for two flat, equally long numpy arrays, split over a few threads (numpy
    releases the GIL while it copies; first-touch page faults of a fresh `dst` parallelise too)."""
    global _copy_pool
    n = dst.size
    k = _copy_threads() if n * dst.itemsize >= (4 << 20) else 1
    if k == 1:
        np.copyto(dst, src)
        return
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _copy_pool = ThreadPoolExecutor(max_workers=4, thread_name_prefix='smrf-copy')
    cuts = [n * i // k for i in range(k + 1)]
    list(_copy_pool.map(lambda i: np.copyto(dst[cuts[i]:cuts[i + 1]], src[cuts[i]:cuts[i + 1]]), range(k)))


class _HostCopy:
    """Device tensor -> numpy through a cached pinned staging buffer (pageable D2H copies run at a fraction of the
    link rate).  start() enqueues the D2H on `stream`; result() waits for it and copies the staging buffer into an
    array the caller owns, so a copy started on a side stream overlaps the kernels that are still running."""

    def __init__(self, t, stream=None):
        self.shape = tuple(t.shape)
        # per thread (virtual ranks are threads of one process) and per role: a staging buffer is never shared
        key = (t.dtype, t.numel(), 0 if stream is None else 1, threading.get_ident())
        buf = _pinned.get(key)
        if buf is None:
            if len(_pinned) > 32:
                _pinned.clear()
            buf = _pinned[key] = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
        self.buf = buf
        self.stream = stream if stream is not None else torch.cuda.current_stream()
        self.src = t                      # keep the device tensor alive until the copy has run
        with torch.cuda.stream(self.stream):
            buf.copy_(t.reshape(-1), non_blocking=True)
        self.done = torch.cuda.Event()
        self.done.record(self.stream)

    def result(self):
        out = np.empty(self.buf.numel(), dtype=self.buf.numpy().dtype)         # pageable, owned by the caller
        self.done.synchronize()
        _threaded_copy(out, self.buf.numpy())
        self.src = None
        return out.reshape(self.shape)


def _to_host(t):
    return _HostCopy(t).result()


def _workspace(nbytes, dev):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=dev)


def _inpaint(lib, grid, ws, tol, unknown=None, guess=None):
    ny, nx = grid.shape
    need = lib.smrf_inpaint_workspace_bytes(ny, nx)
    if ws is None or ws.numel() < need:
        ws = _workspace(need, grid.device)
    info = (C.c_double * 3)()
    _lib.check(lib.smrf_inpaint(_ptr(grid), ny, nx, _code(grid.dtype), _ptr(unknown), _ptr(guess), _ptr(ws), ws.numel(),
                                float(tol), INPAINT_MAX_ITER, info, _stream()), 'smrf_inpaint')
    return _converged({'iterations': int(info[0]), 'residual': float(info[1]), 'unknown': int(info[2])}, tol)


class InpaintWarning(RuntimeWarning):
    """The harmonic solver stopped above its tolerance (iteration cap, or a non-finite residual
    caused by non-finite elevations).  The reference's LSQR reports nothing in that case either,
    but a silently unconverged DTM would flow into the classification."""


def _converged(info, tol, what='inpaint_nans_by_springs'):
    r = info['residual']
    info['converged'] = bool(info['unknown'] == 0 or (np.isfinite(r) and r <= tol))
    if not info['converged']:
        warnings.warn('%s: residual %.3g m after %d iterations (tolerance %.3g m)%s'
                      % (what, r, info['iterations'], tol, '' if np.isfinite(r) else ' -- non-finite elevations in the grid?'),
                      InpaintWarning, stacklevel=3)
    return info


def _progressive(lib, surface, windows, thresholds, mask, when, ws, negate=0, last_out=None):
    ny, nx = surface.shape
    need = lib.smrf_open_workspace_bytes(ny, nx, _code(surface.dtype), int(max(windows)) if len(windows) else 0)
    if ws is None or ws.numel() < need:
        ws = _workspace(need, surface.device)
    w = (C.c_int32 * len(windows))(*[int(v) for v in windows])
    th = (C.c_double * len(windows))(*[float(v) for v in thresholds])
    _lib.check(lib.smrf_progressive_open(_ptr(surface), _ptr(ws), ws.numel(), _ptr(mask), _ptr(when), ny, nx,
                                         _code(surface.dtype), w, th, len(windows), negate, _ptr(last_out), _stream()),
               'smrf_progressive_open')


def _windows(windows):
    if np.isscalar(windows):
        windows = np.arange(windows) + 1                        # neilpy.py:1738-1739
    windows = np.asarray(windows)
    if windows.ndim != 1:
        raise ValueError('windows must be a scalar or a 1-D array of radii')
    if len(windows) and (np.any(windows < 0) or np.any(windows != np.floor(windows))):
        raise ValueError('window radii must be non-negative integers')
    return windows


_factor_cache = {}


def _factors(n, dev):
    key = (n, str(dev))
    if key not in _factor_cache:
        _factor_cache[key] = torch.from_numpy(_spline.notaknot_factors(n)).to(dev).contiguous()
    return _factor_cache[key]


# --------------------------------------------------------------------------- public API
def create_dem(x, y, z, cellsize=1, bin_type='max', inpaint=False, edges=None, use_binned_statistic=False,
               dtype=None):
    """neilpy.create_dem (neilpy.py:1110-1166): returns (I, t)."""
    if use_binned_statistic:
        raise NotImplementedError('use_binned_statistic=True is a broken branch in the reference '
                                  '(it returns a scipy result object, neilpy.py:1148-1149)')
    lib, dev = _lib.load(), _device()
    pts = _Points(x, y, z, dev)
    tdtype = _grid_dtype(dtype, pts.default_dtype)
    grid, _, t, _, _ = _bin(lib, pts, dev, tdtype, cellsize, bin_type, edges)
    if inpaint == True:  # noqa: E712  (the reference's own test)
        _inpaint(lib, grid, None, INPAINT_TOL)
    return (grid if pts.on_device else grid.cpu().numpy()), t


def inpaint_nans_by_springs(A, inplace=False, neighbors=4, tol=INPAINT_TOL, return_info=False):
    """neilpy.inpaint_nans_by_springs (neilpy.py:1227-1271).  Only 4 neighbours, as the reference."""
    lib, dev = _lib.load(), _device()
    if isinstance(A, torch.Tensor):
        on_device = A.is_cuda
        if A.dtype not in (torch.float32, torch.float64):
            raise ValueError('A must be float32 or float64')
        g = A.to(dev) if (inplace and on_device) else A.to(dev, copy=True)
        g = g.contiguous()
    else:
        on_device = False
        arr = np.asarray(A)
        if arr.dtype != np.float32:
            arr = arr.astype(np.float64, copy=False)
        g = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    if g.dim() != 2:
        raise ValueError('A must be 2-D')
    info = _inpaint(lib, g, None, tol)
    if inplace:
        if isinstance(A, torch.Tensor):
            if g.data_ptr() != A.data_ptr():
                A.copy_(g)
        else:
            A[...] = g.cpu().numpy()
        return None
    out = g if on_device else g.cpu().numpy()
    return (out, info) if return_info else out


FDA_TOL = 1e-10          # max |A^T r| of the normal equations at which CGLS stops
FDA_MAX_ITER = 60000


def inpaint_nans_by_fda(A, fast=True, inplace=False, tol=FDA_TOL, return_info=False):
    """neilpy.inpaint_nans_by_fda (neilpy.py:1171-1216): least-squares finite-difference fill of the NaN cells.
    `fast` only prunes equations that cannot touch a NaN in the reference; the answer does not depend on it."""
    lib, dev = _lib.load(), _device()
    if isinstance(A, torch.Tensor):
        on_device = A.is_cuda
        if A.dtype not in (torch.float32, torch.float64):
            raise ValueError('A must be float32 or float64')
        g = (A.to(dev) if (inplace and on_device) else A.to(dev, copy=True)).contiguous()
    else:
        on_device = False
        arr = np.asarray(A)
        if arr.dtype != np.float32:
            arr = arr.astype(np.float64, copy=False)
        g = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    if g.dim() != 2:
        raise ValueError('A must be 2-D')
    ny, nx = g.shape
    ws = _workspace(lib.smrf_inpaint_fda_workspace_bytes(ny, nx), dev)
    info = (C.c_double * 3)()
    _lib.check(lib.smrf_inpaint_fda(_ptr(g), ny, nx, _code(g.dtype), _ptr(ws), ws.numel(), float(tol), FDA_MAX_ITER, info,
                                    _stream()), 'smrf_inpaint_fda')
    info = _converged({'iterations': int(info[0]), 'residual': float(info[1]), 'unknown': int(info[2])}, tol,
                      'inpaint_nans_by_fda')
    if inplace:
        if isinstance(A, torch.Tensor):
            if g.data_ptr() != A.data_ptr():
                A.copy_(g)
        else:
            A[...] = g.cpu().numpy()
        return None
    out = g if on_device else g.cpu().numpy()
    return (out, info) if return_info else out


def progressive_filter(Z, windows, cellsize=1, slope_threshold=.15, return_when_dropped=False):
    """neilpy.progressive_filter (neilpy.py:1659-1680).  `windows` is a 1-D array of radii."""
    lib, dev = _lib.load(), _device()
    windows = np.asarray(windows)
    if isinstance(Z, torch.Tensor):
        on_device = Z.is_cuda
        if Z.dtype not in (torch.float32, torch.float64):
            raise ValueError('Z must be float32 or float64')
        surf = Z.to(dev).contiguous()      # the library never writes the input surface
    else:
        on_device = False
        arr = np.asarray(Z)
        if arr.dtype != np.float32:
            arr = arr.astype(np.float64, copy=False)
        surf = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    if surf.dim() != 2:
        raise ValueError('Z must be 2-D')
    thresholds = slope_threshold * (windows * cellsize)          # neilpy.py:1661
    mask = torch.zeros(surf.shape, dtype=torch.uint8, device=dev)
    when = torch.zeros(surf.shape, dtype=torch.uint8, device=dev) if return_when_dropped else None
    if len(windows):
        _progressive(lib, surf, windows, thresholds, mask, when, None)
    mask = mask.view(torch.bool)
    if not on_device:
        mask = mask.cpu().numpy()
        when = when.cpu().numpy() if when is not None else None
    return (mask, when) if return_when_dropped else mask


def smrf(x, y=None, z=None, cellsize=1, windows=5, slope_threshold=.15, elevation_threshold=.5,
         elevation_scaler=1.25, low_filter_slope=5, low_outlier_fill=False, return_extras=False,
         dtype=None, inpaint_tol=SMRF_INPAINT_TOL, return_stages=None):
    """neilpy.smrf (neilpy.py:1685-1808).

    Returns (Zpro, t, object_cells, is_object_point) [+ extras dict].  `return_stages`, if a
    dict, receives the device tensors of every intermediate stage (for the parity tests).
    """
    lib, dev = _lib.load(), _device()
    windows = _windows(windows)
    pts = _Points(x, y, z, dev)
    tdtype = _grid_dtype(dtype, pts.default_dtype)
    code, st = _code(tdtype), _stream()
    stages = return_stages

    # --- create_dem(bin_type='min')  (:1741) and is_empty_cell (:1742)
    Zmin, empty, t, inv6, _ = _bin(lib, pts, dev, tdtype, cellsize, 'min', None)
    ny, nx = Zmin.shape
    if ny < 4 or nx < 4:
        # scipy's RectBivariateSpline (FITPACK) refuses grids with fewer than 4 rows or columns
        raise ValueError('the grid must have at least 4 rows and 4 columns for the bicubic spline')
    ws = _workspace(max(lib.smrf_inpaint_workspace_bytes(ny, nx),
                        lib.smrf_open_workspace_bytes(ny, nx, code, int(windows.max()) if len(windows) else 0),
                        lib.smrf_spline_workspace_bytes(ny, nx)), dev)
    if stages is not None:
        stages['Zmin_binned'] = Zmin.clone()
    # --- first inpaint (:1743)
    info1 = _inpaint(lib, Zmin, ws, inpaint_tol)
    if stages is not None:
        stages['Zmin_inpainted'] = Zmin.clone()
    # --- low outliers: progressive_filter(-Zmin, [1], cellsize, low_filter_slope)  (:1744)
    low = torch.zeros((ny, nx), dtype=torch.uint8, device=dev)
    one = np.array([1])
    _progressive(lib, Zmin, one, low_filter_slope * (one * cellsize), low, None, ws, negate=1)
    if low_outlier_fill:                                          # :1747-1749
        _lib.check(lib.smrf_merge_punch(_ptr(Zmin), None, _ptr(low), None, None, ny, nx, code, st), 'smrf_merge_punch')
        _inpaint(lib, Zmin, ws, inpaint_tol)
    if stages is not None:
        stages['low_outliers'] = low.clone()
        stages['Zmin_filtered'] = Zmin.clone()
    # --- the progressive morphological filter (:1752-1755)
    obj = torch.zeros((ny, nx), dtype=torch.uint8, device=dev)
    drop = torch.zeros((ny, nx), dtype=torch.uint8, device=dev) if return_extras else None
    opened = None
    if len(windows):
        # the last opening is the surface with the objects shaved off: the seed of the second inpaint
        opened = torch.empty_like(Zmin)
        _progressive(lib, Zmin, windows, slope_threshold * (windows * cellsize), obj, drop, ws, last_out=opened)
    if stages is not None:
        stages['progressive_cells'] = obj.clone()
    # --- object_cells = empty | low | obj; Zpro[object_cells] = nan; second inpaint (:1758-1764)
    Zpro = Zmin
    object_cells = torch.empty((ny, nx), dtype=torch.uint8, device=dev)
    _lib.check(lib.smrf_merge_punch(_ptr(Zpro), _ptr(empty), _ptr(low), _ptr(obj), _ptr(object_cells), ny, nx, code, st),
               'smrf_merge_punch')
    if stages is not None:
        stages['Zpro_punched'] = Zpro.clone()
    info2 = _inpaint(lib, Zpro, ws, inpaint_tol, guess=opened)
    del opened
    early = None
    if not pts.on_device:
        # host in -> host out: the DTM and the cell mask are final here; their D2H copies run on a side stream under
        # the slope / spline / classification kernels that follow
        side = _side_stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        early = (_HostCopy(Zpro, side), _HostCopy(object_cells, side))
    # --- slope raster (:1785-1786) and the two interpolating splines (:1773, :1788)
    S = torch.empty_like(Zpro)
    _lib.check(lib.smrf_slope(_ptr(Zpro), _ptr(S), ny, nx, code, float(cellsize), st), 'smrf_slope')
    rowf, colf = _factors(ny, dev), _factors(nx, dev)
    # the two coefficient sets are interleaved, [ny][nx][2]: one gather sector serves both splines
    coef = torch.empty((ny, nx, 2), dtype=Zpro.dtype, device=dev)
    _lib.check(lib.smrf_spline_prefilter(_ptr(Zpro), _ptr(coef), 2, 0, ny, nx, code, _ptr(rowf), _ptr(colf), _ptr(ws),
                                         ws.numel(), st), 'smrf_spline_prefilter')
    if stages is not None:
        stages['S'] = S.clone()
    _lib.check(lib.smrf_spline_prefilter(_ptr(S), _ptr(coef), 2, 1, ny, nx, code, _ptr(rowf), _ptr(colf), _ptr(ws),
                                         ws.numel(), st), 'smrf_spline_prefilter')
    # --- interpolate + classify every point (:1772-1795)
    is_obj = torch.empty(pts.n, dtype=torch.uint8, device=dev)
    want_vals = return_extras or stages is not None
    elev = torch.empty(pts.n, dtype=torch.float64, device=dev) if want_vals else None
    slp = torch.empty(pts.n, dtype=torch.float64, device=dev) if stages is not None else None
    when_pt = torch.empty(pts.n, dtype=torch.uint8, device=dev) if return_extras else None
    _lib.check(lib.smrf_classify(pts.ptrs[0], pts.ptrs[1], pts.ptrs[2], pts.n, pts.fmt, inv6, _ptr(coef), None,
                                 ny, nx, code, float(elevation_threshold), float(elevation_scaler), _ptr(is_obj),
                                 _ptr(elev), _ptr(slp), _ptr(drop), _ptr(when_pt), st), 'smrf_classify')
    if stages is not None:
        stages.update(Zpro=Zpro, object_cells=object_cells, coef_z=coef[:, :, 0], coef_s=coef[:, :, 1], elevation_values=elev,
                      slope_values=slp, is_object_point=is_obj, inpaint1=info1, inpaint2=info2)

    object_cells = object_cells.view(torch.bool)
    is_obj = is_obj.view(torch.bool)
    extras = None
    if return_extras:                                             # :1797-1801
        zvals = pts.a[:, 2].to(torch.float64) if pts.fmt == _lib.PTS_XYZW_F32 else pts.z.to(torch.float64)
        extras = {'above_ground_height': zvals - elev, 'drop_raster': drop, 'when_dropped': when_pt}
    if not pts.on_device:
        last = _HostCopy(is_obj)
        Zpro = early[0].result()
        object_cells = early[1].result().view(np.bool_)
        is_obj = last.result().view(np.bool_)
        if pts.index is not None:
            import pandas as pd
            is_obj = pd.Series(is_obj, index=pts.index)
        if extras is not None:
            extras = {k: v.cpu().numpy() for k, v in extras.items()}
            if pts.index is not None:
                import pandas as pd
                extras['above_ground_height'] = pd.Series(extras['above_ground_height'], index=pts.index)
    if return_extras:
        return Zpro, t, object_cells, is_obj, extras
    return Zpro, t, object_cells, is_obj
