"""ctypes binding of libsmrf_b200.so (C ABI: include/smrf_b200.h).

There is no CPU fallback: if the library is missing or a symbol does not resolve, the
import of a compute entry point raises.  The library is searched next to this file
(in-tree build, `python -m neilpy_b200.build`) or at $SMRF_B200_LIB.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SMRF_B200_LIB', os.path.join(HERE, 'libsmrf_b200.so'))

F32, F64 = 0, 1
PTS_SOA_F64, PTS_XYZW_F32, PTS_SOA_F32 = 0, 1, 2
BIN_MIN, BIN_MAX = 0, 1

_vp, _i64, _i32, _dbl, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_size_t
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

# name -> (restype, argtypes); mirrors include/smrf_b200.h declaration by declaration
SIGNATURES = {
    'smrf_abi_version': (_i32, []),
    'smrf_last_error': (C.c_char_p, []),
    'smrf_launch_count': (C.c_ulonglong, []),
    'smrf_open_variant': (C.c_char_p, [_i32, _i32]),
    'smrf_extent': (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    'smrf_bin_init': (_i32, [_vp, _i64, _i64, _i32, _i32, _vp]),
    'smrf_bin_accumulate': (_i32, [_vp, _vp, _vp, _i64, _i32, _dp, _vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    'smrf_bin_finalize': (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    'smrf_bin_finalize_partial': (_i32, [_vp, _i64, _i64, _i32, _i32, _vp]),
    'smrf_bin_mark_empty': (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    'smrf_route_plan': (_i32, [_vp, _vp, _i64, _i32, _dp, _i64, _i64, _i64, _i32, _vp, _vp, _vp]),
    'smrf_route_pack': (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'smrf_route_unpack': (_i32, [_vp, _vp, _i64, _vp, _vp]),
    'smrf_bin_accumulate_band': (_i32, [_vp, _vp, _vp, _i64, _i32, _dp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _vp, _vp]),
    'smrf_classify_band': (_i32, [_vp, _vp, _vp, _i64, _i32, _dp, _vp, _i64, _i64, _i64, _i64, _i32, _dbl, _dbl, _vp, _vp]),
    'smrf_inpaint_workspace_bytes': (_sz, [_i64, _i64]),
    'smrf_inpaint_fda_workspace_bytes': (_sz, [_i64, _i64]),
    'smrf_inpaint_fda': (_i32, [_vp, _i64, _i64, _i32, _vp, _sz, _dbl, _i32, _dp, _vp]),
    'smrf_inpaint_layout': (_i32, [_i64, _i64, C.POINTER(C.c_int64)]),
    'smrf_inpaint_setup': (_i32, [_vp, _i64, _i64, _i32, _vp, _sz, _i32, _i32, _vp]),
    'smrf_inpaint_start': (_i32, [_vp, _i64, _i64, _i32, _vp, _sz, _i32, _i32, _dbl, _vp, _i32, _vp, _vp, _vp]),
    'smrf_inpaint_step': (_i32, [_i64, _i64, _vp, _sz, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    'smrf_inpaint_compact': (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _sz, _i32, _i32, _i64, _i32, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'smrf_mg_cycle_part': (_i32, [_i64, _i64, _vp, _sz, _i32, _i32, _i32, _i32, _vp]),
    'smrf_mg_level_layout': (_i32, [_i64, _i64, _i32, C.POINTER(C.c_int64)]),
    'smrf_mg_setup_mask': (_i32, [_vp, _i64, _i64, _vp, _sz, _vp]),
    'smrf_mg_vcycle': (_i32, [_i64, _i64, _vp, _sz, _vp]),
    'smrf_mg_cycle_up_rz': (_i32, [_i64, _i64, _vp, _sz, _i32, _i32, _i32, _vp, _i64, _i64, _vp]),
    'smrf_inpaint_finish': (_i32, [_vp, _i64, _i64, _i32, _vp, _sz, _vp]),
    'smrf_inpaint': (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _sz, _dbl, _i32, _dp, _vp]),
    'smrf_open_workspace_bytes': (_sz, [_i64, _i64, _i32, _i32]),
    'smrf_progressive_open': (_i32, [_vp, _vp, _sz, _vp, _vp, _i64, _i64, _i32, _ip, _dp, _i32, _i32, _vp, _vp]),
    'smrf_open_window': (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _dbl, _i32, _i32, _i64, _i64, _vp]),
    'smrf_open_window_bruteforce': (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    'smrf_merge_punch': (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp]),
    'smrf_slope': (_i32, [_vp, _vp, _i64, _i64, _i32, _dbl, _vp]),
    'smrf_spline_workspace_bytes': (_sz, [_i64, _i64]),
    'smrf_spline_prefilter': (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    'smrf_classify': (_i32, [_vp, _vp, _vp, _i64, _i32, _dp, _vp, _vp, _i64, _i64, _i32, _dbl, _dbl,
                             _vp, _vp, _vp, _vp, _vp, _vp]),
    'smrf_las_decode': (_i32, [_vp, _i64, _i32, _dp, _dp, _vp, _vp, _vp, _vp, _i32, _vp]),
    'smrf_terrain': (_i32, [_vp, _i64, _i64, _i32, _i32, _i32, _dbl, _i32, _dbl, _dbl, _dbl, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp]),
    'smrf_las_write_class': (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i32, _vp]),
}

_lib = None


class SmrfLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library and bind every symbol the header declares."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SmrfLibraryError(
            'libsmrf_b200.so not found at %s -- build it with `python -m neilpy_b200.build` '
            '(there is no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise SmrfLibraryError('libsmrf_b200.so does not export %s' % name) from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc == 0:
        return
    msg = load().smrf_last_error().decode('utf-8', 'replace')
    if rc < 0:
        raise ValueError('%s: %s' % (what, msg))
    raise RuntimeError('%s failed with CUDA error %d: %s' % (what, rc, msg))
