"""The C-ABI shared library loads and exports every symbol include/smrf_b200.h declares
(no compute call is made: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'smrf_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(smrf_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_hot_path():
    names = declared_symbols()
    for need in ('smrf_extent', 'smrf_bin_init', 'smrf_bin_accumulate', 'smrf_bin_finalize', 'smrf_inpaint',
                 'smrf_progressive_open', 'smrf_open_window', 'smrf_merge_punch', 'smrf_slope',
                 'smrf_spline_prefilter', 'smrf_classify'):
        assert need in names


def test_library_exports_every_declared_symbol():
    from neilpy_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert set(declared_symbols()) == set(_lib.SIGNATURES), 'ctypes table out of sync with the header'
    bound = _lib.load()
    assert bound.smrf_abi_version() == 1
    assert bound.smrf_open_variant(_lib.F32, 18).decode().startswith('march')
    assert bound.smrf_open_variant(_lib.F64, 18).decode() == 'direct_generic'


def test_missing_library_fails_loudly(monkeypatch):
    from neilpy_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libsmrf_b200.so')
    with pytest.raises(_lib.SmrfLibraryError):
        _lib.load()


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import numpy as np
    import neilpy_b200
    with pytest.raises(RuntimeError):
        neilpy_b200.progressive_filter(np.zeros((8, 8)), np.array([1]))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'neilpy_b200')
    pat = re.compile(r'^\s*(from|import)\s+oracle|smrf_oracle|oracle/', re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/smrf_b200.h is a C header (no C++ or torch types in the signatures): a C99 program
    includes it, links the library and calls an entry point that needs no GPU."""
    import shutil
    import subprocess
    from neilpy_b200 import _lib
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    src = tmp_path / 'abi.c'
    src.write_text('#include "smrf_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%d %s\\n", smrf_abi_version(), smrf_open_variant(SMRF_F32, 18));'
                   ' return smrf_abi_version() == 1 ? 0 : 1; }\n')
    exe = tmp_path / 'abi'
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', os.path.join(ROOT, 'include'),
                    str(src), '-L', libdir, '-lsmrf_b200', '-Wl,-rpath,' + libdir, '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == '1' and out[1].startswith('march')
