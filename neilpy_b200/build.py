"""In-tree build of libsmrf_b200.so (hand-written CUDA for sm_100a, C ABI in include/smrf_b200.h).

    python -m neilpy_b200.build [--force] [--jobs N]

Every translation unit is compiled with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17
(nvcc cross-compiles without a GPU).  The register-marching opening kernels are fully
unrolled per disk radius, so opening_march_inst.cu is compiled once per radius and the
units build in parallel.  Objects land in neilpy_b200/_build/ (git-ignored); the shared
library lands next to this file so that it travels with the source tree.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, '_build')
LIB = os.path.join(HERE, 'libsmrf_b200.so')
MARCH_MAX_W = 72

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC',
              '-DSMRF_MARCH_MAX_W=%d' % MARCH_MAX_W]


def _units():
    units = []
    for name in ('rank', 'route', 'fda', 'binning', 'grid_ops', 'inpaint', 'spline', 'las', 'terrain', 'opening_generic', 'opening_march'):
        units.append((name + '.o', name + '.cu', []))
    for w in range(MARCH_MAX_W, 0, -1):   # slowest first
        units.append(('opening_march_w%02d.o' % w, 'opening_march_inst.cu', ['-DSMRF_W=%d' % w]))
    return units


def _newest_header():
    t = os.path.getmtime(os.path.join(HERE, '..', 'include', 'smrf_b200.h'))
    for f in os.listdir(CSRC):
        if f.endswith('.cuh'):
            t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    return max(t, os.path.getmtime(__file__))


def _compile(unit, force, hdr_time):
    obj, src, extra = unit
    objp, srcp = os.path.join(OBJ, obj), os.path.join(CSRC, src)
    if (not force and os.path.exists(objp)
            and os.path.getmtime(objp) >= max(os.path.getmtime(srcp), hdr_time)):
        return obj, 0.0, 'cached'
    t0 = time.time()
    cmd = ['nvcc'] + NVCC_FLAGS + extra + ['-c', srcp, '-o', objp]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (obj, ' '.join(cmd), r.stdout + r.stderr))
    return obj, time.time() - t0, 'built'


def build(force=False, jobs=None, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    units = _units()
    hdr_time = _newest_header()
    jobs = jobs or max(1, (os.cpu_count() or 4))
    t0 = time.time()
    rebuilt = False
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        for obj, dt, what in ex.map(lambda u: _compile(u, force, hdr_time), units):
            rebuilt |= what == 'built'
            if verbose and what == 'built':
                print('  [nvcc] %-28s %6.1fs' % (obj, dt), flush=True)
    objs = [os.path.join(OBJ, u[0]) for u in units]
    if rebuilt or force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = ['nvcc', '-shared', '-cudart', 'static', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s' % (r.stdout + r.stderr))
        if verbose:
            print('  [link] %s (%.1fs total)' % (LIB, time.time() - t0), flush=True)
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--jobs', type=int, default=None)
    a = ap.parse_args()
    build(force=a.force, jobs=a.jobs)
    sys.exit(0)
