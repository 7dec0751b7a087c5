"""Generates tests/golden/reference_exec.json: digests of what the REFERENCE's own source for the
SMRF path returns when it is executed in this container.

    python tests/golden/make_reference_exec_golden.py        (needs /root/reference)

`import neilpy` fails here (matplotlib, rasterio, skimage, ... are absent), but the five
functions on the path -- unique_rows, inpaint_nans_by_springs, create_dem, progressive_filter,
smrf (neilpy/neilpy.py:1110-1166, 1221-1271, 1659-1680, 1685-1808), plus inpaint_nans_by_fda (:1171-1216) -- only need numpy, pandas,
scipy and three third-party names.  Their source text is cut out of the reference module with
`ast`, compiled and executed UNMODIFIED, with exactly those three names supplied by the
oracle's restatements (and nothing else from the oracle):

    rasterio.transform.from_origin  -> oracle Affine6.from_origin
    skimage.morphology.disk         -> oracle disk
    skimage.morphology.opening      -> oracle opening (scipy.ndimage grey erosion + dilation)

So every line of the reference's own glue, arithmetic and call order is what produced these
digests; tests/test_oracle_golden.py holds the oracle's own functions to them bit for bit.
Only digests are stored, none of the reference's text.
"""
import ast
import hashlib
import json
import os
import sys

import numpy as np
import pandas as pd
import scipy.ndimage as ndi
import scipy.sparse.linalg  # noqa: F401  (the reference calls sparse.linalg.lsqr)
from scipy import interpolate, sparse, stats

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle import smrf_oracle as O  # noqa: E402

REF = '/root/reference/neilpy/neilpy.py'
NAMES = ('unique_rows', 'inpaint_nans_by_springs', 'inpaint_nans_by_fda', 'create_dem', 'progressive_filter', 'smrf')
NOTEBOOK = dict(cellsize=1, windows=18, slope_threshold=.15, elevation_threshold=.5, elevation_scaler=1.25)


class _Transform:
    from_origin = staticmethod(O.Affine6.from_origin)


class _Rasterio:
    transform = _Transform


def reference_functions():
    src = open(REF).read()
    ns = {'np': np, 'pd': pd, 'ndi': ndi, 'sparse': sparse, 'interpolate': interpolate, 'stats': stats,
          'rasterio': _Rasterio, 'disk': O.disk, 'opening': O.opening}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in NAMES:
            exec(compile(ast.get_source_segment(src, node), REF, 'exec'), ns)
    return ns


def digest(a):
    a = np.ascontiguousarray(np.asarray(a))
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    return '%s%s:%s' % (a.dtype.str, list(a.shape), hashlib.sha256(a.tobytes()).hexdigest()[:20])


def cases():
    """name -> (x, y, z, smrf kwargs)."""
    from neilpy_b200.synth import synth_cloud
    out = {}
    for s in ('samp11', 'samp12'):
        d = np.load(os.path.join(HERE, 'isprs_%s.npz' % s))
        out[s] = (d['x'] / 100.0, d['y'] / 100.0, d['z'] / 100.0, dict(NOTEBOOK))
    x, y, z, _ = synth_cloud(40000, 150.0, 120.0, seed=21)
    out['synth_w6'] = (x, y, z, dict(cellsize=1, windows=6, slope_threshold=.15, elevation_threshold=.5, elevation_scaler=1.25))
    out['synth_cs2_fill'] = (x, y, z, dict(cellsize=2, windows=np.array([1, 2, 4]), slope_threshold=.2, elevation_threshold=.4,
                                           elevation_scaler=1.0, low_outlier_fill=True))
    return out


def run(fns, x, y, z, kw):
    """Everything the path returns, through the functions in `fns` (reference or oracle)."""
    r = {}
    Zpro, t, oc, op, extras = fns['smrf'](x, y, z, return_extras=True, **kw)
    r['smrf.Zpro'], r['smrf.object_cells'], r['smrf.is_object_point'] = digest(Zpro), digest(oc), digest(op)
    r['smrf.t'] = [float(t[i]) for i in range(6)]
    for k in ('above_ground_height', 'drop_raster', 'when_dropped'):
        r['smrf.extras.' + k] = digest(extras[k])
    cs = kw['cellsize']
    Zmin, t1 = fns['create_dem'](x, y, z, cellsize=cs, bin_type='min')
    r['create_dem.min'] = digest(Zmin)
    Zmax, _ = fns['create_dem'](x, y, z, cellsize=cs, bin_type='max', inpaint=True)
    r['create_dem.max.inpaint'] = digest(Zmax)
    filled = fns['inpaint_nans_by_springs'](Zmin)
    r['inpaint_nans_by_springs'] = digest(filled)
    r['inpaint_nans_by_fda'] = digest(fns['inpaint_nans_by_fda'](Zmin))          # neilpy.py:1171-1216
    w = kw['windows']
    w = np.arange(w) + 1 if np.isscalar(w) else w
    mask, when = fns['progressive_filter'](filled, w, cs, kw['slope_threshold'], return_when_dropped=True)
    r['progressive_filter.mask'], r['progressive_filter.when'] = digest(mask), digest(when)
    return r


def oracle_functions():
    return {'smrf': O.smrf, 'create_dem': O.create_dem, 'inpaint_nans_by_springs': O.inpaint_nans_by_springs,
            'inpaint_nans_by_fda': O.inpaint_nans_by_fda, 'progressive_filter': O.progressive_filter}


def main():
    ref = reference_functions()
    out = {}
    for name, (x, y, z, kw) in cases().items():
        out[name] = run(ref, x, y, z, kw)
        mine = run(oracle_functions(), x, y, z, kw)
        diff = [k for k in out[name] if out[name][k] != mine[k]]
        print(name, 'oracle == reference source' if not diff else 'DIFFERS: %s' % diff)
    with open(os.path.join(HERE, 'reference_exec.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
