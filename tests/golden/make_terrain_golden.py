"""Generates tests/golden/terrain_golden.npz: what the REFERENCE's own slope / aspect /
hillshade / pssm return on small synthetic DEMs, and the colours of the reference's shipped
bonemap image.

Run in the build container (needs /root/reference):

    python tests/golden/make_terrain_golden.py

The four functions (neilpy/neilpy.py:456-484, 814-824, 846-867) use numpy only -- pssm also
matplotlib's bone colour map, which is absent here and is supplied by the oracle's restated
table -- so their source text is cut out of the reference module with `ast`, compiled and
executed unmodified.  Nothing of it is stored here, only inputs and outputs.
"""
import ast
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle import terrain_oracle as T  # noqa: E402

REF = '/root/reference/neilpy/neilpy.py'
PNG = '/root/reference/examples/dk22_smrfed_bonemap.png'

CALLS = [('slope', dict()), ('slope', dict(cellsize=2, return_as='radians')), ('slope', dict(cellsize=.5, z_factor=3, return_as='percent')),
         ('aspect', dict()), ('aspect', dict(return_as='radians', flat_as=0)), ('aspect', dict(flat_as=-1)),
         ('hillshade', dict()), ('hillshade', dict(cellsize=2, z_factor=1.5, zenith=30, azimuth=135)),
         ('hillshade', dict(cellsize=5, return_uint8=False)),
         ('pssm', dict(apply_colormap=False)), ('pssm', dict(cellsize=5, ve=1.0, apply_colormap=False)),
         ('pssm', dict(cellsize=2)), ('pssm', dict(cellsize=2, reverse=True))]
DEMS = {'a': (37, 53, 1), 'b': (64, 40, 2), 'c': (5, 6, 3), 'd': (2, 2, 4)}


def reference_functions():
    src = open(REF).read()
    ns = {'np': np, 'plt': T.plt_stub}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ('slope', 'aspect', 'hillshade', 'pssm'):
            exec(compile(ast.get_source_segment(src, node), REF, 'exec'), ns)
    return ns


def main():
    ns = reference_functions()
    out, index = {}, []
    for dname, (ny, nx, seed) in DEMS.items():
        Z = T.synth_dem(ny, nx, seed)
        out['dem_' + dname] = Z
        for k, (fn, kw) in enumerate(CALLS):
            key = '%s_%s_%d' % (dname, fn, k)
            out[key] = ns[fn](Z.copy(), **kw)
            index.append({'key': key, 'dem': dname, 'fn': fn, 'kwargs': kw})
    out['index'] = np.frombuffer(json.dumps(index).encode(), np.uint8)
    from PIL import Image
    rgba = np.array(Image.open(PNG))
    out['bonemap_png_colours'] = np.unique(rgba.reshape(-1, 4), axis=0)
    np.savez_compressed(os.path.join(HERE, 'terrain_golden.npz'), **out)
    print('wrote terrain_golden.npz', len(index), 'results,', os.path.getsize(os.path.join(HERE, 'terrain_golden.npz')), 'bytes')


if __name__ == '__main__':
    main()
