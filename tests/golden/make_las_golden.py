"""Generates tests/golden/las_golden.npz: what the REFERENCE's own `read_las` returns on
synthetic LAS files of all eleven point formats.

Run in the build container (needs /root/reference, which the GPU box lacks):

    python tests/golden/make_las_golden.py

`neilpy` cannot be imported here (matplotlib, rasterio, ... are absent), but `read_las`
(neilpy/neilpy.py:903-1087) needs only struct, numpy and pandas: its source text is cut out
of the reference module with `ast` at generation time, compiled and executed unmodified.
Nothing of it is stored in this repository -- only the file images it was given (written by
oracle/las_oracle.write_las) and the columns / header values it returned.
"""
import ast
import contextlib
import io
import json
import os
import struct
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle import las_oracle as L  # noqa: E402

REF = '/root/reference/neilpy/neilpy.py'

# (name, format, n, version, vlr bytes, trailing bytes, scale, offset)
CASES = [('f%d' % f, f, 257 + 3 * f, (1, 2), 0, 0, (0.01, 0.01, 0.001), (500000.0, 5400000.0, -12.5)) for f in range(11)]
CASES += [('f1_vlr', 1, 1000, (1, 2), 54 + 77, 0, (0.001, 0.001, 0.001), (0.0, 0.0, 0.0)),
          ('f3_v13_wave', 3, 300, (1, 3), 31, 123, (0.01, 0.02, 0.03), (-1e6, 1e7, 0.1)),
          ('f3_v13_nowave', 3, 64, (1, 3), 0, 0, (0.01, 0.01, 0.01), (1.0, 2.0, 3.0)),
          ('f6_v14', 6, 513, (1, 4), 13, 0, (1e-3, 1e-3, 1e-4), (123456.789, 9876543.21, 1000.0)),
          ('f0_one', 0, 1, (1, 2), 0, 0, (0.01, 0.01, 0.01), (0.0, 0.0, 0.0)),
          ('f2_empty', 2, 0, (1, 2), 0, 0, (0.01, 0.01, 0.01), (0.0, 0.0, 0.0))]


def reference_read_las():
    src = open(REF).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'read_las')
    text = ast.get_source_segment(src, node)
    ns = {'struct': struct, 'np': np, 'pd': pd}
    exec(compile(text, REF, 'exec'), ns)
    return ns['read_las']


def main():
    read_las = reference_read_las()
    out, meta = {}, {}
    for name, fmt, n, version, vlr, trailing, scale, offset in CASES:
        rec = L.synth_records(fmt, n, seed=100 + len(out))
        image = L.write_las(rec, fmt, scale=scale, offset=offset, version=version, vlr_bytes=vlr, trailing_bytes=trailing)
        with tempfile.NamedTemporaryFile(suffix='.las', delete=False) as f:
            f.write(image)
            path = f.name
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                header, df = read_las(path)
        finally:
            os.unlink(path)
        out[name + '__file'] = np.frombuffer(image, np.uint8)
        for c in df.columns:
            out[name + '__col__' + c] = df[c].to_numpy()
        meta[name] = {'format': fmt, 'n': n, 'columns': list(df.columns), 'dtypes': [str(df[c].dtype) for c in df.columns],
                      'header': {k: (list(v) if isinstance(v, (tuple, list)) else v) for k, v in header.items()}}
    out['meta'] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, 'las_golden.npz'), **out)
    print('wrote las_golden.npz:', len(CASES), 'cases,', os.path.getsize(os.path.join(HERE, 'las_golden.npz')), 'bytes')


if __name__ == '__main__':
    main()
