"""CPU oracle for the LAS ingest row -- TEST INFRASTRUCTURE ONLY.

Restates `read_las` of thomaspingel/neilpy (neilpy/neilpy.py:903-1087) in numpy/pandas:

    header fields, little-endian, fixed byte positions        neilpy.py:927-967
    LAZ (format id 128..133) and unknown formats raise         neilpy.py:954-964
    version 1.3: points end at `begin_wave_form` if non-zero   neilpy.py:972-977
    records = bytes [point_data_offset : end] viewed with the
      packed per-format record type                            neilpy.py:980-1053
    x = X*scale[0] + offset[0] (int32 -> float64, product
      rounded, then sum rounded), same for y, z                neilpy.py:1056-1059
    return / flag bytes split into their bit fields            neilpy.py:1061-1083

and adds `write_las`, a writer for small synthetic LAS files (the reference has none and
ships no LAS file), used only to build test inputs.

Parity pin: tests/golden/make_las_golden.py executes the reference's own `read_las`
(the function's source text, extracted from /root/reference/neilpy/neilpy.py at
generation time; it needs only struct, numpy and pandas) on synthetic files of all
eleven point formats and stores what it returned in tests/golden/las_golden.npz.
tests/test_las_oracle.py holds this restatement to those outputs exactly, so this
row is pinned to the reference itself.

The classification write-back follows the reference's laspy notebook
(examples/smrf/SMRF Classification using laspy to read and write.ipynb, cell 5:
`classification = 2*(1-is_object_point)`), with laspy's rule that the flag bits
sharing the byte in formats 0-5 are kept.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this
module.  The product (neilpy_b200) never does.
"""
from __future__ import annotations

import struct

import numpy as np
import pandas as pd

# blocks of the LAS 1.4 point record, in file order
_HEAD = [('x', '<i4'), ('y', '<i4'), ('z', '<i4'), ('intensity', '<u2'), ('return_byte', 'u1')]
_LEGACY = [('class', 'u1'), ('scan_angle', 'u1'), ('user_data', 'u1'), ('point_source_id', '<u2')]
_MODERN = [('mixed_byte', 'u1'), ('class', 'u1'), ('user_data', 'u1'), ('scan_angle', '<u2'),
           ('point_source_id', '<u2'), ('gpstime', '<f8')]
_GPS = [('gpstime', '<f8')]
_RGB = [('red', '<u2'), ('green', '<u2'), ('blue', '<u2')]
_NIR = [('near_infrared', '<u2')]
_WAVE = [('wave_packet_descriptor_index', 'u1'), ('byte_offset', '<u8'), ('wave_packet_size', '<u4'),
         ('return_point_waveform_location', '<f4'), ('xt', '<f4'), ('yt', '<f4'), ('zt', '<f4')]

_LAYOUT = {0: _HEAD + _LEGACY,
           1: _HEAD + _LEGACY + _GPS,
           2: _HEAD + _LEGACY + _RGB,
           3: _HEAD + _LEGACY + _GPS + _RGB,
           4: _HEAD + _LEGACY + _GPS + _WAVE,
           5: _HEAD + _LEGACY + _GPS + _RGB + _WAVE,
           6: _HEAD + _MODERN,
           7: _HEAD + _MODERN + _RGB,
           8: _HEAD + _MODERN + _RGB + _NIR,
           9: _HEAD + _MODERN + _WAVE,
           10: _HEAD + _MODERN + _RGB + _NIR + _WAVE}


def record_dtype(fmt):
    return np.dtype(_LAYOUT[fmt])          # packed, no alignment padding


def _bit(v, i):
    return (v & (1 << i)) != 0


def read_las(filename_or_bytes):
    """neilpy.py:903-1087.  Returns (header dict, DataFrame)."""
    if isinstance(filename_or_bytes, (bytes, bytearray, memoryview)):
        data = bytes(filename_or_bytes)
    else:
        with open(filename_or_bytes, 'rb') as f:
            data = f.read()
    u = lambda fmt, a, b: struct.unpack(fmt, data[a:b])          # noqa: E731
    h = {}
    h['file_signature'] = u('<4s', 0, 4)[0].decode('utf-8')
    h['file_source_id'] = u('<H', 4, 6)[0]
    h['global_encoding'] = u('<H', 6, 8)[0]
    h['project_id'] = [u('<L', 8, 12)[0], u('<H', 12, 14)[0], u('<H', 14, 16)[0]]
    h['version_major'] = u('<B', 24, 25)[0]
    h['version_minor'] = u('<B', 25, 26)[0]
    h['version'] = h['version_major'] + h['version_minor'] / 10
    h['system_id'] = u('32s', 26, 58)[0].decode('utf-8').rstrip('\x00')
    h['generating_software'] = u('32s', 58, 90)[0].decode('utf-8').rstrip('\x00')
    h['file_creation_day'] = u('<H', 90, 92)[0]
    h['file_creation_year'] = u('<H', 92, 94)[0]
    h['header_size'] = u('<H', 94, 96)[0]
    h['point_data_offset'] = u('<L', 96, 100)[0]
    h['num_variable_records'] = u('<L', 100, 104)[0]
    fmt = u('<B', 104, 105)[0]
    if 128 <= fmt <= 133:
        raise ValueError('LAZ not yet supported.')
    h['point_data_format_id'] = fmt
    if fmt not in _LAYOUT:
        raise ValueError('Point Data Record Format', fmt, 'not yet supported.')
    h['point_data_record_length'] = u('<H', 105, 107)[0]
    h['num_point_records'] = u('<L', 107, 111)[0]
    h['num_points_by_return'] = u('<5L', 111, 131)
    h['scale'] = u('<3d', 131, 155)
    h['offset'] = u('<3d', 155, 179)
    h['minmax'] = u('<6d', 179, 227)
    end = len(data)
    if h['version'] == 1.3:
        h['begin_wave_form'] = u('<q', 227, 235)[0]
        if h['begin_wave_form'] != 0:
            end = h['begin_wave_form']
    rec = np.frombuffer(data[h['point_data_offset']:end], record_dtype(fmt))
    df = pd.DataFrame(rec)
    for k, name in enumerate('xyz'):
        df[name] = df[name] * h['scale'][k] + h['offset'][k]
    rb = df['return_byte']
    if fmt < 6:
        df['return_number'] = 4 * _bit(rb, 2).astype(np.uint8) + 2 * _bit(rb, 1).astype(np.uint8) + _bit(rb, 0).astype(np.uint8)
        df['return_max'] = 4 * _bit(rb, 5).astype(np.uint8) + 2 * _bit(rb, 4).astype(np.uint8) + _bit(rb, 3).astype(np.uint8)
        df['scan_direction'] = _bit(rb, 6)
        df['edge_of_flight_line'] = _bit(rb, 7)
        del df['return_byte']
    else:
        df['return_number'] = (8 * _bit(rb, 3).astype(np.uint8) + 4 * _bit(rb, 2).astype(np.uint8)
                               + 2 * _bit(rb, 1).astype(np.uint8) + _bit(rb, 0).astype(np.uint8))
        df['return_max'] = (8 * _bit(rb, 7).astype(np.uint8) + 4 * _bit(rb, 6).astype(np.uint8)
                            + 2 * _bit(rb, 5).astype(np.uint8) + _bit(rb, 4).astype(np.uint8))
        del df['return_byte']
        mb = df['mixed_byte']
        df['classification_bit_synthetic'] = _bit(mb, 0)
        df['classification_bit_keypoint'] = _bit(mb, 1)
        df['classification_bit_withheld'] = _bit(mb, 2)
        df['classification_bit_overlap'] = _bit(mb, 3)
        df['scanner_channel'] = 2 * _bit(mb, 5).astype(np.uint8) + 1 * _bit(mb, 4).astype(np.uint8)
        df['scan_direction'] = _bit(mb, 6)
        df['edge_of_flight_line'] = _bit(mb, 7)
        del df['mixed_byte']
    return h, df


def ground_classification(records_bytes, fmt, is_object_point):
    """Record bytes with the classification set to 2*(1-is_object_point) (the laspy notebook,
    cell 5); formats 0-5 keep the three flag bits that share the byte."""
    rec = np.frombuffer(bytes(records_bytes), record_dtype(fmt)).copy()
    code = (2 * (1 - np.asarray(is_object_point).astype(np.int64))).astype(np.uint8)
    rec['class'] = (rec['class'] & 0xE0) | code if fmt < 6 else code
    return rec.tobytes()


# ------------------------------------------------------------------ synthetic LAS writer
def synth_records(fmt, n, seed=0):
    """n random records of point format `fmt`: every byte of every field is exercised."""
    rng = np.random.default_rng(seed)
    dt = record_dtype(fmt)
    raw = rng.integers(0, 256, size=(n, dt.itemsize), dtype=np.uint8)
    rec = np.frombuffer(raw.tobytes(), dt).copy()
    for name in ('x', 'y', 'z'):                         # full int32 range, both signs, the extremes too
        v = rng.integers(-2 ** 31, 2 ** 31, size=n, dtype=np.int64)
        if n >= 2:
            v[0], v[-1] = -2 ** 31, 2 ** 31 - 1
        rec[name] = v.astype(np.int32)
    for name in dt.names:                                # NaN payload bits do not survive a DataFrame round trip
        if dt[name].kind == 'f':
            rec[name] = rng.standard_normal(n).astype(dt[name])
    return rec


def write_las(records, fmt, scale=(0.01, 0.01, 0.001), offset=(500000.0, 5400000.0, -12.5), version=(1, 2),
              vlr_bytes=0, trailing_bytes=0, begin_wave_form=None):
    """A LAS file image (bytes) around `records` (structured array of `record_dtype(fmt)`)."""
    records = np.ascontiguousarray(records)
    n = len(records)
    header_size = 235 if version == (1, 3) else (375 if version == (1, 4) else 227)
    point_offset = header_size + vlr_bytes
    hd = bytearray(header_size)
    hd[0:4] = b'LASF'
    struct.pack_into('<HH', hd, 4, 7, 1)
    struct.pack_into('<LHH', hd, 8, 0xDEADBEEF, 0x1234, 0x5678)
    hd[24], hd[25] = version
    hd[26:26 + 10] = b'neilpyb200'
    hd[58:58 + 11] = b'las_oracle '
    struct.pack_into('<HHHLLBHL', hd, 90, 291, 2026, header_size, point_offset, 1 if vlr_bytes else 0, fmt,
                     records.dtype.itemsize, n)
    struct.pack_into('<5L', hd, 111, n, 0, 0, 0, 0)
    struct.pack_into('<3d', hd, 131, *scale)
    struct.pack_into('<3d', hd, 155, *offset)
    struct.pack_into('<6d', hd, 179, 6.0, 1.0, 5.0, 2.0, 4.0, 3.0)
    body = records.tobytes()
    if version == (1, 3):
        bw = (point_offset + len(body)) if (begin_wave_form is None and trailing_bytes) else (begin_wave_form or 0)
        struct.pack_into('<q', hd, 227, bw)
    rng = np.random.default_rng(99)
    vlr = rng.integers(0, 256, vlr_bytes, dtype=np.uint8).tobytes()
    tail = rng.integers(0, 256, trailing_bytes, dtype=np.uint8).tobytes()
    return bytes(hd) + vlr + body + tail
