"""LAS ingest / egress around the SMRF path.

    read_las(filename)            drop-in for neilpy.read_las (neilpy/neilpy.py:903-1087):
                                  (header dict, pandas DataFrame), same keys, columns, dtypes
    read_las_device(filename)     header + LasPoints: x, y, z (float64) and the classification
                                  byte as CUDA tensors, and the record block itself on the device;
                                  `smrf(pts.x, pts.y, pts.z, ...)` then never touches the host
    write_classification(...)     ground = 2, object = 0 into the records on the device (what the
                                  reference's laspy notebook does with the result of smrf)
    save_las(...)                 the input file image with the updated records
    classify_las(src, dst, ...)   read -> smrf -> write-back -> save

The coordinate decode (int32 * scale + offset in float64) and the classification write-back
run in libsmrf_b200.so (csrc/las.cu: smrf_las_decode, smrf_las_write_class); there is no CPU
fallback for them.  What stays on the host is the 227-byte header (struct parsing) and, for
the DataFrame of `read_las` only, the attribute columns, which are reinterpretations of the
record bytes and single-bit tests with no arithmetic on coordinates.

Deliberately the reference's behaviour, including its limits: the record size is the
format's minimum size (the header's point_data_record_length is reported but not used, so
files with extra bytes per record raise, as np.frombuffer does in the reference); LAZ raises
ValueError('LAZ not yet supported.'); LAS 1.3 files stop at begin_wave_form.  The reference
also prints a notice for formats 6-10; this module does not print.
"""
from __future__ import annotations

import ctypes as C
import struct
from collections import OrderedDict

import numpy as np
import torch

from . import _lib

# minimum record size per point data record format (LAS 1.4 R15, table per format;
# the reference keeps the same table at neilpy.py:925)
RECORD_LENGTH = {0: 20, 1: 28, 2: 26, 3: 34, 4: 57, 5: 63, 6: 30, 7: 36, 8: 38, 9: 59, 10: 67}

# (name, numpy type) blocks of the record after X, Y, Z, intensity and the return byte
_TAIL_LEGACY = (('class', 'u1'), ('scan_angle', 'u1'), ('user_data', 'u1'), ('point_source_id', '<u2'))
_TAIL_MODERN = (('mixed_byte', 'u1'), ('class', 'u1'), ('user_data', 'u1'), ('scan_angle', '<u2'),
                ('point_source_id', '<u2'), ('gpstime', '<f8'))
_OPTIONAL = {'gps': (('gpstime', '<f8'),),
             'rgb': (('red', '<u2'), ('green', '<u2'), ('blue', '<u2')),
             'nir': (('near_infrared', '<u2'),),
             'wave': (('wave_packet_descriptor_index', 'u1'), ('byte_offset', '<u8'), ('wave_packet_size', '<u4'),
                      ('return_point_waveform_location', '<f4'), ('xt', '<f4'), ('yt', '<f4'), ('zt', '<f4'))}
_EXTRA_BLOCKS = {0: (), 1: ('gps',), 2: ('rgb',), 3: ('gps', 'rgb'), 4: ('gps', 'wave'), 5: ('gps', 'rgb', 'wave'),
                 6: (), 7: ('rgb',), 8: ('rgb', 'nir'), 9: ('wave',), 10: ('rgb', 'nir', 'wave')}


def record_fields(fmt):
    """[(name, numpy type, byte offset)] of point format `fmt`, in file order."""
    fields = [('x', '<i4'), ('y', '<i4'), ('z', '<i4'), ('intensity', '<u2'), ('return_byte', 'u1')]
    fields += list(_TAIL_LEGACY if fmt < 6 else _TAIL_MODERN)
    for blk in _EXTRA_BLOCKS[fmt]:
        fields += list(_OPTIONAL[blk])
    out, off = [], 0
    for name, t in fields:
        out.append((name, t, off))
        off += np.dtype(t).itemsize
    assert off == RECORD_LENGTH[fmt]
    return out


def class_offset(fmt):
    return next(off for name, _, off in record_fields(fmt) if name == 'class')


def parse_header(data):
    """The public header block as the reference's dict (neilpy.py:927-977)."""
    if len(data) < 227:
        raise struct.error('LAS header needs 227 bytes')
    h = {}
    sig, source_id, encoding = struct.unpack_from('<4sHH', data, 0)
    h['file_signature'] = sig.decode('utf-8')
    h['file_source_id'], h['global_encoding'] = source_id, encoding
    h['project_id'] = list(struct.unpack_from('<LHH', data, 8))
    h['version_major'], h['version_minor'] = struct.unpack_from('<BB', data, 24)
    h['version'] = h['version_major'] + h['version_minor'] / 10
    sysid, soft = struct.unpack_from('<32s32s', data, 26)
    h['system_id'] = sysid.decode('utf-8').rstrip('\x00')
    h['generating_software'] = soft.decode('utf-8').rstrip('\x00')
    (h['file_creation_day'], h['file_creation_year'], h['header_size'], h['point_data_offset'],
     h['num_variable_records'], fmt) = struct.unpack_from('<HHHLLB', data, 90)
    if 128 <= fmt <= 133:
        raise ValueError('LAZ not yet supported.')
    h['point_data_format_id'] = fmt
    if fmt not in RECORD_LENGTH:
        raise ValueError('Point Data Record Format', fmt, 'not yet supported.')
    h['point_data_record_length'], h['num_point_records'] = struct.unpack_from('<HL', data, 105)
    h['num_points_by_return'] = struct.unpack_from('<5L', data, 111)
    h['scale'] = struct.unpack_from('<3d', data, 131)
    h['offset'] = struct.unpack_from('<3d', data, 155)
    h['minmax'] = struct.unpack_from('<6d', data, 179)
    if h['version'] == 1.3:
        h['begin_wave_form'] = struct.unpack_from('<q', data, 227)[0]
    return h


def point_block(header, file_size):
    """(first byte, end byte, record length) of the point records in the file image
    (neilpy.py:968-980; the size check is np.frombuffer's at :1055)."""
    end = file_size
    if header.get('begin_wave_form', 0) != 0:
        end = header['begin_wave_form']
    lo = header['point_data_offset']
    length = RECORD_LENGTH[header['point_data_format_id']]
    nbytes = max(0, end - lo)
    if nbytes % length:
        raise ValueError('buffer size must be a multiple of element size')
    return lo, lo + nbytes, length


def _bit(v, i):
    return (v & np.uint8(1 << i)) != 0


def attribute_columns(block, fmt):
    """Every column of the reference's DataFrame except x, y, z, in its order
    (neilpy.py:980-1083): strided views of the record bytes, then the bit fields."""
    length = RECORD_LENGTH[fmt]
    block = np.ascontiguousarray(block, dtype=np.uint8)
    n = block.size // length
    rows = block.reshape(n, length)
    raw = OrderedDict()
    for name, t, off in record_fields(fmt):
        if name in ('x', 'y', 'z'):
            continue
        size = np.dtype(t).itemsize
        raw[name] = np.ascontiguousarray(rows[:, off:off + size]).view(t).reshape(n)
    rb = raw.pop('return_byte')
    u8 = lambda b: b.astype(np.uint8)      # noqa: E731
    if fmt < 6:
        raw['return_number'] = u8(rb & 7)
        raw['return_max'] = u8((rb >> 3) & 7)
        raw['scan_direction'] = _bit(rb, 6)
        raw['edge_of_flight_line'] = _bit(rb, 7)
    else:
        mb = raw.pop('mixed_byte')
        raw['return_number'] = u8(rb & 15)
        raw['return_max'] = u8(rb >> 4)
        raw['classification_bit_synthetic'] = _bit(mb, 0)
        raw['classification_bit_keypoint'] = _bit(mb, 1)
        raw['classification_bit_withheld'] = _bit(mb, 2)
        raw['classification_bit_overlap'] = _bit(mb, 3)
        raw['scanner_channel'] = u8((mb >> 4) & 3)
        raw['scan_direction'] = _bit(mb, 6)
        raw['edge_of_flight_line'] = _bit(mb, 7)
    return raw


class LasPoints:
    """Device-resident points of one LAS file."""

    def __init__(self, header, records, x, y, z, classification, image, lo, hi):
        self.header, self.records = header, records
        self.x, self.y, self.z, self.classification = x, y, z, classification
        self._image, self._lo, self._hi = image, lo, hi

    def __len__(self):
        return int(self.x.numel())

    @property
    def format(self):
        return self.header['point_data_format_id']


def _read_image(filename_or_bytes):
    if isinstance(filename_or_bytes, (bytes, bytearray, memoryview, np.ndarray)):
        return np.frombuffer(filename_or_bytes, dtype=np.uint8)
    return np.fromfile(filename_or_bytes, dtype=np.uint8)


def decode_records(records, n, fmt, scale, offset, want_class=True):
    """records: uint8 CUDA tensor holding n packed records.  Returns x, y, z (float64) and the
    classification byte (or None) as CUDA tensors.  smrf_las_decode."""
    from .api import _ptr, _stream
    lib = _lib.load()
    dev = records.device
    x, y, z = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
    cls = torch.empty(n, dtype=torch.uint8, device=dev) if want_class else None
    sc = (C.c_double * 3)(*[float(v) for v in scale])
    of = (C.c_double * 3)(*[float(v) for v in offset])
    _lib.check(lib.smrf_las_decode(_ptr(records), n, RECORD_LENGTH[fmt], sc, of, _ptr(x), _ptr(y), _ptr(z), _ptr(cls),
                                   class_offset(fmt), _stream()), 'smrf_las_decode')
    return x, y, z, cls


def read_las_device(filename_or_bytes, want_class=True):
    """(header, LasPoints).  The record block goes to the device as it lies in the file (one
    pinned staging copy) and is decoded there."""
    from .api import _device
    dev = _device()
    image = _read_image(filename_or_bytes)
    header = parse_header(image[:235].tobytes())
    lo, hi, length = point_block(header, image.size)
    n = (hi - lo) // length
    records = torch.empty(max(hi - lo, 1), dtype=torch.uint8, device=dev)     # cudaMalloc'd: 256-byte aligned
    if n:
        staged = torch.empty(hi - lo, dtype=torch.uint8, pin_memory=True)
        staged.numpy()[:] = image[lo:hi]
        records[:hi - lo].copy_(staged, non_blocking=True)
    x, y, z, cls = decode_records(records, n, header['point_data_format_id'], header['scale'], header['offset'], want_class)
    torch.cuda.current_stream().synchronize()                                 # `staged` may go away now
    return header, LasPoints(header, records[:hi - lo], x, y, z, cls, image, lo, hi)


def read_las(filename):
    """Drop-in for neilpy.read_las: (header, DataFrame)."""
    import pandas as pd
    header, pts = read_las_device(filename, want_class=False)
    fmt = header['point_data_format_id']
    cols = OrderedDict()
    cols['x'], cols['y'], cols['z'] = (t.cpu().numpy() for t in (pts.x, pts.y, pts.z))
    cols.update(attribute_columns(pts._image[pts._lo:pts._hi], fmt))     # already in the reference's column order
    return header, pd.DataFrame(cols)


def write_classification(pts, is_object_point, ground_code=2, object_code=0):
    """records[i].classification = object_code if is_object_point[i] else ground_code, on the
    device (smrf_las_write_class); formats 0-5 keep the flag bits sharing the byte."""
    from .api import _ptr, _stream
    lib = _lib.load()
    obj = torch.as_tensor(np.asarray(is_object_point) if not isinstance(is_object_point, torch.Tensor) else is_object_point)
    obj = obj.to(device=pts.records.device).view(-1)
    obj = obj.view(torch.uint8) if obj.dtype == torch.bool else (obj != 0).view(torch.uint8)
    if obj.numel() != len(pts):
        raise ValueError('is_object_point must have one entry per point')
    fmt = pts.format
    _lib.check(lib.smrf_las_write_class(_ptr(pts.records), len(pts), RECORD_LENGTH[fmt], class_offset(fmt),
                                        0xE0 if fmt < 6 else 0, _ptr(obj.contiguous()), int(ground_code), int(object_code),
                                        _stream()), 'smrf_las_write_class')
    return pts


def save_las(filename, pts):
    """The file image `pts` was read from, with the device's (possibly re-classified) records."""
    out = pts._image.copy()
    out[pts._lo:pts._hi] = pts.records.cpu().numpy()
    if filename is None:
        return out.tobytes()
    out.tofile(filename)
    return None


def classify_las(src, dst, **smrf_kwargs):
    """read_las_device -> smrf -> write_classification -> save_las.  Returns smrf's tuple
    (device tensors) so that the DTM can be written out too."""
    from .api import smrf
    header, pts = read_las_device(src)
    res = smrf(pts.x, pts.y, pts.z, **smrf_kwargs)
    write_classification(pts, res[3])
    save_las(dst, pts)
    return res
