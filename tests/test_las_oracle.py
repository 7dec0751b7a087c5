"""The LAS oracle (oracle/las_oracle.py) against what the reference's own read_las returned
(tests/golden/las_golden.npz, made by tests/golden/make_las_golden.py), and the host-side
LAS logic of neilpy_b200.las that needs no GPU."""
import json
import os

import numpy as np
import pytest

from oracle import las_oracle as L

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden():
    g = np.load(os.path.join(HERE, 'golden', 'las_golden.npz'))
    meta = json.loads(bytes(g['meta']).decode())
    return g, meta


GOLD, META = load_golden()


def header_equal(a, b):
    if set(a) != set(b):
        return False
    for k in a:
        va = list(a[k]) if isinstance(a[k], (tuple, list)) else a[k]
        vb = list(b[k]) if isinstance(b[k], (tuple, list)) else b[k]
        if va != vb:
            return False
    return True


@pytest.mark.parametrize('name', sorted(META))
def test_oracle_matches_reference_read_las(name):
    m = META[name]
    header, df = L.read_las(bytes(GOLD[name + '__file']))
    assert header_equal(header, m['header'])
    assert list(df.columns) == m['columns']
    assert [str(df[c].dtype) for c in df.columns] == m['dtypes']
    assert len(df) == m['n']
    for c in df.columns:
        assert np.array_equal(df[c].to_numpy(), GOLD[name + '__col__' + c]), c


def test_oracle_errors_like_reference():
    rec = L.synth_records(1, 10)
    img = bytearray(L.write_las(rec, 1))
    img[104] = 129                                   # LAZ-compressed format id
    with pytest.raises(ValueError, match='LAZ'):
        L.read_las(bytes(img))
    img[104] = 11
    with pytest.raises(ValueError):
        L.read_las(bytes(img))
    img[104] = 1
    with pytest.raises(ValueError):                  # np.frombuffer: size not a multiple of the record
        L.read_las(bytes(img) + b'\x00')


def test_host_header_and_columns_match_golden():
    """neilpy_b200.las: header parse and the byte/bit columns are host logic (no arithmetic on
    coordinates); they must agree with the reference for every format."""
    from neilpy_b200 import las
    for name, m in META.items():
        img = bytes(GOLD[name + '__file'])
        header = las.parse_header(img)
        assert header_equal(header, m['header']), name
        lo, hi, length = las.point_block(header, len(img))
        assert (hi - lo) == m['n'] * length
        cols = las.attribute_columns(np.frombuffer(img, np.uint8)[lo:hi], header['point_data_format_id'])
        want = [c for c in m['columns'] if c not in 'xyz']
        assert list(cols) == want, name
        for c in want:
            got = cols[c]
            assert str(got.dtype) == m['dtypes'][m['columns'].index(c)], (name, c)
            assert np.array_equal(got, GOLD[name + '__col__' + c]), (name, c)


def test_host_errors():
    from neilpy_b200 import las
    rec = L.synth_records(0, 4)
    img = bytearray(L.write_las(rec, 0))
    img[104] = 130
    with pytest.raises(ValueError, match='LAZ'):
        las.parse_header(bytes(img))
    img[104] = 42
    with pytest.raises(ValueError):
        las.parse_header(bytes(img))
    img[104] = 0
    h = las.parse_header(bytes(img))
    with pytest.raises(ValueError):
        las.point_block(h, len(img) + 3)


def test_ground_classification_oracle():
    for fmt in (1, 7):
        rec = L.synth_records(fmt, 50, seed=5)
        obj = np.random.default_rng(1).random(50) < 0.4
        out = np.frombuffer(L.ground_classification(rec.tobytes(), fmt, obj), L.record_dtype(fmt))
        code = np.where(obj, 0, 2)
        if fmt < 6:
            assert np.array_equal(out['class'] & 0x1F, code)
            assert np.array_equal(out['class'] & 0xE0, rec['class'] & 0xE0)
        else:
            assert np.array_equal(out['class'], code)
        for nme in rec.dtype.names:
            if nme != 'class':
                assert np.array_equal(out[nme], rec[nme])
