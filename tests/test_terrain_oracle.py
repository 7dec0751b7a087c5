"""The terrain oracle (oracle/terrain_oracle.py) against what the reference's own slope /
aspect / hillshade / pssm returned (tests/golden/terrain_golden.npz), and the colour table
against the colours of the reference's shipped bonemap image."""
import json
import os

import numpy as np
import pytest

from oracle import terrain_oracle as T

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'terrain_golden.npz'))
INDEX = json.loads(bytes(GOLD['index']).decode())


def same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=a.dtype.kind == 'f')


@pytest.mark.parametrize('item', INDEX, ids=[i['key'] for i in INDEX])
def test_oracle_matches_reference_functions(item):
    Z = GOLD['dem_' + item['dem']]
    got = getattr(T, item['fn'])(Z.copy(), **item['kwargs'])
    assert same(got, GOLD[item['key']])


def test_bone_table_is_the_one_in_the_reference_png():
    """Every distinct colour of examples/dk22_smrfed_bonemap.png (pssm + plt.imsave) is an entry of
    floor(255 * bone_r), and all 256 entries occur."""
    png = GOLD['bonemap_png_colours']
    lut = (T.bone_lut(reverse=True) * 255).astype(np.uint8)
    assert len(png) == 256
    assert {tuple(c) for c in png} == {tuple(c) for c in lut}
    # bone and bone_r are each other's mirror image up to rounding of the table construction
    assert np.abs(T.bone_lut(False)[::-1] - T.bone_lut(True)).max() < 1e-12


def test_product_table_equals_oracle_table():
    from neilpy_b200 import terrain
    for rev in (False, True):
        assert np.array_equal(terrain.bone_table(reverse=rev), T.bone_lut(reverse=rev))
