"""slope / aspect / hillshade / pssm on the device (neilpy_b200.terrain -> smrf_terrain) against
what the reference's own functions returned (tests/golden/terrain_golden.npz).

Tolerances: sqrt / products / sums are correctly rounded on both sides; atan, atan2, sin, cos
are CUDA's (<= 2 ulp) versus the host libm's, so float outputs are held to 1e-13 relative
(+1e-13 absolute) and the uint8 outputs to at most one cell per thousand off by one level
(a product landing within an ulp of a rounding tie)."""
import json
import os

import numpy as np
import pytest

from oracle import terrain_oracle as T

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'terrain_golden.npz'))
INDEX = json.loads(bytes(GOLD['index']).decode())
RTOL = ATOL = 1e-13


def check(got, want, lut=None):
    assert got.shape == want.shape and got.dtype == want.dtype
    if want.dtype == np.uint8:
        d = np.abs(got.astype(int) - want.astype(int))
        assert d.max() <= 1 and (d > 0).sum() <= max(1, d.size // 1000), (int(d.max()), int((d > 0).sum()))
    elif want.ndim == 3:                              # rgba: exact table entries, same cells as the index
        bad = (got != want).any(axis=2)
        assert bad.sum() <= max(1, bad.size // 1000), int(bad.sum())
        table = {tuple(c) for c in lut}
        assert all(tuple(c) in table for c in np.unique(got.reshape(-1, 4), axis=0))
    else:
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.allclose(got, want, rtol=RTOL, atol=ATOL, equal_nan=True), float(np.nanmax(np.abs(got - want)))


@pytest.mark.parametrize('item', INDEX, ids=[i['key'] for i in INDEX])
def test_matches_reference_functions(item):
    from neilpy_b200 import terrain
    Z = GOLD['dem_' + item['dem']]
    got = getattr(terrain, item['fn'])(Z.copy(), **item['kwargs'])
    lut = None
    if item['fn'] == 'pssm' and item['kwargs'].get('apply_colormap', True):
        lut = np.ascontiguousarray(terrain.bone_table(reverse=not item['kwargs'].get('reverse', False)))
    check(got, GOLD[item['key']], lut)


def test_large_grid_device_in_device_out_and_float32():
    import torch
    from neilpy_b200 import terrain
    Z = T.synth_dem(1031, 2053, seed=9)
    Zd = torch.as_tensor(Z).cuda()
    P = terrain.pssm(Zd, cellsize=2, apply_colormap=False)
    assert P.is_cuda and P.dtype == torch.uint8
    check(P.cpu().numpy(), T.pssm(Z, cellsize=2, apply_colormap=False))
    check(terrain.hillshade(Zd, cellsize=2).cpu().numpy(), T.hillshade(Z, cellsize=2))
    check(terrain.slope(Zd, cellsize=2, return_as='percent').cpu().numpy(), T.slope(Z, cellsize=2, return_as='percent'))
    check(terrain.aspect(Zd).cpu().numpy(), T.aspect(Z))
    rgba = terrain.pssm(Zd, cellsize=2)
    assert rgba.shape == (1031, 2053, 4) and float(rgba[..., 3].min()) == 1.0
    # float32 grid: widened per cell, i.e. the float64 result of the float32 values
    Z32 = Z.astype(np.float32)
    check(terrain.hillshade(Z32, cellsize=2), T.hillshade(Z32.astype(np.float64), cellsize=2))


def test_flat_grid_nan_cells_and_errors():
    from neilpy_b200 import terrain
    Z = np.full((9, 11), 5.0)
    assert np.isnan(terrain.aspect(Z)).all()
    assert (terrain.aspect(Z, flat_as=-1) == -1).all()
    assert (terrain.slope(Z) == 0).all()
    assert np.array_equal(terrain.hillshade(Z), T.hillshade(Z))
    assert (terrain.pssm(Z, apply_colormap=False) == 0).all()
    Z[4, 5] = np.nan
    H = terrain.hillshade(Z)
    assert H[4, 5] == np.uint8(180) and H[4, 4] == 0 and H[0, 0] == np.uint8(180)   # neighbours of the NaN see a NaN gradient
    with pytest.raises(ValueError):
        terrain.slope(Z, return_as='grads')
    with pytest.raises(ValueError):
        terrain.aspect(Z, return_as='percent')
    with pytest.raises(ValueError):
        terrain.pssm(np.zeros(5))
