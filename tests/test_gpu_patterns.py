"""GPU parity on structured worst cases for the emission path: checkerboards (every row full, every +x halo cell
visible), flat floors (rows of 64 set bits between empty rows), thin vertical walls (one bit per row), voxels on chunk
corners, values above 63 (colour bits overlap the shadow bit) -- splat and mesh against the oracle, byte for byte."""
import numpy as np
import pytest

import voxplat_b200 as vpb
from voxplat_b200 import worldgen
from test_gpu_mesh import check_mesh
from test_gpu_splat import check_splat

pytestmark = pytest.mark.gpu


def build(rb, bits, fn):
    """World whose voxel at world coordinates (x,y,z) is fn(x,y,z) (vectorised over index grids)."""
    R = 1 << rb
    nx, ny, nz = (1 << b for b in bits)
    z, y, x = np.meshgrid(np.arange(nz * R), np.arange(ny * R), np.arange(nx * R), indexing="ij")
    vol = fn(x, y, z).astype(np.uint8)
    dense = vol.reshape(nz, R, ny, R, nx, R).transpose(0, 2, 4, 1, 3, 5).reshape(nx * ny * nz, R ** 3)
    return worldgen.World(1, rb, bits, dense=dense)


PATTERNS = {
    "checkerboard": lambda x, y, z: ((x + y + z) & 1) * (1 + (x * 7 + y * 13 + z * 29) % 255),
    "checkerboard2": lambda x, y, z: (((x >> 1) + (y >> 1) + (z >> 1)) & 1) * 200,
    "floors": lambda x, y, z: (y % 9 == 4) * (65 + x % 190),
    "walls_x": lambda x, y, z: (x % 16 == 15) * 255,
    "walls_z": lambda x, y, z: (z % 16 == 0) * (1 + y % 63),
    "pillars": lambda x, y, z: ((x % 5 == 0) & (z % 7 == 0)) * 77,
    "corners": lambda x, y, z: (((x % 32 == 0) | (x % 32 == 31)) & ((y % 32 == 0) | (y % 32 == 31)) & ((z % 32 == 0) | (z % 32 == 31))) * 129,
    "solid_with_holes": lambda x, y, z: 255 - 255 * ((x % 11 == 3) & (y % 6 == 2) & (z % 4 == 1)),
}


@pytest.mark.parametrize("name", sorted(PATTERNS))
@pytest.mark.parametrize("rb,bits", [(5, (1, 1, 1)), (6, (1, 0, 1))])
def test_pattern_splat_and_mesh(name, rb, bits):
    w = build(rb, bits, PATTERNS[name])
    assert check_splat(w) > 0
    if rb == 5 or not name.startswith("checkerboard"):          # a 64^3 checkerboard has 3 faces per voxel: 44 MB per chunk
        check_mesh(w)


def test_pattern_128():
    w = build(7, (1, 0, 0), PATTERNS["checkerboard"])           # two 64-bit words per row, all of them full
    assert check_splat(w) > 0
