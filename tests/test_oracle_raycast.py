"""Pins the restated pick ray (oracle vo_raycast -- the arithmetic of the device kernel k_raycast, operation for operation)
to the COMPILED reference's chunkset_edit_raycast_until_solid (chunkset/edit.c:248-314) on thousands of seeded rays:
random, camera-like, axis-aligned / planar (zero components: NaN and infinite distances), starting outside the world on
either side, and leaving through a 0-face (unsigned coordinates stick at 0xFFFFFFFF)."""
import numpy as np
import pytest

import helpers
from voxplat_b200 import worldgen

pytestmark = pytest.mark.skipif(not helpers.ref_available(), reason="oracle/_ref/libvoxref.so not built")


def rays(w, n, seed):
    rng = np.random.default_rng(seed)
    X, Y, Z = w.dims
    o = np.stack([rng.uniform(0, X, n), rng.uniform(0, Y, n), rng.uniform(0, Z, n)], axis=1).astype(np.float32)
    v = rng.normal(size=(n, 3)).astype(np.float32)
    k = n // 10
    o[:k, 1] = Y - 1.5                                          # camera-like: from above, looking down
    v[:k, 1] = -np.abs(v[:k, 1]) - 0.2
    v[k:k + 20] = [0, -1, 0]                                    # axis-aligned and planar rays
    v[k + 20:k + 40] = [1, 0, 0]
    v[k + 40:k + 60] = [0, 0, -1]
    v[k + 60:2 * k, 2] = 0
    o[2 * k:3 * k, 0] = X + rng.uniform(1, 20, k).astype(np.float32)          # start outside, both sides
    v[2 * k:3 * k, 0] = -np.abs(v[2 * k:3 * k, 0]) - 0.1
    o[3 * k:4 * k, 2] = -rng.uniform(1, 20, k).astype(np.float32)
    v[3 * k:4 * k, 2] = np.abs(v[3 * k:4 * k, 2]) + 0.1
    o[4 * k:5 * k, 0] = rng.uniform(0, 3, k).astype(np.float32)               # leave through x = 0
    v[4 * k:5 * k] = [-1, 0.01, 0.02]
    o[5 * k:6 * k] = np.floor(o[5 * k:6 * k])                                 # origins on cell corners (ties)
    v[6 * k:7 * k] = np.sign(v[6 * k:7 * k])                                  # exact diagonals (ties at every step)
    return o, v


@pytest.mark.parametrize("rb,bits,kind", [(5, (2, 1, 2), "terrain"), (4, (2, 1, 2), "random"), (6, (1, 0, 1), "terrain")])
def test_restated_raycast_equals_reference(rb, bits, kind):
    w = worldgen.World(31, rb, bits) if kind == "terrain" else helpers.random_world(31, rb, bits, density=0.02, null_frac=0.3)
    r, o = helpers.RefWorld(w), helpers.OracleWorld(w)
    org, vec = rays(w, 3000, 12 + rb)
    hits = 0
    for i in range(len(org)):
        want = r.raycast(org[i], vec[i])
        got = o.raycast(org[i], vec[i])
        assert got == want, (i, org[i], vec[i], got, want)
        hits += want[0] > 0
    assert 0 < hits < len(org)
