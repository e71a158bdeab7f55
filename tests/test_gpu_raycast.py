"""Device pick rays (vp_raycast, vp_edit.cu) against the COMPILED reference's chunkset_edit_raycast_until_solid
(chunkset/edit.c:248-314): the voxel hit, the cell where the walk ended and the normal written must be identical for
seeded random rays, axis-aligned rays (zero components: the reference's NaN / infinity distances), rays that start
outside the world and rays that leave through a 0-face (the reference's unsigned coordinates stick at 0xFFFFFFFF)."""
import ctypes as C

import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
from test_gpu_splat import upload_world

pytestmark = pytest.mark.gpu


def reference_rays(r, origins, vectors):
    n = len(origins)
    vox, coords, nrm = np.zeros(n, np.uint8), np.zeros((n, 3), np.uint32), np.zeros((n, 3), np.int8)
    for i in range(n):
        o = (C.c_float * 3)(*[float(x) for x in origins[i]])
        v = (C.c_float * 3)(*[float(x) for x in vectors[i]])
        c = (C.c_uint32 * 3)()
        m = (C.c_int8 * 3)()
        vox[i] = r.lib.vr_raycast(r.set, o, v, c, m)
        coords[i] = list(c)
        nrm[i] = list(m)
    return vox, coords, nrm


@pytest.mark.skipif(not helpers.ref_available(), reason="oracle/_ref/libvoxref.so not built")
@pytest.mark.parametrize("rb,bits,kind", [(5, (2, 1, 2), "terrain"), (4, (2, 1, 2), "random"), (6, (1, 0, 1), "terrain")])
def test_raycast_matches_compiled_reference(rb, bits, kind):
    w = worldgen.World(31, rb, bits) if kind == "terrain" else helpers.random_world(31, rb, bits, density=0.02, null_frac=0.3)
    r = helpers.RefWorld(w)
    rng = np.random.default_rng(12)
    X, Y, Z = w.dims
    n = 400
    o = np.stack([rng.uniform(0, X, n), rng.uniform(0, Y, n), rng.uniform(0, Z, n)], axis=1).astype(np.float32)
    v = rng.normal(size=(n, 3)).astype(np.float32)
    # camera-like rays: from above, looking down at an angle
    o[:100, 1] = Y - 1.5
    v[:100, 1] = -np.abs(v[:100, 1]) - 0.2
    # axis-aligned and planar rays (zero components)
    v[100:110] = [0, -1, 0]
    v[110:120] = [1, 0, 0]
    v[120:130] = [0, 0, -1]
    v[130:150, 2] = 0
    # rays that start outside the world, on both sides
    o[150:170, 0] = X + rng.uniform(1, 20, 20).astype(np.float32)
    v[150:170, 0] = -np.abs(v[150:170, 0]) - 0.1
    o[170:190, 2] = -rng.uniform(1, 20, 20).astype(np.float32)
    v[170:190, 2] = np.abs(v[170:190, 2]) + 0.1
    # rays that leave through the x = 0 face
    o[190:210, 0] = rng.uniform(0, 3, 20).astype(np.float32)
    v[190:210] = [-1, 0.01, 0.02]
    want = reference_rays(r, o, v)
    ctx = vpb.Context(rb, bits)
    try:
        upload_world(ctx, w)
        got = ctx.raycast(o, v)
    finally:
        ctx.close()
    assert want[0].any() and not want[0].all()                 # hits and misses both occur
    for k, name in enumerate(("voxel", "coord", "normal")):
        bad = np.nonzero((got[k] != want[k]).reshape(n, -1).any(axis=1))[0]
        assert len(bad) == 0, (name, bad[:5], got[k][bad[:5]], want[k][bad[:5]])
