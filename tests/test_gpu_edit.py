"""Device-side brush edits (vp_edit.cu) against the COMPILED reference's chunkset_edit_sphere + shadow_place_update:
voxels of every chunk, the whole height map and the dirty set must match after a burst of place / remove edits,
and the rebuilt buffers of the dirty chunks must match the oracle on the edited world (config C5's loop)."""
import ctypes as C

import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
from test_gpu_splat import upload_world

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not helpers.ref_available(), reason="oracle/_ref/libvoxref.so not built")
@pytest.mark.parametrize("rb,bits", [(5, (2, 1, 2)), (4, (2, 2, 2))])
def test_edit_sphere_matches_compiled_reference(rb, bits):
    w = worldgen.World(606, rb, bits)
    r = helpers.RefWorld(w)
    lib = r.lib
    lib.vr_chunk_voxels.restype = C.c_void_p
    lib.vr_shadow_ptr.restype = C.c_void_p
    ctx = vpb.Context(rb, bits, mesh_arena_bytes=256 << 20)
    try:
        upload_world(ctx, w)
        for i in range(w.n_chunks):
            lib.vr_chunk_dirty(r.set, C.c_uint32(i), 1)
        rng = np.random.default_rng(9)
        X, Y, Z = w.dims
        edits = [(0, 3, 0, 3, 9), (X - 1, Y - 2, Z - 1, 4, 7), (w.R, 1, w.R, 5, 63), (w.R - 1, 10, 2 * w.R, 4, 0)]      # corners, borders, y < 2
        edits += [(int(rng.integers(0, X)), int(rng.integers(0, min(Y, 60))), int(rng.integers(0, Z)), int(rng.integers(1, 7)),
                   int(rng.choice([0, 63, 17]))) for _ in range(30)]
        for (x, y, z, rad, v) in edits:
            lib.vr_edit_sphere(r.set, x, y, z, rad, v)
            dirty = ctx.edit_sphere(x, y, z, rad, v)
            want_dirty = [i for i in range(w.n_chunks) if lib.vr_chunk_dirty(r.set, C.c_uint32(i), 1)]
            assert sorted(dirty.tolist()) == want_dirty, (x, y, z, rad, v)
        ids = np.arange(w.n_chunks, dtype=np.uint32)
        got = ctx.download_chunks_dense(ids)
        for i in range(w.n_chunks):
            p = lib.vr_chunk_voxels(r.set, C.c_uint32(i))
            want = np.frombuffer(C.string_at(p, w.N), np.uint8)
            assert np.array_equal(got[i], want), i
        n = w.shw * Z
        want_sh = np.frombuffer(C.string_at(lib.vr_shadow_ptr(r.set), n * 2), np.uint16)
        assert np.array_equal(ctx.download_shadow_rows(0, Z), want_sh)
        # and the rebuild of the edited world is the reference's
        res, splat, mesh = ctx.rebuild_batch(ids, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
        for k in range(w.n_chunks):
            g, it = r.splat(k)
            off = int(res["svl_offset"][k])
            assert np.array_equal(res["svl_items"][k], it), k
            assert np.array_equal(splat[off:off + g.size * 2].view(np.int16), g), k
            v, x = r.mesh(k)
            vo, io = int(res["vbo_offset"][k]), int(res["ibo_offset"][k])
            assert np.array_equal(mesh[vo:vo + v.size * 2].view(np.int16), v), k
            assert np.array_equal(mesh[io:io + x.size * 4].view(np.uint32), x), k
    finally:
        ctx.close()
