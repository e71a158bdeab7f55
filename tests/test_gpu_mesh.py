"""GPU parity: near-field quad mesh (vp_mesh.cu) against the oracle, byte for byte, through the C ABI."""
import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
from test_gpu_splat import upload_world

pytestmark = pytest.mark.gpu


def check_mesh(w, ids=None, also_splat=False):
    o = helpers.OracleWorld(w)
    ctx = vpb.Context(w.root_bitw, w.max_bitw, mesh_arena_bytes=max(64 << 20, w.n_chunks * w.N * 90),
                      splat_arena_bytes=max(64 << 20, w.n_chunks * (w.R + 1) ** 3 * 10))
    try:
        upload_world(ctx, w)
        ids = np.arange(w.n_chunks, dtype=np.uint32) if ids is None else np.asarray(ids, np.uint32)
        flags = vpb.VP_REBUILD_MESH | (vpb.VP_REBUILD_SPLAT if also_splat else 0)
        res, splat, mesh = ctx.rebuild_batch(ids, flags)
        faces = 0
        for k, cid in enumerate(ids):
            vbo, ibo = o.mesh(int(cid))
            assert res["vbo_items"][k] == vbo.size and res["ibo_items"][k] == ibo.size, (cid, res["vbo_items"][k], vbo.size)
            vo, io = int(res["vbo_offset"][k]), int(res["ibo_offset"][k])
            gv = mesh[vo:vo + vbo.size * 2].view(np.int16)
            gi = mesh[io:io + ibo.size * 4].view(np.uint32)
            if not np.array_equal(gv, vbo):
                bad = np.nonzero(gv != vbo)[0]
                f = bad[0] // 16
                raise AssertionError("chunk %d: %d vbo int16 differ, first in face %d:\n got  %s\n want %s" % (
                    cid, bad.size, f, gv[f * 16:f * 16 + 16], vbo[f * 16:f * 16 + 16]))
            assert np.array_equal(gi, ibo), cid
            faces += ibo.size // 6
            if also_splat:
                geom, items = o.splat(int(cid))
                off = int(res["svl_offset"][k])
                assert np.array_equal(res["svl_items"][k], items)
                assert np.array_equal(splat[off:off + geom.size * 2].view(np.int16), geom)
        return faces
    finally:
        ctx.close()


@pytest.mark.parametrize("rb,bits", [(4, (1, 1, 1)), (5, (1, 1, 1)), (6, (1, 0, 1)), (7, (1, 0, 0))])
@pytest.mark.parametrize("density", [0.03, 0.5, 0.95])
def test_random_worlds(rb, bits, density):
    w = helpers.random_world(900 + 7 * rb + int(density * 100), rb, bits, density=density, null_frac=0.25)
    assert check_mesh(w) > 0


def test_neighbourhood_corners():
    # 3x3x3 chunks: the centre chunk sees all 26 neighbours (AO across edges and corners)
    w = helpers.random_world(77, 4, (2, 2, 2), density=0.45, null_frac=0.1)
    check_mesh(w)


@pytest.mark.parametrize("rb,bits", [(4, (2, 1, 2)), (5, (2, 1, 2)), (6, (1, 1, 1))])
def test_terrain_worlds(rb, bits):
    w = worldgen.World(4321, rb, bits)
    assert check_mesh(w, also_splat=True) > 0


def test_single_chunk_wrapper():
    w = worldgen.World(5, 5, (1, 1, 1))
    o = helpers.OracleWorld(w)
    ctx = vpb.Context(w.root_bitw, w.max_bitw)
    try:
        upload_world(ctx, w)
        for cid in range(w.n_chunks):
            vbo, ibo = ctx.chunk_make_mesh(cid)
            wv, wi = o.mesh(cid)
            assert np.array_equal(vbo, wv) and np.array_equal(ibo, wi)
    finally:
        ctx.close()
