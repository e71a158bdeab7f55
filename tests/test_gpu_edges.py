"""Edge cases through the C ABI: empty batches, an all-air world, a single chunk world, chunks on every world face,
ids outside the slab, repeated ids."""
import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
from test_gpu_splat import upload_world

pytestmark = pytest.mark.gpu


def test_empty_batch_and_all_air_world():
    ctx = vpb.Context(5, (1, 1, 1))
    try:
        res, splat, mesh = ctx.rebuild_batch(np.zeros(0, np.uint32), vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
        assert len(res) == 0 and splat.size == 0 and mesh.size == 0
        ids = np.arange(8, dtype=np.uint32)
        res, splat, mesh = ctx.rebuild_batch(ids, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)       # every chunk is the null chunk
        assert not res["svl_items_total"].any() and not res["vbo_items"].any() and not res["ibo_items"].any()
        words, offs = ctx.encode_chunks_rle(ids)
        assert words.tolist() == [32 ** 3, 0] * 8                                                     # the shared null stream
    finally:
        ctx.close()


def test_single_chunk_world_and_full_solid():
    for density in (1.0, 0.5):
        w = helpers.random_world(21, 4, (0, 0, 0), density=density, null_frac=0.0)
        o = helpers.OracleWorld(w)
        ctx = vpb.Context(4, (0, 0, 0))
        try:
            upload_world(ctx, w)
            res, splat, mesh = ctx.rebuild_batch(np.array([0], np.uint32), vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
            g, it = o.splat(0)
            v, x = o.mesh(0)
            assert np.array_equal(res["svl_items"][0], it)
            assert np.array_equal(splat[:g.size * 2].view(np.int16), g)
            vo, io = int(res["vbo_offset"][0]), int(res["ibo_offset"][0])
            assert np.array_equal(mesh[vo:vo + v.size * 2].view(np.int16), v) and np.array_equal(mesh[io:io + x.size * 4].view(np.uint32), x)
        finally:
            ctx.close()


def test_bad_ids_are_rejected_and_duplicates_work():
    w = worldgen.World(4, 4, (1, 1, 1))
    o = helpers.OracleWorld(w)
    ctx = vpb.Context(4, (1, 1, 1))
    try:
        upload_world(ctx, w)
        with pytest.raises(vpb.VoxplatError) as e:
            ctx.rebuild_batch(np.array([8], np.uint32))                   # only ids 0..7 exist
        assert e.value.code == -6
        res, splat, _ = ctx.rebuild_batch(np.array([3, 3, 1], np.uint32))   # a chunk may be listed twice
        for k, cid in enumerate([3, 3, 1]):
            g, it = o.splat(cid)
            off = int(res["svl_offset"][k])
            assert np.array_equal(splat[off:off + g.size * 2].view(np.int16), g)
    finally:
        ctx.close()


def test_rebuild_in_two_parts_equals_one_call():
    """vp_rebuild_device_part(0) + (1) (the slab step that hides the border exchange) on a single-slab context and on a
    middle slab: same buffers per chunk as vp_rebuild_device."""
    w = worldgen.World(11, 5, (1, 1, 3))
    per_row = 4
    for slab_rows in [None, (2, 5)]:
        ctx = vpb.Context(5, (1, 1, 3), slab=slab_rows) if slab_rows else vpb.Context(5, (1, 1, 3))
        try:
            z0, z1 = slab_rows if slab_rows else (0, 8)
            own = np.arange(z0 * per_row, z1 * per_row, dtype=np.uint32)
            nn = own[w.solid[own] > 0]
            ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
            ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
            ctx.batch_prepare(own, flags=vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
            ctx.rebuild_device()
            r1, sb1, mb1 = ctx.rebuild_device_results()
            s1, m1 = ctx.arena_download(0, sb1), ctx.arena_download(1, mb1)
            ctx.rebuild_device_part(0)
            ctx.rebuild_device_part(1)
            r2, sb2, mb2 = ctx.rebuild_device_results()
            s2, m2 = ctx.arena_download(0, sb2), ctx.arena_download(1, mb2)
            assert sb1 == sb2 and mb1 == mb2
            for k in range(len(own)):
                assert np.array_equal(r1["svl_items"][k], r2["svl_items"][k]) and r1["vbo_items"][k] == r2["vbo_items"][k]
                a, b, nb = int(r1["svl_offset"][k]), int(r2["svl_offset"][k]), int(r1["svl_items_total"][k]) * 2
                assert np.array_equal(s1[a:a + nb], s2[b:b + nb])
                a, b, nb = int(r1["vbo_offset"][k]), int(r2["vbo_offset"][k]), int(r1["vbo_items"][k]) * 2
                assert np.array_equal(m1[a:a + nb], m2[b:b + nb])
                a, b, nb = int(r1["ibo_offset"][k]), int(r2["ibo_offset"][k]), int(r1["ibo_items"][k]) * 4
                assert np.array_equal(m1[a:a + nb], m2[b:b + nb])
        finally:
            ctx.close()
