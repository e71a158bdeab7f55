"""GPU parity: cull + LOD + splat emission (vp_splat.cu) against the oracle, byte for byte, through the C ABI."""
import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen

pytestmark = pytest.mark.gpu


def upload_world(ctx, w):
    ids = w.nonnull_ids()
    if len(ids):
        ctx.upload_chunks_dense(ids, np.ascontiguousarray(w.dense[ids]))
    ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])


def check_splat(w, ids=None):
    o = helpers.OracleWorld(w)
    ctx = vpb.Context(w.root_bitw, w.max_bitw, splat_arena_bytes=max(64 << 20, w.n_chunks * (w.R + 1) ** 3 * 10))
    try:
        upload_world(ctx, w)
        ids = np.arange(w.n_chunks, dtype=np.uint32) if ids is None else np.asarray(ids, np.uint32)
        res, splat, _ = ctx.rebuild_batch(ids, vpb.VP_REBUILD_SPLAT)
        total = 0
        for k, cid in enumerate(ids):
            geom, items = o.splat(int(cid))
            assert np.array_equal(res["svl_items"][k], items), (cid, res["svl_items"][k], items)
            assert res["svl_items_total"][k] == geom.size
            off = int(res["svl_offset"][k])
            got = splat[off:off + geom.size * 2].view(np.int16)
            if not np.array_equal(got, geom):
                bad = np.nonzero(got != geom)[0]
                raise AssertionError("chunk %d: %d of %d int16 differ, first at %d: got %s want %s" % (
                    cid, bad.size, geom.size, bad[0], got[bad[0] // 4 * 4:bad[0] // 4 * 4 + 4], geom[bad[0] // 4 * 4:bad[0] // 4 * 4 + 4]))
            total += geom.size
        return total
    finally:
        ctx.close()


@pytest.mark.parametrize("rb,bits", [(4, (1, 1, 1)), (5, (1, 1, 1)), (6, (1, 0, 1)), (7, (1, 0, 0))])
@pytest.mark.parametrize("density", [0.02, 0.5, 0.97])
def test_random_worlds(rb, bits, density):
    w = helpers.random_world(7 * rb + int(density * 100), rb, bits, density=density, null_frac=0.25)
    assert check_splat(w) > 0


def test_mixed_density_and_full_chunks():
    # full-solid and empty chunks next to each other: only halo cells / border faces are visible
    w = helpers.random_world(3, 5, (2, 1, 1), density=[1.0, 0.0, 1.0, 0.3], null_frac=0.0)
    check_splat(w)


@pytest.mark.parametrize("rb,bits", [(4, (2, 1, 2)), (5, (2, 1, 2)), (6, (1, 1, 1))])
def test_terrain_worlds(rb, bits):
    w = worldgen.World(1234, rb, bits)
    assert check_splat(w) > 0


def test_single_chunk_wrapper_and_repeatability():
    w = worldgen.World(99, 5, (1, 1, 1))
    o = helpers.OracleWorld(w)
    ctx = vpb.Context(w.root_bitw, w.max_bitw)
    try:
        upload_world(ctx, w)
        for cid in range(w.n_chunks):
            geom, items = ctx.chunk_make_splatlists(cid)
            want, witems = o.splat(cid)
            assert np.array_equal(items, witems) and np.array_equal(geom, want)
        # two full rebuilds give the same bytes per chunk (arena placement may differ)
        ids = np.arange(w.n_chunks, dtype=np.uint32)
        r1, s1, _ = ctx.rebuild_batch(ids)
        a = [bytes(s1[int(r["svl_offset"]):int(r["svl_offset"]) + int(r["svl_items_total"]) * 2]) for r in r1]
        r2, s2, _ = ctx.rebuild_batch(ids)
        b = [bytes(s2[int(r["svl_offset"]):int(r["svl_offset"]) + int(r["svl_items_total"]) * 2]) for r in r2]
        assert a == b
    finally:
        ctx.close()


def test_arena_overflow_is_reported():
    w = helpers.random_world(11, 5, (1, 1, 1), density=0.5, null_frac=0.0)
    ctx = vpb.Context(w.root_bitw, w.max_bitw, splat_arena_bytes=1 << 16)
    try:
        upload_world(ctx, w)
        with pytest.raises(vpb.VoxplatError) as e:
            ctx.rebuild_batch(np.arange(w.n_chunks, dtype=np.uint32))
        assert e.value.code == -4
    finally:
        ctx.close()


def test_dense_roundtrip_and_null_detection():
    w = helpers.random_world(12, 4, (1, 1, 1), density=0.3, null_frac=0.5)
    ctx = vpb.Context(w.root_bitw, w.max_bitw)
    try:
        ids = np.arange(w.n_chunks, dtype=np.uint32)
        ctx.upload_chunks_dense(ids, w.dense)           # all-zero chunks become null chunks
        back = ctx.download_chunks_dense(ids)
        assert np.array_equal(back, w.dense)
    finally:
        ctx.close()
