"""Full-size parity (BASELINE config C2: 2048x256x2048, 4096 chunks of 64^3): every chunk's splat buffer and the
near-field chunks' mesh buffers are hashed (FNV-1a 64) and compared with the reference run on the host cores
(compiled reference when oracle/_ref travelled, else the oracle port), plus size-independent properties."""
import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import slab, worldgen

pytestmark = pytest.mark.gpu


def test_c2_world_byte_exact_hashes_and_properties():
    rb, bits = 6, (5, 2, 5)
    w = worldgen.World(1234, rb, bits)
    ids = np.arange(w.n_chunks, dtype=np.uint32)
    near = ids[slab.near_camera_flags(ids, rb, bits)]
    checker = helpers.RefWorld(w) if helpers.ref_available() else helpers.OracleWorld(w)
    _, want_splat, want_counts = checker.rebuild(ids, 0)
    _, want_mesh, want_mcounts = checker.rebuild(near, 1)

    ctx = vpb.Context(rb, bits, splat_arena_bytes=1 << 30, mesh_arena_bytes=1 << 30, rle_arena_bytes=1 << 30)
    try:
        nn = w.nonnull_ids()
        ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
        ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
        flags = np.where(slab.near_camera_flags(ids, rb, bits), 3, 1).astype(np.uint8)
        res, splat, mesh = ctx.rebuild_batch(ids, per_chunk_flags=flags)
        # counts
        assert np.array_equal(res["svl_items"], want_counts[:, :5])
        assert np.array_equal(res["vbo_items"][near], want_mcounts[:, 5]) and np.array_equal(res["ibo_items"][near], want_mcounts[:, 6])
        # bytes, chunk by chunk, through the hash the reference harness uses
        for k in range(w.n_chunks):
            n = int(res["svl_items_total"][k])
            if n or want_splat[k] != helpers.fnv1a(np.zeros(0, np.uint8)):
                off = int(res["svl_offset"][k])
                assert helpers.fnv1a(splat[off:off + n * 2]) == int(want_splat[k]), k
        lib = helpers.oracle_lib()
        import ctypes as C
        for j, k in enumerate(near):
            vo, io = int(res["vbo_offset"][k]), int(res["ibo_offset"][k])
            v = np.ascontiguousarray(mesh[vo:vo + int(res["vbo_items"][k]) * 2])
            x = np.ascontiguousarray(mesh[io:io + int(res["ibo_items"][k]) * 4])
            h = lib.vo_fnv1a(C.c_void_p(v.ctypes.data), C.c_uint64(v.nbytes), C.c_uint64(0))
            h = lib.vo_fnv1a(C.c_void_p(x.ctypes.data), C.c_uint64(x.nbytes), C.c_uint64(h))
            assert int(h) == int(want_mesh[j]), k
        # size-independent properties: the per-chunk buffers tile the arena exactly (no overlap, no gap) ...
        sizes = res["svl_items_total"].astype(np.int64) * 2
        order = np.argsort(res["svl_offset"][sizes > 0])
        offs = res["svl_offset"][sizes > 0][order].astype(np.int64)
        assert offs[0] == 0 and np.array_equal(offs[1:], (offs + sizes[sizes > 0][order])[:-1])
        # ... every LOD level is no larger than the one below, every index quad references its own 4 vertices ...
        it = res["svl_items"].astype(np.int64)
        assert (it[:, 1:] <= it[:, :-1]).all() and (res["ibo_items"] * 16 == res["vbo_items"] * 6).all()
        # ... and the device RLE codec round-trips the whole world
        words, woffs = ctx.encode_chunks_rle(nn)
        ctx.upload_chunks_rle(nn, words, woffs)
        sample = nn[:: max(1, len(nn) // 64)]
        assert np.array_equal(ctx.download_chunks_dense(sample), w.dense[sample])
        res2, splat2, _ = ctx.rebuild_batch(ids, vpb.VP_REBUILD_SPLAT)
        assert np.array_equal(res2["svl_items"], res["svl_items"])           # idempotent on identical input
    finally:
        ctx.close()
