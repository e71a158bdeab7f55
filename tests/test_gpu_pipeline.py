"""The pipelined one-call e2e path (vp_rebuild_from_rle) gives the same per-chunk bytes as the oracle for any
block count, including mesh chunks whose neighbours live in other blocks."""
import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
from test_gpu_rle import encode_world

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_blocks", [1, 3, 8, 64])
def test_rebuild_from_rle_matches_oracle(n_blocks):
    w = worldgen.World(314, 5, (2, 1, 3))                 # 4 x 2 x 8 chunks of 32^3
    o = helpers.OracleWorld(w)
    words, offs = encode_world(w)
    ids = np.arange(w.n_chunks, dtype=np.uint32)
    flags = np.array([3 if (i % 5) else 1 for i in ids], np.uint8)        # most chunks splat + mesh
    ctx = vpb.Context(w.root_bitw, w.max_bitw, mesh_arena_bytes=256 << 20)
    try:
        ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
        for rep in range(2):                                  # second call takes the staged (pipelined download) path
            res, splat, mesh = ctx.rebuild_from_rle(ids, words, offs, per_chunk_flags=flags, n_blocks=n_blocks)
            for k, cid in enumerate(ids):
                g, it = o.splat(int(cid))
                off = int(res["svl_offset"][k])
                assert np.array_equal(res["svl_items"][k], it), (rep, cid)
                assert np.array_equal(splat[off:off + g.size * 2].view(np.int16), g), (rep, cid)
                if flags[k] & 2:
                    v, x = o.mesh(int(cid))
                    vo, io = int(res["vbo_offset"][k]), int(res["ibo_offset"][k])
                    assert res["vbo_items"][k] == v.size
                    assert np.array_equal(mesh[vo:vo + v.size * 2].view(np.int16), v), (rep, cid)
                    assert np.array_equal(mesh[io:io + x.size * 4].view(np.uint32), x), (rep, cid)
                else:
                    assert res["vbo_items"][k] == 0
    finally:
        ctx.close()


def test_rebuild_from_rle_subset_and_unsorted_ids():
    w = worldgen.World(99, 5, (1, 1, 2))
    o = helpers.OracleWorld(w)
    words, offs = encode_world(w)
    all_ids = np.arange(w.n_chunks, dtype=np.uint32)
    ctx = vpb.Context(w.root_bitw, w.max_bitw)
    try:
        ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
        ctx.upload_chunks_rle(all_ids, words, offs)             # everything resident first
        pick = np.array([5, 2, 7], np.uint32)                    # unsorted subset: falls back to one block
        sub_words = np.concatenate([words[int(offs[i]):int(offs[i + 1])] for i in pick])
        sub_offs = np.concatenate([[0], np.cumsum([int(offs[i + 1] - offs[i]) for i in pick])]).astype(np.uint64)
        res, splat, _ = ctx.rebuild_from_rle(pick, sub_words, sub_offs, n_blocks=4)
        for k, cid in enumerate(pick):
            g, it = o.splat(int(cid))
            off = int(res["svl_offset"][k])
            assert np.array_equal(splat[off:off + g.size * 2].view(np.int16), g)
    finally:
        ctx.close()
