"""GPU parity against the committed golden fixtures (outputs of the compiled reference)."""
import numpy as np
import pytest

import voxplat_b200 as vpb
from test_golden import FIXTURES, load_fixture, expected_chunks
from test_gpu_splat import upload_world

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", FIXTURES)
def test_cuda_matches_golden(name):
    w, z = load_fixture(name)
    ctx = vpb.Context(w.root_bitw, w.max_bitw, mesh_arena_bytes=256 << 20)
    try:
        upload_world(ctx, w)
        ids = np.arange(w.n_chunks, dtype=np.uint32)
        res, splat, mesh = ctx.rebuild_batch(ids, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
        for c, want_splat, items, vbo, ibo in expected_chunks(w, z):
            assert np.array_equal(res["svl_items"][c], items), c
            off = int(res["svl_offset"][c])
            assert np.array_equal(splat[off:off + want_splat.size * 2].view(np.int16), want_splat), c
            vo, io = int(res["vbo_offset"][c]), int(res["ibo_offset"][c])
            assert np.array_equal(mesh[vo:vo + vbo.size * 2].view(np.int16), vbo), c
            assert np.array_equal(mesh[io:io + ibo.size * 4].view(np.uint32), ibo), c
        if z["rle"].size:
            words, offs = ctx.encode_chunks_rle(ids)
            assert np.array_equal(offs, z["rle_offsets"]) and np.array_equal(words, z["rle"])
    finally:
        ctx.close()
