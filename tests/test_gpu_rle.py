"""GPU parity: RLE codec kernels (vp_rle.cu) against the oracle through the C ABI."""
import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen

pytestmark = pytest.mark.gpu


def encode_world(w):
    """Oracle-encoded streams of every chunk: (words, offsets)."""
    streams = [helpers.rle_encode(w.dense[i]) for i in range(w.n_chunks)]
    offs = np.zeros(w.n_chunks + 1, np.uint64)
    offs[1:] = np.cumsum([s.size for s in streams])
    return np.concatenate(streams), offs


@pytest.mark.parametrize("rb,bits,kind", [(4, (1, 1, 1), "random"), (5, (1, 1, 1), "terrain"), (6, (1, 0, 1), "terrain"),
                                           (6, (1, 0, 0), "random"), (7, (0, 0, 1), "terrain")])
def test_upload_rle_decode_and_encode_roundtrip(rb, bits, kind):
    w = worldgen.World(31 + rb, rb, bits) if kind == "terrain" else helpers.random_world(31 + rb, rb, bits, density=0.4, maxv=255)
    words, offs = encode_world(w)
    ctx = vpb.Context(rb, bits, rle_arena_bytes=max(64 << 20, int(words.size) * 4 + (1 << 20)))
    try:
        ids = np.arange(w.n_chunks, dtype=np.uint32)
        ctx.upload_chunks_rle(ids, words, offs)                      # rle_decompress on the device
        assert np.array_equal(ctx.download_chunks_dense(ids), w.dense)
        got_words, got_offs = ctx.encode_chunks_rle(ids)              # rle_compress on the device
        assert np.array_equal(got_offs, offs)
        assert np.array_equal(got_words, words)
    finally:
        ctx.close()


def test_rle_upload_feeds_the_rebuild():
    w = worldgen.World(8, 5, (2, 1, 1))
    o = helpers.OracleWorld(w)
    words, offs = encode_world(w)
    ctx = vpb.Context(w.root_bitw, w.max_bitw)
    try:
        ids = np.arange(w.n_chunks, dtype=np.uint32)
        ctx.upload_chunks_rle(ids, words, offs)
        ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
        res, splat, mesh = ctx.rebuild_batch(ids, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
        for k, cid in enumerate(ids):
            geom, items = o.splat(int(cid))
            off = int(res["svl_offset"][k])
            assert np.array_equal(splat[off:off + geom.size * 2].view(np.int16), geom)
            vbo, ibo = o.mesh(int(cid))
            vo, io = int(res["vbo_offset"][k]), int(res["ibo_offset"][k])
            assert np.array_equal(mesh[vo:vo + vbo.size * 2].view(np.int16), vbo)
            assert np.array_equal(mesh[io:io + ibo.size * 4].view(np.uint32), ibo)
    finally:
        ctx.close()


def test_flat_codec_edge_cases():
    ctx = vpb.Context(5, (0, 0, 0))
    try:
        rng = np.random.default_rng(3)
        cases = [np.zeros(4096, np.uint8),                                        # one run: {4096, 0}
                 np.full(32768, 7, np.uint8),
                 np.arange(4096, dtype=np.uint32).astype(np.uint8),               # every byte starts a run
                 (np.arange(65536) // 3 % 256).astype(np.uint8),
                 rng.integers(0, 2, 262144).astype(np.uint8),                     # adversarial: ~N/2 runs (reference scratch would overflow, u4)
                 np.repeat(rng.integers(0, 256, 2048).astype(np.uint8), 1024)]   # 2 M bytes, long runs
        for d in cases:
            enc = ctx.rle_compress(d)
            want = helpers.rle_encode(d)
            assert np.array_equal(enc, want)
            assert np.array_equal(ctx.rle_decompress(want, d.size), d)
        # KAT from SURVEY 8(c)
        d = np.array([0, 0, 0, 5, 5, 7] + [0] * 9 + [9], np.uint8)
        assert list(ctx.rle_compress(d)) == [0x00000003, 0x05000002, 0x07000001, 0x00000009, 0x09000001, 0]
        # malformed stream (wrong length) is reported
        with pytest.raises(vpb.VoxplatError):
            ctx.upload_chunks_rle(np.array([0], np.uint32), np.array([5 | (3 << 24), 0], np.uint32), np.array([0, 2], np.uint64))
    finally:
        ctx.close()
