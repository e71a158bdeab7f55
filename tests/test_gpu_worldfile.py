"""World file: vp_world_save writes the reference exporter's layout (deadcode.c:320-350) with the
streams of the reference's rle_compress; vp_world_load restores a world whose rebuild is byte-identical."""
import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
from test_gpu_rle import encode_world

pytestmark = pytest.mark.gpu


def expected_file(w):
    """The file as the reference would write it: header, rle_compress stream of every chunk (oracle encoder, checked
    against the compiled reference in test_oracle_vs_reference), whole shadow map."""
    words, _ = encode_world(w)
    head = bytes([0x89]) + b"VOXPLAT" + bytes([w.root_bitw, *w.max_bitw])
    return head + words.astype("<u4").tobytes() + w.shadow[:w.shw * w.dims[2]].astype("<u2").tobytes()


@pytest.mark.parametrize("rb,bits,kind", [(5, (2, 1, 2), "terrain"), (4, (1, 1, 2), "random"), (6, (1, 0, 1), "terrain")])
def test_save_matches_reference_layout_and_load_restores(tmp_path, rb, bits, kind):
    w = worldgen.World(77 + rb, rb, bits) if kind == "terrain" else helpers.random_world(77 + rb, rb, bits, density=0.3, null_frac=0.3)
    path = str(tmp_path / "default.bin")
    ids = np.arange(w.n_chunks, dtype=np.uint32)
    ctx = vpb.Context(rb, bits)
    try:
        nn = w.nonnull_ids()
        ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
        ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
        size = ctx.save_world(path)
    finally:
        ctx.close()
    data = open(path, "rb").read()
    assert size == len(data)
    assert data == expected_file(w)
    assert vpb.world_file_info(path) == (rb, tuple(bits), len(data))

    o = helpers.OracleWorld(w)
    ctx = vpb.Context(rb, bits)
    try:
        ctx.load_world(path)
        assert np.array_equal(ctx.download_chunks_dense(ids), w.dense)
        assert np.array_equal(ctx.download_shadow_rows(0, w.dims[2]), w.shadow[:w.shw * w.dims[2]])
        res, splat, mesh = ctx.rebuild_batch(ids, flags=vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
        for k, cid in enumerate(ids):
            g, it = o.splat(int(cid))
            off = int(res["svl_offset"][k])
            assert np.array_equal(res["svl_items"][k], it)
            assert np.array_equal(splat[off:off + g.size * 2].view(np.int16), g)
            v, x = o.mesh(int(cid))
            vo, io = int(res["vbo_offset"][k]), int(res["ibo_offset"][k])
            assert np.array_equal(mesh[vo:vo + v.size * 2].view(np.int16), v)
            assert np.array_equal(mesh[io:io + x.size * 4].view(np.uint32), x)
    finally:
        ctx.close()


def test_load_rejects_bad_files(tmp_path):
    w = worldgen.World(5, 4, (1, 1, 1))
    good = expected_file(w)
    ctx = vpb.Context(4, (1, 1, 1))
    try:
        cases = {"magic": b"\x89VOXPLAX" + good[8:], "geometry": good[:8] + bytes([4, 1, 1, 2]) + good[12:],
                 "truncated": good[:-6], "trailing": good + b"\0\0\0\0", "cut stream": good[:40]}
        for name, blob in cases.items():
            p = tmp_path / (name.replace(" ", "_") + ".bin")
            p.write_bytes(blob)
            with pytest.raises(vpb.VoxplatError):
                ctx.load_world(str(p))
        with pytest.raises(vpb.VoxplatError):
            ctx.load_world(str(tmp_path / "missing.bin"))
        (tmp_path / "ok.bin").write_bytes(good)
        ctx.load_world(str(tmp_path / "ok.bin"))                 # the context still works after the failures
        assert np.array_equal(ctx.download_chunks_dense(np.arange(8, dtype=np.uint32)), w.dense)
    finally:
        ctx.close()
