"""vp_multi: several z-slabs behind one C handle, one host thread, border planes pushed peer to peer
(voxplat_b200/csrc/vp_multi.cu).  Every chunk of the world must come out byte-identical to the oracle evaluated on the
WHOLE world -- the only thing a slab cut can break is the cross-slab halo (mesher.c:391-432 for the cull, :119-171 for
mesh AO).  On a 1-GPU box the slabs share device 0 (the exchange is then plain device memory); with >= 2 GPUs the same
test also runs with one slab per device, i.e. over NVLink peer access."""
import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen

pytestmark = pytest.mark.gpu


def device_sets(nslab):
    import torch
    sets = [[0] * nslab]
    if torch.cuda.device_count() >= nslab:
        sets.append(list(range(nslab)))
    return sets


def check_world(w, devices, flags):
    o = helpers.OracleWorld(w)
    m = vpb.MultiContext(w.root_bitw, w.max_bitw, devices, splat_arena_bytes=64 << 20, mesh_arena_bytes=256 << 20)
    try:
        ids = np.arange(w.n_chunks, dtype=np.uint32)
        nn = ids[w.solid[ids] > 0]
        m.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
        m.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
        for _ in range(2):                       # twice: the second step reuses ghost slices, events and scratch
            res, owner, splat, mesh = m.rebuild_batch(ids, flags)
            assert set(owner.tolist()) == set(range(len(devices)))
            for k, cid in enumerate(ids):
                sp, me = splat[owner[k]], mesh[owner[k]]
                if flags & vpb.VP_REBUILD_SPLAT:
                    g, it = o.splat(int(cid))
                    off = int(res["svl_offset"][k])
                    assert np.array_equal(res["svl_items"][k], it), (cid, devices)
                    assert np.array_equal(sp[off:off + g.size * 2].view(np.int16), g), (cid, devices)
                if flags & vpb.VP_REBUILD_MESH:
                    v, x = o.mesh(int(cid))
                    vo, io = int(res["vbo_offset"][k]), int(res["ibo_offset"][k])
                    assert res["vbo_items"][k] == v.size and res["ibo_items"][k] == x.size, (cid, devices)
                    assert np.array_equal(me[vo:vo + v.size * 2].view(np.int16), v), (cid, devices)
                    assert np.array_equal(me[io:io + x.size * 4].view(np.uint32), x), (cid, devices)
    finally:
        m.close()


@pytest.mark.parametrize("nslab", [2, 4])
def test_multi_random_world_matches_oracle(nslab):
    w = helpers.random_world(91, 4, (1, 1, 3), density=0.4, null_frac=0.2)
    for devs in device_sets(nslab):
        check_world(w, devs, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)


def test_multi_terrain_world_matches_oracle():
    w = worldgen.World(77, 5, (2, 1, 2))
    for devs in device_sets(2):
        check_world(w, devs, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
    w = worldgen.World(78, 6, (1, 0, 2))
    for devs in device_sets(4):
        check_world(w, devs, vpb.VP_REBUILD_SPLAT)


def test_multi_subset_and_rle_upload():
    """A batch that names only some chunks (what the dispatcher sends after an edit), chunks uploaded as RLE streams."""
    w = worldgen.World(5, 5, (1, 1, 2))
    o = helpers.OracleWorld(w)
    m = vpb.MultiContext(w.root_bitw, w.max_bitw, [0, 0], splat_arena_bytes=64 << 20, mesh_arena_bytes=128 << 20)
    try:
        ids = np.arange(w.n_chunks, dtype=np.uint32)
        nn = ids[w.solid[ids] > 0]
        streams = [helpers.rle_encode(w.dense[i]) for i in nn]
        offs = np.zeros(len(nn) + 1, np.uint64)
        offs[1:] = np.cumsum([s.size for s in streams])
        m.upload_chunks_rle(nn, np.concatenate(streams), offs)
        m.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
        per_row = 1 << (w.max_bitw[0] + w.max_bitw[1])
        sub = np.array([per_row * 1 + 1, per_row * 2, 0, per_row * 3 + 2], np.uint32)      # both sides of the cut at row 2
        res, owner, splat, mesh = m.rebuild_batch(sub, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
        assert owner.tolist() == [0, 1, 0, 1]
        for k, cid in enumerate(sub):
            g, it = o.splat(int(cid))
            off = int(res["svl_offset"][k])
            assert np.array_equal(res["svl_items"][k], it)
            assert np.array_equal(splat[owner[k]][off:off + g.size * 2].view(np.int16), g)
            v, x = o.mesh(int(cid))
            vo, io = int(res["vbo_offset"][k]), int(res["ibo_offset"][k])
            assert np.array_equal(mesh[owner[k]][vo:vo + v.size * 2].view(np.int16), v)
            assert np.array_equal(mesh[owner[k]][io:io + x.size * 4].view(np.uint32), x)
    finally:
        m.close()


def test_multi_rejects_bad_device_counts():
    with pytest.raises(vpb.VoxplatError):
        vpb.MultiContext(4, (1, 1, 2), [0, 0, 0])              # 3 slabs do not divide 4 chunk rows
