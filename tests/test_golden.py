"""Oracle restatement against the committed golden fixtures (generated from the compiled reference by
tests/golden/make_golden.py) -- runs everywhere, no GPU and no /root/reference needed."""
import json
import os

import numpy as np
import pytest

import helpers
from voxplat_b200 import worldgen

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["terrain_r16", "terrain_r32", "random_r16"]


def load_fixture(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    w = worldgen.World(0, int(z["root_bitw"]), tuple(int(b) for b in z["max_bitw"]), dense=z["dense"])
    w.shadow[:] = z["shadow"]
    return w, z


def expected_chunks(w, z):
    """Per chunk: (splat int16, items[5], vbo int16, ibo uint32) slices of the fixture."""
    so = np.concatenate([[0], np.cumsum(z["svl_items"].sum(axis=1))]).astype(np.int64)
    vo = np.concatenate([[0], np.cumsum(z["vbo_items"])]).astype(np.int64)
    io = np.concatenate([[0], np.cumsum(z["ibo_items"])]).astype(np.int64)
    for c in range(w.n_chunks):
        yield c, z["splat"][so[c]:so[c + 1]], z["svl_items"][c], z["vbo"][vo[c]:vo[c + 1]], z["ibo"][io[c]:io[c + 1]]


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_golden(name):
    w, z = load_fixture(name)
    o = helpers.OracleWorld(w)
    for c, splat, items, vbo, ibo in expected_chunks(w, z):
        g, it = o.splat(c)
        assert np.array_equal(it, items) and np.array_equal(g, splat), c
        v, x = o.mesh(c)
        assert np.array_equal(v, vbo) and np.array_equal(x, ibo), c
    if z["rle"].size:
        offs = z["rle_offsets"].astype(np.int64)
        for c in range(w.n_chunks):
            words = z["rle"][offs[c]:offs[c + 1]]
            assert np.array_equal(helpers.rle_encode(w.dense[c]), words)
            assert np.array_equal(helpers.rle_decode(words, w.N), w.dense[c])


def test_known_answer_vectors():
    """8x4x4 world of two 4^3 chunks, three voxels; RLE KAT."""
    kat = json.load(open(os.path.join(GOLD, "kat_r4.json")))
    dense = np.zeros((2, 64), np.uint8)
    for x, y, z, v in kat["writes"]:
        dense[x // 4, (z * 4 + y) * 4 + x % 4] = v
    w = worldgen.World(0, 2, (1, 0, 0), dense=dense)
    w.shadow[:] = 0                                   # the reference run had an all-zero shadow map
    o = helpers.OracleWorld(w)
    for c, exp in enumerate(kat["chunks"]):
        g, it = o.splat(c)
        assert list(it) == exp["svl_items"] and g.tolist() == exp["svl"]
        v, x = o.mesh(c)
        assert v.tolist() == exp["vbo"] and x.tolist() == exp["ibo"]
    # spot values quoted in the survey
    c0 = kat["chunks"][0]
    assert c0["svl_items"] == [12, 12, 8, 4, 4]
    assert c0["svl"][:12] == [1, 1, 1, 5, 4, 1, 1, 9, 3, 3, 3, 33]
    assert len(c0["vbo"]) == 208 and len(c0["ibo"]) == 78 and c0["ibo"][:6] == [1, 3, 0, 3, 1, 2]
    assert [v & 0xFFFF for v in c0["vbo"][:4]] == [2, 1, 1, 0xC505]
    d = np.array(kat["rle"]["data"], np.uint8)
    assert helpers.rle_encode(d).tolist() == kat["rle"]["words"] == [3, 0x05000002, 0x07000001, 9, 0x09000001, 0]


# ---- fixtures of the rows next to the path (rays, LOD nodes, edits), generated from the compiled reference too ----
def load_secondary(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    w, _ = load_fixture(str(z["world"]))
    return w, z


@pytest.mark.parametrize("name", ["rays_terrain_r32", "rays_random_r16"])
def test_restated_raycast_matches_golden(name):
    w, z = load_secondary(name)
    o = helpers.OracleWorld(w)
    for i in range(len(z["voxels"])):
        hit, coord, normal = o.raycast(z["origins"][i], z["vectors"][i])
        assert (hit, coord, normal) == (int(z["voxels"][i]), z["coords"][i].tolist(), z["normals"][i].tolist()), i


def test_restated_node_gather_matches_golden():
    import ctypes as C
    w, z = load_secondary("nodes_terrain_r16")
    o = helpers.OracleWorld(w)
    lib = helpers.oracle_lib()
    lib.vo_lod_node.restype = C.c_uint32
    svl, items = [], np.zeros((w.n_chunks, 5), np.uint32)
    for c in range(w.n_chunks):
        g, it = o.splat(c)
        svl.append(np.ascontiguousarray(g if g.size else np.zeros(4, np.int16)))
        items[c] = it
    ptrs = (C.c_void_p * w.n_chunks)(*[a.ctypes.data for a in svl])
    cbits = (C.c_int32 * 3)(*w.max_bitw)
    at = 0
    for lod, node, n in z["nodes"]:
        got = np.zeros(max(int(n), 1), np.int16)
        assert lib.vo_lod_node(cbits, int(lod), int(node), ptrs, helpers.vp(items), helpers.vp(got)) == n, (lod, node)
        assert np.array_equal(got[:n], z["data"][at:at + n]), (lod, node)
        at += int(n)
    assert at == z["data"].size


def test_host_stand_in_edits_match_golden():
    """tests/hoststore.HostContext.edit_sphere (the stand-in the CPU dry runs use) against the reference's edit burst."""
    import hoststore
    w, z = load_secondary("edits_terrain_r32")
    ctx = hoststore.HostContext(w.root_bitw, w.max_bitw)
    nn = w.nonnull_ids()
    ctx.upload_chunks_dense(nn, w.dense[nn])
    ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
    offs = z["dirty_offsets"]
    for k, (x, y, zz, r, v) in enumerate(z["edits"].tolist()):
        dirty = ctx.edit_sphere(x, y, zz, r, v)
        assert sorted(dirty.tolist()) == z["dirty"][offs[k]:offs[k + 1]].tolist(), k
    assert np.array_equal(ctx.download_chunks_dense(np.arange(w.n_chunks)), z["dense"])
    assert np.array_equal(ctx.download_shadow_rows(0, w.dims[2]), z["shadow"])
