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
