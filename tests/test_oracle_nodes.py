"""CPU check of the node-gather restatement on the SURVEY 8(c) known-answer world (two 4^3 chunks)."""
import ctypes as C
import json
import os

import numpy as np

import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_node_gather_on_kat_world():
    kat = json.load(open(os.path.join(GOLD, "kat_r4.json")))
    lib = helpers.oracle_lib()
    lib.vo_lod_node.restype = C.c_uint32
    svl = [np.array(c["svl"], np.int16) for c in kat["chunks"]]
    items = np.array([c["svl_items"] for c in kat["chunks"]], np.uint32)
    ptrs = (C.c_void_p * 2)(*[a.ctypes.data for a in svl])
    bits = (C.c_int32 * 3)(1, 0, 0)
    # lod 0: one node per chunk = its level-0 segment
    for node in (0, 1):
        out = np.zeros(64, np.int16)
        n = lib.vo_lod_node(bits, 0, node, ptrs, helpers.vp(items), helpers.vp(out))
        assert n == items[node, 0] and out[:n].tolist() == svl[node][:n].tolist()
    # lod 1: both chunks fall into node 0 (x outer): chunk 0's level-1 segment, then chunk 1's
    out = np.zeros(64, np.int16)
    n = lib.vo_lod_node(bits, 1, 0, ptrs, helpers.vp(items), helpers.vp(out))
    assert n == items[0, 1] + items[1, 1]
    assert out[:n].tolist() == svl[0][12:24].tolist() + svl[1][4:8].tolist()


import pytest  # noqa: E402


@pytest.mark.skipif(not helpers.ref_available(), reason="oracle/_ref/libvoxref.so not built")
@pytest.mark.parametrize("rb,bits", [(4, (2, 1, 2)), (4, (0, 2, 1)), (5, (2, 1, 1))])
def test_node_gather_pinned_to_the_reference_gfx_update_svl(rb, bits):
    """SURVEY 8(f) f2 pinned: the reference's own dispatcher publishes every chunk, its own gfx_update_svl (compiled
    unmodified, GL calls captured) gathers the node buffers; the restated gather vo_lod_node must give the same bytes for
    every node of every level."""
    from voxplat_b200 import worldgen
    w = worldgen.World(2718, rb, bits)
    r = helpers.RefWorld(w)
    r.run_engine_with_gfx()
    o = helpers.OracleWorld(w)
    lib = helpers.oracle_lib()
    lib.vo_lod_node.restype = C.c_uint32
    svl, items = [], np.zeros((w.n_chunks, 5), np.uint32)
    for c in range(w.n_chunks):
        g, it = o.splat(c)
        svl.append(np.ascontiguousarray(g if g.size else np.zeros(4, np.int16)))
        items[c] = it
    ptrs = (C.c_void_p * w.n_chunks)(*[a.ctypes.data for a in svl])
    cbits = (C.c_int32 * 3)(*bits)
    seen = 0
    for lod in range(5):
        n_nodes = 1 << sum(b - min(lod, b) for b in bits)
        for node in range(n_nodes):
            want_n, want = r.node_buffer(lod, node)
            got_n = lib.vo_lod_node(cbits, lod, node, ptrs, helpers.vp(items), None)
            assert got_n == want_n, (lod, node)
            if want_n:
                got = np.zeros(got_n, np.int16)
                lib.vo_lod_node(cbits, lod, node, ptrs, helpers.vp(items), helpers.vp(got))
                assert np.array_equal(got, want), (lod, node)
                seen += 1
    assert seen >= 5
