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
