"""LOD-node aggregation (vp_nodes.cu) against the restated gather of gfx_update_svl (oracle vo_lod_node)."""
import ctypes as C

import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
from test_gpu_splat import upload_world

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rb,bits", [(4, (2, 1, 3)), (5, (3, 1, 2)), (4, (0, 2, 1))])
def test_lod_nodes_match_oracle(rb, bits):
    w = worldgen.World(2718, rb, bits)
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
    ctx = vpb.Context(rb, bits)
    try:
        upload_world(ctx, w)
        ctx.rebuild_batch(np.arange(w.n_chunks, dtype=np.uint32), vpb.VP_REBUILD_SPLAT)
        for lod in range(5):
            nodes, buf, ms = ctx.build_lod_nodes(lod)
            for node in range(len(nodes)):
                want_n = lib.vo_lod_node(cbits, lod, node, ptrs, helpers.vp(items), None)
                assert nodes["items"][node] == want_n, (lod, node)
                if want_n:
                    want = np.zeros(want_n, np.int16)
                    lib.vo_lod_node(cbits, lod, node, ptrs, helpers.vp(items), helpers.vp(want))
                    off = int(nodes["offset"][node])
                    assert np.array_equal(buf[off:off + want_n * 2].view(np.int16), want), (lod, node)
            # every splat of the level lands in exactly one node
            assert int(nodes["items"].sum()) == int(items[:, lod].sum())
    finally:
        ctx.close()
