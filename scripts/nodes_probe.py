"""LOD-node aggregation on the C2 world: device time per level and achieved copy bandwidth."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
w = worldgen.World(1234, 6, (5, 2, 5))
ctx = vpb.Context(6, (5, 2, 5), splat_arena_bytes=1 << 30)
nn = w.nonnull_ids()
ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
ctx.rebuild_batch(np.arange(w.n_chunks, dtype=np.uint32), vpb.VP_REBUILD_SPLAT)
out = []
for lod in range(5):
    best = 1e9
    for _ in range(5):
        nodes, _, ms = ctx.build_lod_nodes(lod, download=False)
        best = min(best, ms)
    b = int(nodes["items"].sum()) * 2
    out.append({"lod": lod, "nodes": len(nodes), "non_empty": int((nodes["items"] > 0).sum()), "bytes": b, "kernel_ms": round(best, 4),
                "copy_GBps_read_plus_write": round(2 * b / best / 1e6, 1)})
print(json.dumps(out))
