"""e2e (host RLE in -> host buffers out) of the C2 world for several pipeline block counts."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import voxplat_b200 as vpb
from voxplat_b200 import worldgen, slab
rb, bits = 6, (5, 2, 5)
w = worldgen.World(1234, rb, bits)
nn = w.nonnull_ids()
ctx = vpb.Context(rb, bits, splat_arena_bytes=2 << 30, mesh_arena_bytes=2 << 30, rle_arena_bytes=1 << 30)
ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
words, offs = ctx.encode_chunks_rle(nn)
pw = torch.from_numpy(words).pin_memory()
ids = np.arange(w.n_chunks, dtype=np.uint32)
near = slab.near_camera_flags(ids, rb, bits)
flags = np.where(near, 3, 1).astype(np.uint8)[nn]
for nb in [int(x) for x in os.environ.get("BLOCKS", "4,8,16,32,64").split(",")]:
    for _ in range(3):
        ctx.rebuild_from_rle(nn, pw, offs, per_chunk_flags=flags, n_blocks=nb)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.rebuild_from_rle(nn, pw, offs, per_chunk_flags=flags, n_blocks=nb)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    print("n_blocks %2d: %.2f ms  %.1f Gvoxel/s" % (nb, ms, w.n_chunks * w.N / ms / 1e6), flush=True)
