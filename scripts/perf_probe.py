"""Quick device-side timing of the rebuild kernels on the C2 world (2048x256x2048, 64^3 chunks)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import voxplat_b200 as vpb
from voxplat_b200 import worldgen

rb = int(os.environ.get("RB", 6))
bits = tuple(int(x) for x in os.environ.get("BITS", "5,2,5").split(","))
flags = int(os.environ.get("FLAGS", 1))
steps = int(os.environ.get("STEPS", 10))
t0 = time.time()
w = worldgen.World(1234, rb, bits)
print("world %s gen %.1fs nonnull %d/%d" % (w.dims, time.time() - t0, len(w.nonnull_ids()), w.n_chunks), flush=True)
ctx = vpb.Context(rb, bits, splat_arena_bytes=3 << 30, mesh_arena_bytes=(6 << 30) if flags & 2 else (16 << 20))
nn = w.nonnull_ids()
t0 = time.time()
ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
print("upload %.2fs" % (time.time() - t0), flush=True)
ids = np.arange(w.n_chunks, dtype=np.uint32)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
ctx.batch_prepare(ids, flags)
for _ in range(3):
    ctx.rebuild_device()
torch.cuda.synchronize()
res, sb, mb = ctx.rebuild_device_results()
vox = w.n_chunks * w.N
nn_vox = len(nn) * w.N
splats = int(res["svl_items_total"].sum()) // 4
faces = int(res["vbo_items"].sum()) // 16
alg = nn_vox + 3 * len(nn) * w.R ** 2 + 12 * splats + (64 * faces)
print("splats %d (%.1f MB) faces %d (%.1f MB) alg bytes %.1f MB" % (splats, sb / 1e6, faces, mb / 1e6, alg / 1e6))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record(stream)
for i in range(steps):
    ctx.rebuild_device()
    ev[i + 1].record(stream)
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
m = float(np.median(ms))
print("ms/step median %.3f min %.3f max %.3f" % (m, min(ms), max(ms)))
print("Gvoxel/s %.1f (all chunks) ; achieved %.1f GB/s algorithmic = %.1f%% of 6527.8" % (vox / m / 1e6, alg / m / 1e6, alg / m / 1e6 / 65.278))
