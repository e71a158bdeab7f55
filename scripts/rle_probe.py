"""Device time of the RLE kernels on the C2 world (CUDA events via torch on the context stream)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
w = worldgen.World(1234, 6, (5, 2, 5))
ctx = vpb.Context(6, (5, 2, 5), rle_arena_bytes=1 << 30)
nn = w.nonnull_ids()
ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
words, offs = ctx.encode_chunks_rle(nn)
pw = torch.from_numpy(words).pin_memory()
import time
for name, fn in [("encode (device kernels + D2H of %.0f MB)" % (words.nbytes / 1e6), lambda: ctx.encode_chunks_rle(nn)),
                 ("upload_rle (H2D + decode + xfaces)", lambda: ctx.upload_chunks_rle(nn, pw, offs))]:
    for _ in range(3): fn()
    t0 = time.perf_counter()
    for _ in range(10): fn()
    print("%-45s %.3f ms per call" % (name, (time.perf_counter() - t0) * 100))
print("rle words", words.size, "bytes/voxel", words.nbytes / (len(nn) * w.N))
