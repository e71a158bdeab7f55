import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np, helpers, voxplat_b200 as vpb
ctx = vpb.Context(5, (0, 0, 0))
rng = np.random.default_rng(3)
cases = [np.zeros(4096, np.uint8), np.full(32768, 7, np.uint8), np.arange(4096, dtype=np.uint32).astype(np.uint8),
         (np.arange(65536) // 3 % 256).astype(np.uint8), rng.integers(0, 2, 262144).astype(np.uint8),
         np.repeat(rng.integers(0, 256, 2048).astype(np.uint8), 1024)]
for k, d in enumerate(cases):
    t = time.time(); want = helpers.rle_encode(d); print(k, 'oracle enc', want.size, time.time() - t, flush=True)
    t = time.time(); enc = ctx.rle_compress(d); print(k, 'gpu enc', enc.size, np.array_equal(enc, want), time.time() - t, flush=True)
    t = time.time(); dec = ctx.rle_decompress(want, d.size); print(k, 'gpu dec', np.array_equal(dec, d), time.time() - t, flush=True)
