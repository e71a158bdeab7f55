"""Device world generation vs host generation + upload on the C2 world (2048x256x2048, 64^3 chunks)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
rb, bits = 6, (5, 2, 5)
ctx = vpb.Context(rb, bits)
ctx.generate_world(1234); torch.cuda.synchronize()
t0 = time.perf_counter(); ctx.generate_world(1234); torch.cuda.synchronize(); t_dev = time.perf_counter() - t0
t0 = time.perf_counter(); w = worldgen.World(1234, rb, bits); t_host = time.perf_counter() - t0
nn = w.nonnull_ids()
t0 = time.perf_counter(); ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn])); ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]]); t_up = time.perf_counter() - t0
print("device generate_world %.1f ms ; host generator (%d threads) %.1f ms + upload %.1f ms" % (t_dev * 1e3, os.cpu_count(), t_host * 1e3, t_up * 1e3))
tmp = "/tmp/vp_world.bin"
t0 = time.perf_counter(); n = ctx.save_world(tmp); t_s = time.perf_counter() - t0
t0 = time.perf_counter(); ctx.load_world(tmp); t_l = time.perf_counter() - t0
print("world file %.1f MB: save %.1f ms, load %.1f ms" % (n / 1e6, t_s * 1e3, t_l * 1e3))
os.remove(tmp)
