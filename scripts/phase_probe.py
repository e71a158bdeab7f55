"""Per-phase CTA time of k_splat (instrumented build: VP_NVCC_EXTRA=-DVP_PROFILE_PHASES python voxplat_b200/build.py --force)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
w = worldgen.World(1234, 6, (5, 2, 5))
ctx = vpb.Context(6, (5, 2, 5), splat_arena_bytes=3 << 30)
nn = w.nonnull_ids()
ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
ctx.batch_prepare(np.arange(w.n_chunks, dtype=np.uint32), 1)
lib = vpb.load_library()
out = (C.c_ulonglong * 8)()
for _ in range(3): ctx.rebuild_device()
lib.vp_debug_phase_cycles(out, 1)
for _ in range(5): ctx.rebuild_device()
lib.vp_debug_phase_cycles(out, 1)
v = np.array(list(out), dtype=np.float64)
names = ["0 init/zero", "1 stream+bits", "2 zero lv + vis", "3 LOD", "4 counts+scan", "5 cluster exchange", "6 emission", "-"]
for n, x in zip(names, v): print("%-22s %6.1f%%  %8.0f cycles/CTA" % (n, 100 * x / v.sum(), x / 5 / (int(os.environ.get("CL", 4)) * len(nn))))
