"""Config C1 (BASELINE.json): one 64^3 worldgen chunk at chunk coordinate (1,0,1) of the seed-1234 2048x256x2048
world: RLE decode + cull + 5 LOD splat lists + mesh + RLE encode.  GPU latency through the host-facing C ABI
(host RLE in, host buffers out) next to the reference's CPU code single-threaded (the README's 1-4 ms case)."""
import json, os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen

rb, bits = 6, (5, 2, 5)
w = worldgen.World(1234, rb, bits)
cid = (1 * 4 + 0) * 32 + 1
ctx = vpb.Context(rb, bits)
nn = w.nonnull_ids()
ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
ids = np.array([cid], np.uint32)
words, offs = ctx.encode_chunks_rle(ids)
o = helpers.OracleWorld(w)
lat = []
for i in range(60):
    t0 = time.perf_counter()
    ctx.upload_chunks_rle(ids, words, offs)                                      # H2D + rle_decompress on the device
    res, splat, mesh = ctx.rebuild_batch(ids, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)   # cull + LOD + splat + mesh, D2H
    enc, _ = ctx.encode_chunks_rle(ids)                                          # rle_compress on the device, D2H
    lat.append((time.perf_counter() - t0) * 1e3)
g, it = o.splat(cid); v, x = o.mesh(cid)
ok = (np.array_equal(splat[int(res["svl_offset"][0]):][:g.size * 2].view(np.int16), g) and
      np.array_equal(mesh[int(res["vbo_offset"][0]):][:v.size * 2].view(np.int16), v) and np.array_equal(enc, words))
out = {"chunk": [1, 0, 1], "gpu_ms_median": float(np.median(lat[10:])), "gpu_ms_min": float(np.min(lat[10:])), "parity": bool(ok),
       "splat_items": int(res["svl_items_total"][0]), "mesh_faces": int(res["ibo_items"][0]) // 6, "rle_words": int(words.size)}
if helpers.ref_available():
    r = helpers.RefWorld(w)
    lib = helpers.ref_lib()
    ts = []
    for i in range(5):
        t0 = time.perf_counter()
        r.splat(cid); r.mesh(cid)
        buf = np.zeros(w.N + 1, np.uint32)
        lib.vr_rle_compress(helpers.vp(w.dense[cid]), C.c_uint32(w.N), helpers.vp(buf), C.c_uint32(buf.size))
        ts.append((time.perf_counter() - t0) * 1e3)
    out["reference_cpu_ms_single_thread"] = float(np.median(ts))
print(json.dumps(out))
