"""Config C5 (BASELINE.json): bursts of sphere edits (chunkset_edit_sphere, radius 4, alternating place 63 /
remove 0) at chunk sizes 32 and 128; per burst: re-upload the dirty chunks + shadow rows, rebuild them
(splat + mesh), copy the results to the host.  Reports ms per burst and Gvoxel/s over the dirty set.
The edit itself (voxel writes, shadow_place_update) stays host C in the reference (edit.c:179-244); here the
equivalent numpy edit is applied to the host copy of the world."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxplat_b200 as vpb
from voxplat_b200 import worldgen


def sphere_edit(w, cx, cy, cz, r, v):
    """Host-side edit with the reference's rule: |p - c| < r, y >= 2 (edit.c:151,221); returns dirty chunk ids
    = every chunk in the AABB (c-r-1 .. c+r+1) (edit.c:187-202,241-242)."""
    R, rb = w.R, w.root_bitw
    X, Y, Z = w.dims
    nx, ny = 1 << w.max_bitw[0], 1 << w.max_bitw[1]
    dirty = set()
    for gx in range((cx - r - 1) >> rb, ((cx + r + 1) >> rb) + 1):
        for gy in range((cy - r - 1) >> rb, ((cy + r + 1) >> rb) + 1):
            for gz in range((cz - r - 1) >> rb, ((cz + r + 1) >> rb) + 1):
                if 0 <= gx < nx and 0 <= gy < ny and 0 <= gz < (1 << w.max_bitw[2]):
                    dirty.add((gz * ny + gy) * nx + gx)
    for z in range(max(cz - r - 1, 0), min(cz + r + 1, Z)):
        for y in range(max(cy - r - 1, 2), min(cy + r + 1, Y)):
            for x in range(max(cx - r - 1, 0), min(cx + r + 1, X)):
                if (x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2 < r * r:
                    cid = (((z >> rb) * ny) + (y >> rb)) * nx + (x >> rb)
                    w.dense[cid, (((z & (R - 1)) << rb | (y & (R - 1))) << rb) | (x & (R - 1))] = v
                    if v:          # shadow_place_update (shadow.h:77-89)
                        idx = x + y + w.shw * z
                        if not (w.shadow[idx] >= y + 1 or w.shadow[idx + 1] >= y + 1):
                            w.shadow[idx] = y
    return np.array(sorted(dirty), np.uint32)


def run_device(rb, bits, bursts):
    """Same bursts with the edit applied on the device (vp_edit_sphere): no voxel upload at all."""
    w = worldgen.World(1234, rb, bits)
    ctx = vpb.Context(rb, bits, mesh_arena_bytes=1 << 30, splat_arena_bytes=1 << 30)
    nn = w.nonnull_ids()
    ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
    ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
    rng = np.random.default_rng(5)
    X, Y, Z = w.dims
    lat, nd = [], 0
    for b in range(bursts):
        x, z = int(rng.integers(8, X - 8)), int(rng.integers(8, Z - 8))
        h = int(worldgen.lib().vpw_height(__import__("ctypes").byref(worldgen.params(1234, rb, bits)), x, z))
        t0 = time.perf_counter()
        dirty = ctx.edit_sphere(x, h, z, 4, 63 if b % 2 == 0 else 0)
        res, splat, mesh = ctx.rebuild_batch(dirty, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
        lat.append((time.perf_counter() - t0) * 1e3)
        nd += len(dirty)
    ctx.close()
    lat = np.array(lat[5:])
    return {"chunk": 1 << rb, "mode": "edit on device", "ms_per_burst_median": float(np.median(lat)), "ms_per_burst_p95": float(np.percentile(lat, 95)),
            "dirty_chunks_per_burst": nd / bursts}


def run(rb, bits, bursts):
    w = worldgen.World(1234, rb, bits)
    ctx = vpb.Context(rb, bits, mesh_arena_bytes=1 << 30, splat_arena_bytes=1 << 30)
    nn = w.nonnull_ids()
    ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
    ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
    rng = np.random.default_rng(5)
    X, Y, Z = w.dims
    lat, vox = [], 0
    for b in range(bursts):
        x, z = int(rng.integers(8, X - 8)), int(rng.integers(8, Z - 8))
        h = int(worldgen.lib().vpw_height(__import__("ctypes").byref(worldgen.params(1234, rb, bits)), x, z))
        dirty = sphere_edit(w, x, h, z, 4, 63 if b % 2 == 0 else 0)
        z0, z1 = max(z - 6, 0), min(z + 6, Z)
        t0 = time.perf_counter()
        ctx.upload_chunks_dense(dirty, np.ascontiguousarray(w.dense[dirty]))
        ctx.upload_shadow_rows(z0, w.shadow[z0 * w.shw:z1 * w.shw])
        res, splat, mesh = ctx.rebuild_batch(dirty, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
        lat.append((time.perf_counter() - t0) * 1e3)
        vox += len(dirty) * w.N
    ctx.close()
    lat = np.array(lat[5:])
    return {"chunk": 1 << rb, "world": list(w.dims), "bursts": bursts, "ms_per_burst_median": float(np.median(lat)),
            "ms_per_burst_p95": float(np.percentile(lat, 95)), "dirty_chunks_per_burst": vox / w.N / bursts,
            "gvoxel_per_s_dirty_set": vox / bursts / (float(np.median(lat)) * 1e-3) / 1e9}


if __name__ == "__main__":
    out = [run(5, (4, 2, 4), 200), run(7, (3, 1, 3), 200), run_device(5, (4, 2, 4), 200), run_device(7, (3, 1, 3), 200)]
    print(json.dumps(out))
