"""Generate the golden fixtures from the COMPILED, UNMODIFIED reference (oracle/_ref/libvoxref.so).
Run in the container that has /root/reference:   python tests/golden/make_golden.py
Fixtures hold the input world (dense voxels + shadow map) and the reference's outputs for every chunk:
splat buffer + svl_items[5], mesh VBO/IBO, RLE stream.  They are small (tens of KB, npz-compressed)."""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from voxplat_b200 import worldgen  # noqa: E402


def dump_world(name, w, with_rle=True):
    r = helpers.RefWorld(w)
    lib = helpers.ref_lib()
    out = {"root_bitw": w.root_bitw, "max_bitw": np.array(w.max_bitw), "dense": w.dense, "shadow": w.shadow}
    splat, items, vbo, ibo, nv, ni, rle, rle_off = [], [], [], [], [], [], [], [0]
    for cid in range(w.n_chunks):
        g, it = r.splat(cid)
        splat.append(g); items.append(it)
        v, x = r.mesh(cid)
        vbo.append(v); ibo.append(x); nv.append(v.size); ni.append(x.size)
        if with_rle:      # random data has more than N/4 runs: the reference's encoder scratch would overflow (rle.c:49)
            buf = np.zeros(w.N + 1, np.uint32)
            k = lib.vr_rle_compress(helpers.vp(w.dense[cid]), C.c_uint32(w.N), helpers.vp(buf), C.c_uint32(buf.size))
            rle.append(buf[:k].copy()); rle_off.append(rle_off[-1] + k)
    out.update(splat=np.concatenate(splat), svl_items=np.array(items, np.uint32), vbo=np.concatenate(vbo), ibo=np.concatenate(ibo),
               vbo_items=np.array(nv, np.uint32), ibo_items=np.array(ni, np.uint32), rle=np.concatenate(rle) if rle else np.zeros(0, np.uint32),
               rle_offsets=np.array(rle_off, np.uint64))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "chunks", w.n_chunks, "splat items", out["splat"].size, "faces", out["ibo"].size // 6, "rle words", out["rle"].size)


def dump_kat():
    """SURVEY 8(c) known-answer world: 8x4x4 voxels as 2x1x1 chunks of 4^3, three voxels set through
    chunkset_edit_write; the reference's outputs verbatim."""
    lib = helpers.ref_lib()
    s = C.c_void_p(lib.vr_world_create(2, 1, 0, 0))
    writes = [(1, 1, 1, 5), (4, 1, 1, 9), (3, 3, 3, 33)]
    for x, y, z, v in writes:
        lib.vr_edit_write(s, x, y, z, v)
    kat = {"root_bitw": 2, "max_bitw": [1, 0, 0], "writes": writes, "chunks": []}
    for cid in (0, 1):
        g = np.zeros(4096, np.int16); items = (C.c_uint32 * 5)()
        n = lib.vr_chunk_splat(s, cid, helpers.vp(g), 4096, items)
        v = np.zeros(4096, np.int16); x = np.zeros(4096, np.uint32); nv, ni = C.c_uint32(), C.c_uint32()
        lib.vr_chunk_mesh(s, cid, helpers.vp(v), 4096, helpers.vp(x), 4096, C.byref(nv), C.byref(ni))
        kat["chunks"].append({"svl_items": list(items), "svl": g[:n].tolist(), "vbo": v[:nv.value].tolist(), "ibo": x[:ni.value].tolist()})
    d = np.array([0, 0, 0, 5, 5, 7] + [0] * 9 + [9], np.uint8)
    buf = np.zeros(32, np.uint32)
    k = lib.vr_rle_compress(helpers.vp(d), C.c_uint32(16), helpers.vp(buf), C.c_uint32(32))
    kat["rle"] = {"data": d.tolist(), "words": buf[:k].tolist()}
    json.dump(kat, open(os.path.join(HERE, "kat_r4.json"), "w"))
    print("kat", kat["chunks"][0]["svl_items"], kat["rle"]["words"])


def load_world(name):
    z = np.load(os.path.join(HERE, name + ".npz"))
    w = worldgen.World(0, int(z["root_bitw"]), tuple(int(b) for b in z["max_bitw"]), dense=z["dense"])
    w.shadow[:] = z["shadow"]
    return w


def dump_rays(name, world_fixture, n=600, seed=7):
    """Pick rays (chunkset_edit_raycast_until_solid, edit.c:248-314) on the world of an existing fixture: random, camera-like,
    axis-aligned, starting outside, leaving through a 0-face, on cell corners, exact diagonals."""
    w = load_world(world_fixture)
    r = helpers.RefWorld(w)
    rng = np.random.default_rng(seed)
    X, Y, Z = w.dims
    o = np.stack([rng.uniform(0, X, n), rng.uniform(0, Y, n), rng.uniform(0, Z, n)], axis=1).astype(np.float32)
    v = rng.normal(size=(n, 3)).astype(np.float32)
    k = n // 10
    o[:k, 1] = Y - 1.5
    v[:k, 1] = -np.abs(v[:k, 1]) - 0.2
    v[k:k + 10] = [0, -1, 0]
    v[k + 10:k + 20] = [1, 0, 0]
    v[k + 20:k + 30] = [0, 0, -1]
    v[k + 30:2 * k, 2] = 0
    o[2 * k:3 * k, 0] = X + rng.uniform(1, 20, k).astype(np.float32)
    v[2 * k:3 * k, 0] = -np.abs(v[2 * k:3 * k, 0]) - 0.1
    o[3 * k:4 * k, 2] = -rng.uniform(1, 20, k).astype(np.float32)
    v[3 * k:4 * k, 2] = np.abs(v[3 * k:4 * k, 2]) + 0.1
    o[4 * k:5 * k, 0] = rng.uniform(0, 3, k).astype(np.float32)
    v[4 * k:5 * k] = [-1, 0.01, 0.02]
    o[5 * k:6 * k] = np.floor(o[5 * k:6 * k])
    v[6 * k:7 * k] = np.sign(v[6 * k:7 * k])
    vox, coords, nrm = np.zeros(n, np.uint8), np.zeros((n, 3), np.uint32), np.zeros((n, 3), np.int8)
    for i in range(n):
        vox[i], coords[i], nrm[i] = r.raycast(o[i], v[i])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), world=world_fixture, origins=o, vectors=v, voxels=vox, coords=coords, normals=nrm)
    print(name, "rays", n, "hits", int((vox > 0).sum()))


def dump_nodes(name, world_fixture):
    """LOD-node buffers as the reference's own gfx_update_svl (gfx/vsplat.c:197-338, GL calls captured) leaves them, for
    every node of every level, after its own dispatcher published every chunk."""
    w = load_world(world_fixture)
    r = helpers.RefWorld(w)
    r.run_engine_with_gfx()
    items, bufs = [], []
    for lod in range(5):
        for node in range(1 << sum(b - min(lod, b) for b in w.max_bitw)):
            n, buf = r.node_buffer(lod, node)
            items.append((lod, node, n))
            bufs.append(buf)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), world=world_fixture, nodes=np.array(items, np.uint32),
                        data=np.concatenate(bufs) if bufs else np.zeros(0, np.int16))
    print(name, "nodes", len(items), "items", int(sum(i[2] for i in items)))


def dump_edits(name, world_fixture, n=24, seed=11):
    """A burst of chunkset_edit_sphere calls (edit.c:179-244): the edits, the reference's dirty list after each, and the final
    voxels + height map."""
    w = load_world(world_fixture)
    r = helpers.RefWorld(w)
    lib = r.lib
    lib.vr_chunk_voxels.restype = C.c_void_p
    lib.vr_shadow_ptr.restype = C.c_void_p
    for i in range(w.n_chunks):
        lib.vr_chunk_dirty(r.set, C.c_uint32(i), 1)
    rng = np.random.default_rng(seed)
    X, Y, Z = w.dims
    edits = [(0, 3, 0, 3, 9), (X - 1, Y - 2, Z - 1, 4, 7), (w.R, 1, w.R, 5, 63)]
    edits += [(int(rng.integers(0, X)), int(rng.integers(0, min(Y, 60))), int(rng.integers(0, Z)), int(rng.integers(1, 7)), int(rng.choice([0, 63, 17])))
              for _ in range(n - len(edits))]
    dirty, dirty_off = [], [0]
    for (x, y, z, rad, v) in edits:
        r.edit_sphere(x, y, z, rad, v)
        d = [i for i in range(w.n_chunks) if lib.vr_chunk_dirty(r.set, C.c_uint32(i), 1)]
        dirty += d
        dirty_off.append(len(dirty))
    dense = np.stack([np.frombuffer(C.string_at(lib.vr_chunk_voxels(r.set, C.c_uint32(i)), w.N), np.uint8) for i in range(w.n_chunks)])
    shadow = np.frombuffer(C.string_at(lib.vr_shadow_ptr(r.set), w.shw * Z * 2), np.uint16)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), world=world_fixture, edits=np.array(edits, np.int32), dirty=np.array(dirty, np.uint32),
                        dirty_offsets=np.array(dirty_off, np.uint32), dense=dense, shadow=shadow)
    print(name, "edits", len(edits), "dirty entries", len(dirty))


if __name__ == "__main__":
    assert helpers.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    dump_kat()
    dump_world("terrain_r16", worldgen.World(1234, 4, (2, 1, 2)))
    dump_world("terrain_r32", worldgen.World(1234, 5, (1, 1, 1)))
    dump_world("random_r16", helpers.random_world(42, 4, (1, 1, 1), density=0.35, null_frac=0.25), with_rle=False)
    dump_secondary()


def dump_secondary():
    """Fixtures on top of the worlds above (the world fixtures themselves are not regenerated by this entry point)."""
    dump_rays("rays_terrain_r32", "terrain_r32")
    dump_rays("rays_random_r16", "random_r16", seed=8)
    dump_nodes("nodes_terrain_r16", "terrain_r16")
    dump_edits("edits_terrain_r32", "terrain_r32")
