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


if __name__ == "__main__":
    assert helpers.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    dump_kat()
    dump_world("terrain_r16", worldgen.World(1234, 4, (2, 1, 2)))
    dump_world("terrain_r32", worldgen.World(1234, 5, (1, 1, 1)))
    dump_world("random_r16", helpers.random_world(42, 4, (1, 1, 1), density=0.35, null_frac=0.25), with_rle=False)
