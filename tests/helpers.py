"""Test-side access to the checkers: the plain-C oracle restatement (oracle/libvoxoracle.so) and, when it
was built in this tree, the compiled unmodified reference (oracle/_ref/libvoxref.so).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline legs may import this."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libvoxoracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libvoxref.so")


class VoWorld(C.Structure):
    _fields_ = [("rb", C.c_int32), ("bits", C.c_int32 * 3), ("chunks", C.c_void_p), ("shadow", C.c_void_p)]


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.vo_chunk_splat.restype = C.c_uint32
        _oracle.vo_chunk_mesh_faces.restype = C.c_uint32
        _oracle.vo_rle_encode.restype = C.c_uint32
        _oracle.vo_rle_decode.restype = C.c_uint32
        _oracle.vo_fnv1a.restype = C.c_uint64
        _oracle.vo_fnv1a.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        _oracle.vo_world_rebuild.restype = C.c_double
    return _oracle


def ref_available():
    return os.path.exists(REF_SO)


def ref_lib():
    """The compiled reference (None when oracle/_ref was not built, e.g. /root/reference absent)."""
    global _ref
    if _ref is None and ref_available():
        _ref = C.CDLL(REF_SO)
        _ref.vr_world_create.restype = C.c_void_p
        _ref.vr_world_rebuild.restype = C.c_double
        _ref.vr_chunk_rle.restype = C.c_uint32
        _ref.vr_rle_compress.restype = C.c_uint32
        _ref.vr_fnv1a.restype = C.c_uint64
        _ref.vr_init(C.c_uint64(6 << 30))
        _ref.vr_set_scratch_scale(12)
    return _ref


def vp(a):
    return C.c_void_p(a.ctypes.data)


class OracleWorld:
    """vo_world view over a voxplat_b200.worldgen.World (dense chunks + padded shadow map)."""

    def __init__(self, world):
        self.world = world
        self.lib = oracle_lib()
        ptrs = world.chunk_ptrs()
        self._ptrs = (C.c_void_p * len(ptrs))(*[p if p else None for p in ptrs])
        self.w = VoWorld(world.root_bitw, (C.c_int32 * 3)(*world.max_bitw), C.cast(self._ptrs, C.c_void_p),
                         world.shadow.ctypes.data)
        self.R, self.N = world.R, world.N

    def splat(self, cid):
        out = np.zeros((self.R + 1) ** 3 * 5, np.int16)
        items = (C.c_uint32 * 5)()
        n = self.lib.vo_chunk_splat(C.byref(self.w), C.c_uint32(cid), vp(out), items)
        return out[:n].copy(), np.array(list(items), np.uint32)

    def mesh(self, cid):
        f = self.lib.vo_chunk_mesh_faces(C.byref(self.w), C.c_uint32(cid))
        vbo = np.zeros(max(f, 1) * 16, np.int16)
        ibo = np.zeros(max(f, 1) * 6, np.uint32)
        nv, ni = C.c_uint32(), C.c_uint32()
        self.lib.vo_chunk_mesh(C.byref(self.w), C.c_uint32(cid), vp(vbo), vp(ibo), C.byref(nv), C.byref(ni))
        assert nv.value == f * 16 and ni.value == f * 6
        return vbo[:nv.value].copy(), ibo[:ni.value].copy()

    def rebuild(self, ids, mode, nthreads=0):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        hashes = np.zeros(len(ids), np.uint64)
        counts = np.zeros((len(ids), 8), np.uint32)
        t = self.lib.vo_world_rebuild(C.byref(self.w), vp(ids), C.c_uint32(len(ids)), C.c_int(mode), C.c_int(nthreads),
                                      vp(hashes), vp(counts))
        return t, hashes, counts


def rle_encode(data):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    out = np.zeros(data.size + 1, np.uint32)
    n = oracle_lib().vo_rle_encode(vp(data), C.c_uint32(data.size), vp(out))
    return out[:n].copy()


def rle_decode(words, cap):
    words = np.ascontiguousarray(words, dtype=np.uint32)
    out = np.zeros(cap, np.uint8)
    n = oracle_lib().vo_rle_decode(vp(words), vp(out), C.c_uint32(cap))
    return out[:n].copy()


def fnv1a(a):
    a = np.ascontiguousarray(a)
    return int(oracle_lib().vo_fnv1a(C.c_void_p(a.ctypes.data), C.c_uint64(a.nbytes), C.c_uint64(0)))


class RefWorld:
    """The same world inside the compiled reference (struct ChunkSet built through its own API)."""

    def __init__(self, world):
        self.lib = ref_lib()
        assert self.lib is not None
        self.world = world
        self.set = C.c_void_p(self.lib.vr_world_create(world.root_bitw, *world.max_bitw))
        for i in range(world.n_chunks):
            if world.solid[i]:
                self.lib.vr_world_set_chunk(self.set, C.c_uint32(i), C.c_void_p(world.dense[i].ctypes.data))
        self.lib.vr_world_set_shadow(self.set, vp(world.shadow), C.c_uint32(world.shadow.size))
        self.R, self.N = world.R, world.N

    def splat(self, cid):
        out = np.zeros((self.R + 1) ** 3 * 5, np.int16)
        items = (C.c_uint32 * 5)()
        n = self.lib.vr_chunk_splat(self.set, C.c_uint32(cid), vp(out), C.c_uint32(out.size), items)
        return out[:n].copy(), np.array(list(items), np.uint32)

    def mesh(self, cid):
        vbo = np.zeros(self.N * 48, np.int16)
        ibo = np.zeros(self.N * 18, np.uint32)
        nv, ni = C.c_uint32(), C.c_uint32()
        self.lib.vr_chunk_mesh(self.set, C.c_uint32(cid), vp(vbo), C.c_uint32(vbo.size), vp(ibo), C.c_uint32(ibo.size),
                               C.byref(nv), C.byref(ni))
        return vbo[:nv.value].copy(), ibo[:ni.value].copy()

    def rebuild(self, ids, mode, nthreads=0):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        hashes = np.zeros(len(ids), np.uint64)
        counts = np.zeros((len(ids), 8), np.uint32)
        t = self.lib.vr_world_rebuild(self.set, vp(ids), C.c_uint32(len(ids)), C.c_int(mode), C.c_int(nthreads),
                                      vp(hashes), vp(counts))
        return t, hashes, counts


def random_world(seed, root_bitw, max_bitw, density=0.3, null_frac=0.2, maxv=255, shadow_random=True):
    """Adversarial (non-terrain) world: random voxels, some all-air chunks, random shadow map."""
    from voxplat_b200 import worldgen
    rng = np.random.default_rng(seed)
    n = 1 << sum(max_bitw)
    N = 1 << (3 * root_bitw)
    dense = np.zeros((n, N), np.uint8)
    for i in range(n):
        if rng.random() < null_frac:
            continue
        d = density if not isinstance(density, (list, tuple)) else density[i % len(density)]
        dense[i] = (rng.random(N) < d) * rng.integers(1, maxv + 1, N)
    # Reference UB guard: an air voxel at world (0,0,0) under a solid (0,0,1) makes chunk_make_mesh sample
    # shadow_map[(0 + 0-1) + 0] = shadow_map[0xFFFFFFFF] (mesher.c:315-316, shadow.h:60) -- the value cannot
    # change the result (the compare is `< 0`), but the load can fault.  Keep that voxel solid.
    dense[0, 0] = 1
    w = worldgen.World(seed, root_bitw, max_bitw, dense=dense)
    if shadow_random:
        Y = w.dims[1]
        w.shadow[:w.shw * w.dims[2]] = rng.integers(0, Y + 2, w.shw * w.dims[2])
    return w
