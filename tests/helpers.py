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

    def raycast(self, origin, vector, normal=(0, 0, 0)):
        """Restated pick ray (vo_raycast): returns (voxel, coord[3], normal[3])."""
        o = (C.c_float * 3)(*[float(x) for x in origin])
        v = (C.c_float * 3)(*[float(x) for x in vector])
        c = (C.c_uint32 * 3)()
        m = (C.c_int8 * 3)(*normal)
        hit = self.lib.vo_raycast(C.byref(self.w), o, v, c, m)
        return hit, list(c), list(m)

    def rebuild(self, ids, mode, nthreads=0, hashed=True):
        """hashed=False: a timing run (no FNV pass over the outputs inside the clock); the hashes come back zero."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        hashes = np.zeros(len(ids), np.uint64)
        counts = np.zeros((len(ids), 8), np.uint32)
        t = self.lib.vo_world_rebuild(C.byref(self.w), vp(ids), C.c_uint32(len(ids)), C.c_int(mode), C.c_int(nthreads),
                                      vp(hashes) if hashed else None, vp(counts))
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
        for i in np.nonzero(world.solid)[0]:
            self.lib.vr_world_set_chunk(self.set, C.c_uint32(int(i)), C.c_void_p(world.dense[int(i)].ctypes.data))
        if hasattr(world, "shadow_pieces"):
            # a sparse world: only the rows it holds are written into the reference's (zero-filled, padded) map
            self.lib.vr_shadow_ptr.restype = C.c_void_p
            base = self.lib.vr_shadow_ptr(self.set)
            for z0, rows in world.shadow_pieces:
                C.memmove(base + z0 * world.shw * 2, rows.ctypes.data, rows.nbytes)
        else:
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

    def rebuild(self, ids, mode, nthreads=0, hashed=True):
        """hashed=False: a timing run (no FNV pass over the outputs inside the clock); the hashes come back zero."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        hashes = np.zeros(len(ids), np.uint64)
        counts = np.zeros((len(ids), 8), np.uint32)
        t = self.lib.vr_world_rebuild(self.set, vp(ids), C.c_uint32(len(ids)), C.c_int(mode), C.c_int(nthreads),
                                      vp(hashes) if hashed else None, vp(counts))
        return t, hashes, counts

    def run_engine_with_gfx(self, max_passes=600):
        """Drive the reference's OWN dispatcher (chunkset_manage, chunkset.c:246-507) until every chunk's splat list is
        published, handing each published chunk to the reference's gfx_update_svl (gfx/vsplat.c:197-338, compiled
        unmodified against the GL capture shim) the way the game loop does.  Afterwards node_buffer() reads what the
        reference would have left in GPU memory."""
        import time
        lib, n = self.lib, self.world.n_chunks
        for i in range(n):                                    # request a splat rebuild of every chunk (game.c:621-624)
            lib.vr_chunk_set_make_mesh(self.set, C.c_uint32(i), 1)
            lib.vr_chunk_set_make_mesh(self.set, C.c_uint32(i), 0)
        idle = 0
        for _ in range(max_passes):
            lib.vr_manage(self.set)
            new = 0
            for i in range(n):
                if lib.vr_chunk_svl_dirty(self.set, C.c_uint32(i)):
                    lib.vr_gfx_update_svl(self.set, C.c_uint32(i))
                    new += 1
            pending = sum(lib.vr_chunk_pending(self.set, C.c_uint32(i)) for i in range(n))
            idle = idle + 1 if (new == 0 and pending == 0) else 0
            if idle >= 2:
                return
            time.sleep(0.03)
        raise AssertionError("the reference dispatcher did not drain")

    def node_buffer(self, lod, node):
        """(int16 items, bytes) of the level-`lod` node buffer gfx_update_svl built (GeometrySVL.vbo_items + buffer data)."""
        data, size = C.c_void_p(), C.c_uint64()
        self.lib.vr_gsvl_node.restype = C.c_uint32
        items = self.lib.vr_gsvl_node(self.set, C.c_int(lod), C.c_uint32(node), C.byref(data), C.byref(size))
        if not items:
            return 0, np.zeros(0, np.int16)
        assert size.value == items * 2
        return items, np.frombuffer(C.string_at(data, size.value), np.int16)

    def raycast(self, origin, vector, normal=(0, 0, 0)):
        """chunkset_edit_raycast_until_solid (edit.c:248-314) as is: returns (voxel, coord[3], normal[3])."""
        o = (C.c_float * 3)(*[float(x) for x in origin])
        v = (C.c_float * 3)(*[float(x) for x in vector])
        c = (C.c_uint32 * 3)()
        m = (C.c_int8 * 3)(*normal)
        hit = self.lib.vr_raycast(self.set, o, v, c, m)
        return hit, list(c), list(m)

    def edit_sphere(self, x, y, z, r, v):
        """chunkset_edit_sphere (edit.c:179-244) on the reference's own ChunkSet."""
        self.lib.vr_edit_sphere(self.set, x, y, z, r, v)


class SparseWorld:
    """Some chunk rows (z rows of chunks) of a synthetic world held on the host; every other chunk is absent, i.e. air for
    whoever reads it.  This is what a parity check of a few chunk rows of a very large world needs (bench.py's parity
    guard): a chunk of row r has the reference's splat buffer when rows r and r+1 are held, and the reference's mesh
    buffer when rows r-1, r, r+1 are (rows outside the world count as held).  repeat_bits: the world is that smaller
    world repeated along z (bench.py's weak scaling)."""

    def __init__(self, seed, root_bitw, max_bitw, rows, repeat_bits=None):
        from voxplat_b200 import worldgen
        self.seed, self.root_bitw, self.max_bitw = seed, root_bitw, tuple(max_bitw)
        self.R = 1 << root_bitw
        self.N = self.R ** 3
        self.n_chunks = 1 << sum(max_bitw)
        self.dims = tuple((1 << b) * self.R for b in max_bitw)
        self.shw = self.dims[0] + self.dims[1]
        nz = 1 << max_bitw[2]
        per_row = 1 << (max_bitw[0] + max_bitw[1])
        self.rows = sorted(set(int(r) for r in rows if 0 <= r < nz))
        self.ids = np.concatenate([np.arange(r * per_row, (r + 1) * per_row, dtype=np.uint32) for r in self.rows])
        if repeat_bits is not None and tuple(repeat_bits) != tuple(max_bitw):
            base_rows = 1 << repeat_bits[2]
            gen_ids = (((self.ids // per_row) % base_rows) * per_row + self.ids % per_row).astype(np.uint32)
            self._dense, solid = worldgen.gen_chunks(seed, root_bitw, repeat_bits, gen_ids)
        else:
            self._dense, solid = worldgen.gen_chunks(seed, root_bitw, max_bitw, self.ids)
        self.solid = np.zeros(self.n_chunks, np.uint32)
        self.solid[self.ids] = solid
        self._pos = np.full(self.n_chunks, -1, np.int64)
        self._pos[self.ids] = np.arange(len(self.ids))
        ptrs = self.chunk_ptrs()
        # the height map: rows of the held chunk rows only (a voxel row z depends on the chunks of chunk row z / R alone)
        self.shadow = np.zeros(self.shw * self.dims[2] + worldgen.shadow_pad(root_bitw, max_bitw), np.uint16)   # untouched pages cost nothing
        self.shadow_pieces = []
        for r in self.rows:
            rows_r = worldgen.shadow_rows(seed, root_bitw, max_bitw, ptrs, r * self.R, (r + 1) * self.R)
            self.shadow[r * self.R * self.shw:(r + 1) * self.R * self.shw] = rows_r
            self.shadow_pieces.append((r * self.R, rows_r))

    class _Dense:
        def __init__(self, w):
            self.w = w

        def __getitem__(self, i):
            k = self.w._pos[i]
            assert k >= 0, "chunk %d is not held by this sparse world" % i
            return self.w._dense[k]

    @property
    def dense(self):
        return SparseWorld._Dense(self)

    def chunk_ptrs(self):
        base = self._dense.ctypes.data
        ptrs = [0] * self.n_chunks
        for k, cid in enumerate(self.ids):
            if self.solid[cid]:
                ptrs[int(cid)] = base + k * self.N
        return ptrs

    def checkable(self, mesh):
        """Chunk rows whose reference buffers this world can produce."""
        nz = 1 << self.max_bitw[2]
        have = set(self.rows)
        ok = lambda r: r < 0 or r >= nz or r in have
        return [r for r in self.rows if ok(r + 1) and (not mesh or ok(r - 1))]


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
