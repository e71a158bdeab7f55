"""TEST INFRASTRUCTURE: a host stand-in for voxplat_b200.Context, backed by the oracle restatement, with the same
method surface bench.py and voxplat_b200.slab use.  It exists so that the multi-rank HOST logic -- the slab split, the
border-plane exchange, the order of the collectives, the parity guard's bookkeeping, the JSON assembly -- can run on CPU
under a gloo process group (tests/test_bench_dryrun.py).  It is never imported by the product or by bench.py: the product
has no CPU path.

Like the device context it holds only its slab's chunks plus one ghost chunk row on each side, and a ghost row holds only
the border slice the neighbour sent (halo_unpack): a wrong exchange gives wrong buffers, exactly as on the GPUs."""
import ctypes as C

import numpy as np

import helpers
from voxplat_b200 import api as vapi, worldgen


class HostContext:
    def __init__(self, root_bitw, max_bitw, device=0, slab=None, **_arenas):
        self.rb, self.bits = root_bitw, tuple(max_bitw)
        self.R, self.N = 1 << root_bitw, 1 << (3 * root_bitw)
        self.nx, self.ny, self.nz = (1 << b for b in self.bits)
        self.per_row = self.nx * self.ny
        self.z0, self.z1 = slab if slab is not None else (0, self.nz)
        self.n_chunks = self.per_row * self.nz
        self.chunks = {}                         # id -> dense bytes (own rows and ghost rows)
        self.shw = (self.nx + self.ny) * self.R
        self.shadow = np.zeros(self.shw * self.nz * self.R + worldgen.shadow_pad(root_bitw, max_bitw), np.uint16)
        self.batch = None
        self.version, self.cache = 0, {}
        self.launches = 0
        self.last = None

    # ---- residency ----
    def set_stream(self, handle):
        pass

    def close(self):
        self.chunks.clear()

    def _own(self, cid):
        return self.z0 <= cid // self.per_row < self.z1

    def upload_chunks_dense(self, ids, dense):
        dense = np.asarray(dense).reshape(len(ids), self.N)
        for k, cid in enumerate(np.asarray(ids)):
            assert self._own(int(cid))
            if dense[k].any():
                self.chunks[int(cid)] = np.array(dense[k], np.uint8)
            else:
                self.chunks.pop(int(cid), None)
        self.version += 1

    def upload_chunks_rle(self, ids, words, word_offsets):
        words = np.asarray(words)
        for k, cid in enumerate(np.asarray(ids)):
            d = helpers.rle_decode(np.ascontiguousarray(words[int(word_offsets[k]):int(word_offsets[k + 1])]), self.N)
            assert d.size == self.N
            if d.any():
                self.chunks[int(cid)] = d
            else:
                self.chunks.pop(int(cid), None)
        self.version += 1

    def encode_chunks_rle(self, ids):
        streams = [helpers.rle_encode(self.chunks[int(c)]) if int(c) in self.chunks else np.array([self.N, 0], np.uint32) for c in ids]
        offs = np.zeros(len(ids) + 1, np.uint64)
        offs[1:] = np.cumsum([s.size for s in streams])
        return (np.concatenate(streams) if streams else np.zeros(0, np.uint32)), offs

    def chunks_resident(self, ids):
        return np.array([int(c) in self.chunks for c in ids], np.uint8)

    def upload_shadow_rows(self, z0, rows):
        rows = np.asarray(rows).reshape(-1)
        self.shadow[z0 * self.shw:z0 * self.shw + rows.size] = rows
        self.version += 1

    upload_shadow_rows_async = upload_shadow_rows

    def download_shadow_rows(self, z0, z1):
        return self.shadow[z0 * self.shw:z1 * self.shw].copy()

    def generate_world(self, seed):
        ids = np.arange(self.z0 * self.per_row, self.z1 * self.per_row, dtype=np.uint32)
        more = np.arange(self.z0 * self.per_row, min(self.z1 + 1, self.nz) * self.per_row, dtype=np.uint32)
        dense, solid = worldgen.gen_chunks(seed, self.rb, self.bits, more)
        ptrs = [0] * self.n_chunks
        for k, cid in enumerate(more):
            if solid[k]:
                ptrs[int(cid)] = dense.ctypes.data + k * self.N
        sz0, sz1 = self.z0 * self.R, min(self.nz * self.R, self.z1 * self.R + 17)
        self.upload_shadow_rows(sz0, worldgen.shadow_rows(seed, self.rb, self.bits, ptrs, sz0, sz1))
        self.upload_chunks_dense(ids, dense[:len(ids)])

    def edit_sphere(self, cx, cy, cz, r, v):
        """chunkset_edit_sphere + shadow_place_update (edit.c:179-244, shadow.h:77-89) on the host copy; returns the dirty list."""
        rb, R = self.rb, self.R
        X, Y, Z = self.nx * R, self.ny * R, self.nz * R
        dirty = []
        for gx in range((cx - r - 1) >> rb, ((cx + r + 1) >> rb) + 1):
            for gy in range((cy - r - 1) >> rb, ((cy + r + 1) >> rb) + 1):
                for gz in range((cz - r - 1) >> rb, ((cz + r + 1) >> rb) + 1):
                    if 0 <= gx < self.nx and 0 <= gy < self.ny and 0 <= gz < self.nz:
                        dirty.append((gz * self.ny + gy) * self.nx + gx)
        for cid in dirty:                                      # chunk_open_rw gives a null chunk its own zeroed voxels
            if self._own(cid) and cid not in self.chunks:
                self.chunks[cid] = np.zeros(self.N, np.uint8)
        for gx in sorted({d % self.nx for d in dirty}):        # the reference's visiting order: chunks x, y, z; cells x, y, z
            for gy in sorted({(d // self.nx) % self.ny for d in dirty}):
                for gz in sorted({d // self.per_row for d in dirty}):
                    for x in range(max(cx - r - 1, gx << rb), min(cx + r + 1, (gx + 1) << rb)):
                        for y in range(max(cy - r - 1, gy << rb, 2), min(cy + r + 1, (gy + 1) << rb)):
                            for z in range(max(cz - r - 1, gz << rb), min(cz + r + 1, (gz + 1) << rb)):
                                if x < 0 or z < 0 or x >= X or y >= Y or z >= Z or (x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2 >= r * r:
                                    continue
                                cid = (gz * self.ny + gy) * self.nx + gx
                                if self._own(cid):
                                    self.chunks[cid][(((z & (R - 1)) << rb | (y & (R - 1))) << rb) | (x & (R - 1))] = v
                                if v:
                                    idx = x + y + self.shw * z
                                    if not (self.shadow[idx] >= y + 1 or self.shadow[idx + 1] >= y + 1):
                                        self.shadow[idx] = y
        self.version += 1
        return np.array(dirty, np.uint32)

    # ---- slab borders (voxplat_b200.slab.SlabRebuilder) ----
    def halo_plane_bytes(self):
        return self.per_row * self.R * self.R

    def border_stream_handle(self):
        return 0

    def _plane(self, ptr):
        return np.ctypeslib.as_array((C.c_uint8 * self.halo_plane_bytes()).from_address(ptr)).reshape(self.per_row, self.R * self.R)

    def halo_pack(self, which, ptr):
        row, zs = (self.z0, 0) if which == 0 else (self.z1 - 1, self.R - 1)
        buf = self._plane(ptr)
        for k in range(self.per_row):
            c = self.chunks.get(row * self.per_row + k)
            buf[k] = 0 if c is None else c[zs * self.R * self.R:(zs + 1) * self.R * self.R]

    def halo_unpack(self, which, ptr):
        row, zs = (self.z1, 0) if which == 0 else (self.z0 - 1, self.R - 1)
        buf = self._plane(ptr)
        for k in range(self.per_row):
            ghost = np.zeros(self.N, np.uint8)
            ghost[zs * self.R * self.R:(zs + 1) * self.R * self.R] = buf[k]
            cid = row * self.per_row + k
            old = self.chunks.get(cid)
            if ghost.any():
                if old is None or not np.array_equal(old, ghost):
                    self.chunks[cid] = ghost
                    self.version += 1
            elif old is not None:
                del self.chunks[cid]
                self.version += 1

    # ---- rebuild ----
    def batch_prepare(self, ids, flags=vapi.VP_REBUILD_SPLAT, per_chunk_flags=None):
        ids = np.asarray(ids, np.uint32)
        f = np.full(len(ids), flags, np.uint8) if per_chunk_flags is None else np.asarray(per_chunk_flags, np.uint8)
        self.batch = (ids, f)

    def _rebuild(self, ids, flags):
        key = (self.version, ids.tobytes(), flags.tobytes())
        if key in self.cache:
            return self.cache[key]
        ptrs = (C.c_void_p * self.n_chunks)(*[self.chunks[i].ctypes.data if i in self.chunks else None for i in range(self.n_chunks)])
        lib = helpers.oracle_lib()
        w = helpers.VoWorld(self.rb, (C.c_int32 * 3)(*self.bits), C.cast(ptrs, C.c_void_p), self.shadow.ctypes.data)
        res = np.zeros(len(ids), vapi.RESULT_DTYPE)
        splat, mesh = [], []
        so = mo = 0
        geom = np.zeros((self.R + 1) ** 3 * 5, np.int16)
        for k, cid in enumerate(ids):
            if flags[k] & vapi.VP_REBUILD_SPLAT:
                items = (C.c_uint32 * 5)()
                n = lib.vo_chunk_splat(C.byref(w), C.c_uint32(int(cid)), helpers.vp(geom), items)
                res["svl_items"][k] = list(items)
                res["svl_items_total"][k] = n
                res["svl_offset"][k] = so
                splat.append(geom[:n].copy().view(np.uint8))
                so += n * 2
            if flags[k] & vapi.VP_REBUILD_MESH:
                f = lib.vo_chunk_mesh_faces(C.byref(w), C.c_uint32(int(cid)))
                vbo, ibo = np.zeros(max(f, 1) * 16, np.int16), np.zeros(max(f, 1) * 6, np.uint32)
                nv, ni = C.c_uint32(), C.c_uint32()
                lib.vo_chunk_mesh(C.byref(w), C.c_uint32(int(cid)), helpers.vp(vbo), helpers.vp(ibo), C.byref(nv), C.byref(ni))
                res["vbo_items"][k], res["ibo_items"][k] = nv.value, ni.value
                res["vbo_offset"][k] = mo
                mesh.append(vbo[:nv.value].view(np.uint8))
                mo += nv.value * 2
                res["ibo_offset"][k] = mo
                mesh.append(ibo[:ni.value].view(np.uint8))
                mo += ni.value * 4
        out = (res, np.concatenate(splat) if splat else np.zeros(0, np.uint8), np.concatenate(mesh) if mesh else np.zeros(0, np.uint8))
        self.cache = {key: out}
        return out

    def rebuild_device(self):
        self.last = self._rebuild(*self.batch)
        self.launches += 4

    def rebuild_device_part(self, part):
        if part == 1:
            self.rebuild_device()

    def rebuild_device_results(self):
        res, splat, mesh = self.last
        return res.copy(), splat.size, mesh.size

    def arena_download(self, which, nbytes):
        return self.last[1 + which][:nbytes].copy()

    def rebuild_batch(self, ids, flags=vapi.VP_REBUILD_SPLAT, per_chunk_flags=None):
        self.batch_prepare(ids, flags, per_chunk_flags)
        self.rebuild_device()
        return self.last[0].copy(), self.last[1].copy(), self.last[2].copy()

    def rebuild_from_rle(self, ids, words, word_offsets, flags=vapi.VP_REBUILD_SPLAT, per_chunk_flags=None, n_blocks=8):
        self.upload_chunks_rle(ids, np.asarray(words), word_offsets)
        return self.rebuild_batch(ids, flags, per_chunk_flags)

    def kernel_launches(self):
        return self.launches

    # ---- the rest of the Context surface the GPU tests use (tests/test_unverified_gpu_tests_dryrun.py) ----
    def download_chunks_dense(self, ids):
        out = np.zeros((len(ids), self.N), np.uint8)
        for k, cid in enumerate(ids):
            if int(cid) in self.chunks:
                out[k] = self.chunks[int(cid)]
        return out

    def rle_compress(self, data):
        return helpers.rle_encode(data)

    def rle_decompress(self, words, cap_bytes):
        return helpers.rle_decode(words, cap_bytes)

    def raycast(self, origins, vectors, normals=None):
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        v = np.ascontiguousarray(vectors, dtype=np.float32).reshape(-1, 3)
        n = len(o)
        ptrs = (C.c_void_p * self.n_chunks)(*[self.chunks[i].ctypes.data if i in self.chunks else None for i in range(self.n_chunks)])
        w = helpers.VoWorld(self.rb, (C.c_int32 * 3)(*self.bits), C.cast(ptrs, C.c_void_p), self.shadow.ctypes.data)
        vox, coords = np.zeros(n, np.uint8), np.zeros((n, 3), np.uint32)
        nrm = np.zeros((n, 3), np.int8) if normals is None else np.array(normals, np.int8).reshape(n, 3)
        lib = helpers.oracle_lib()
        for i in range(n):
            c, m = (C.c_uint32 * 3)(), (C.c_int8 * 3)(*[int(x) for x in nrm[i]])
            vox[i] = lib.vo_raycast(C.byref(w), (C.c_float * 3)(*o[i].tolist()), (C.c_float * 3)(*v[i].tolist()), c, m)
            coords[i], nrm[i] = list(c), list(m)
        return vox, coords, nrm

    def build_lod_nodes(self, lod, download=True):
        """The gather of gfx_update_svl over the splat lists of the last rebuild (which must have covered every chunk, in id order)."""
        res, splat, _ = self.last
        assert len(res) == self.n_chunks
        lib = helpers.oracle_lib()
        lib.vo_lod_node.restype = C.c_uint32
        lists = [np.ascontiguousarray(splat[int(res["svl_offset"][c]):int(res["svl_offset"][c]) + int(res["svl_items_total"][c]) * 2]) if res["svl_items_total"][c]
                 else np.zeros(8, np.uint8) for c in range(self.n_chunks)]
        ptrs = (C.c_void_p * self.n_chunks)(*[a.ctypes.data for a in lists])
        items = np.ascontiguousarray(res["svl_items"], np.uint32)
        cbits = (C.c_int32 * 3)(*self.bits)
        nn = 1 << sum(b - min(lod, b) for b in self.bits)
        nodes = np.zeros(nn, vapi.NODE_DTYPE)
        bufs, off = [], 0
        for node in range(nn):
            n = lib.vo_lod_node(cbits, lod, node, ptrs, helpers.vp(items), None)
            nodes["items"][node], nodes["offset"][node] = n, off
            if n:
                out = np.zeros(n, np.int16)
                lib.vo_lod_node(cbits, lod, node, ptrs, helpers.vp(items), helpers.vp(out))
                bufs.append(out.view(np.uint8))
                off += n * 2
        return nodes, (np.concatenate(bufs) if bufs else np.zeros(0, np.uint8)), 0.0

    def kernel_ms_history(self, n):
        return np.full(n, 0.4), np.full(n, 0.1)
