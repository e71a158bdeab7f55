"""ctypes binding of the C ABI in include/voxplat_b200.h.

This is plumbing only: every data transformation happens in the CUDA library
(voxplat_b200/libvoxplat_b200.so).  There is no CPU fallback -- if the library is missing or no CUDA
device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VOXPLAT_B200_LIB", os.path.join(HERE, "libvoxplat_b200.so"))      # override: kernel tuning variants

VP_REBUILD_SPLAT = 1
VP_REBUILD_MESH = 2
VP_OK = 0
VP_ERR_ARENA_FULL = -4

STATUS_NAMES = {0: "VP_OK", -1: "VP_ERR_ARG", -2: "VP_ERR_NO_DEVICE", -3: "VP_ERR_CUDA", -4: "VP_ERR_ARENA_FULL",
                -5: "VP_ERR_RLE", -6: "VP_ERR_NOT_RESIDENT"}


class VoxplatError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (STATUS_NAMES.get(code, code), msg))
        self.code = code


class VpConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("root_bitw", C.c_int32), ("max_bitw", C.c_int32 * 3),
                ("slab_z0", C.c_int32), ("slab_z1", C.c_int32),
                ("splat_arena_bytes", C.c_uint64), ("mesh_arena_bytes", C.c_uint64), ("rle_arena_bytes", C.c_uint64)]


# numpy mirror of vp_chunk_result (48 bytes)
RESULT_DTYPE = np.dtype([("svl_offset", "<u8"), ("svl_items", "<u4", (5,)), ("svl_items_total", "<u4"),
                         ("vbo_offset", "<u8"), ("ibo_offset", "<u8"), ("vbo_items", "<u4"), ("ibo_items", "<u4")])
assert RESULT_DTYPE.itemsize == 56
NODE_DTYPE = np.dtype([("offset", "<u8"), ("items", "<u4"), ("members", "<u4")])

_lib = None


def world_file_info(path):
    """(root_bitw, max_bitw, file_bytes) of a world file; needs no device."""
    lib = load_library()
    rb, mb, nb = C.c_int32(), (C.c_int32 * 3)(), C.c_uint64()
    rc = lib.vp_world_file_info(os.fsencode(path), C.byref(rb), C.byref(mb), C.byref(nb))
    if rc != VP_OK:
        raise RuntimeError("voxplat_b200: %s is not a VOXPLAT world file (error %d)" % (path, rc))
    return int(rb.value), tuple(int(x) for x in mb), int(nb.value)


def load_library():
    """Load the CUDA C-ABI library; raises if it has not been built (python -m voxplat_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("voxplat_b200: %s not found -- run `python -m voxplat_b200.build` (nvcc, sm_100a). "
                           "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sig = {
        "vp_ctx_create": (C.c_int, [C.POINTER(VpConfig), C.POINTER(vp)]),
        "vp_ctx_destroy": (None, [vp]),
        "vp_last_error": (C.c_char_p, [vp]),
        "vp_version": (C.c_char_p, []),
        "vp_ctx_resize_arenas": (C.c_int, [vp, C.c_uint64, C.c_uint64]),
        "vp_ctx_set_stream": (C.c_int, [vp, vp]),
        "vp_ctx_synchronize": (C.c_int, [vp]),
        "vp_kernel_launches": (C.c_uint64, [vp, C.c_int]),
        "vp_upload_chunks_dense": (C.c_int, [vp, vp, C.c_uint32, vp]),
        "vp_upload_chunks_rle": (C.c_int, [vp, vp, C.c_uint32, vp, vp]),
        "vp_set_chunks_null": (C.c_int, [vp, vp, C.c_uint32]),
        "vp_chunks_resident": (C.c_int, [vp, vp, C.c_uint32, vp]),
        "vp_download_chunks_dense": (C.c_int, [vp, vp, C.c_uint32, vp]),
        "vp_encode_chunks_rle": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint64, vp]),
        "vp_upload_shadow_rows": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp]),
        "vp_upload_shadow_rows_async": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp]),
        "vp_rle_compress": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32, C.POINTER(C.c_uint32)]),
        "vp_rle_decompress": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32, C.POINTER(C.c_uint32)]),
        "vp_rebuild_batch": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, vp, vp, C.POINTER(vp), C.POINTER(vp)]),
        "vp_rebuild_from_rle": (C.c_int, [vp, vp, C.c_uint32, vp, vp, C.c_uint32, vp, C.c_uint32, vp, C.POINTER(vp), C.POINTER(vp)]),
        "vp_batch_prepare": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32]),
        "vp_rebuild_device": (C.c_int, [vp]),
        "vp_rebuild_device_part": (C.c_int, [vp, C.c_int]),
        "vp_rebuild_device_results": (C.c_int, [vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "vp_kernel_ms_history": (C.c_int, [vp, C.c_uint32, vp, vp]),
        "vp_splat_arena_device": (vp, [vp]),
        "vp_mesh_arena_device": (vp, [vp]),
        "vp_arena_download": (C.c_int, [vp, C.c_int, vp, C.c_uint64]),
        "vp_chunk_make_splatlists": (C.c_int64, [vp, C.c_uint32, vp, C.c_uint64, vp]),
        "vp_chunk_make_mesh": (C.c_int, [vp, C.c_uint32, vp, C.c_uint64, C.POINTER(C.c_uint32), vp, C.c_uint64, C.POINTER(C.c_uint32)]),
        "vp_edit_sphere": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint8, vp, C.c_uint32, C.POINTER(C.c_uint32)]),
        "vp_download_shadow_rows": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp]),
        "vp_build_lod_nodes": (C.c_int, [vp, C.c_uint32, vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(C.c_float)]),
        "vp_generate_world": (C.c_int, [vp, C.c_uint32]),
        "vp_world_save": (C.c_int, [vp, C.c_char_p, C.POINTER(C.c_uint64)]),
        "vp_world_load": (C.c_int, [vp, C.c_char_p]),
        "vp_world_file_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32 * 3), C.POINTER(C.c_uint64)]),
        "vp_raycast": (C.c_int, [vp, C.c_uint32, vp, vp, vp, vp, vp]),
        "vp_halo_plane_bytes": (C.c_uint64, [vp]),
        "vp_ctx_border_stream": (vp, [vp]),
        "vp_device_count": (C.c_int32, []),
        "vp_multi_create": (C.c_int, [C.POINTER(VpConfig), vp, C.c_int32, C.POINTER(vp)]),
        "vp_multi_destroy": (None, [vp]),
        "vp_multi_last_error": (C.c_char_p, [vp]),
        "vp_multi_devices": (C.c_int32, [vp]),
        "vp_multi_ctx": (vp, [vp, C.c_int32]),
        "vp_multi_owner": (C.c_int32, [vp, C.c_uint32]),
        "vp_multi_upload_chunks_dense": (C.c_int, [vp, vp, C.c_uint32, vp]),
        "vp_multi_upload_chunks_rle": (C.c_int, [vp, vp, C.c_uint32, vp, vp]),
        "vp_multi_set_chunks_null": (C.c_int, [vp, vp, C.c_uint32]),
        "vp_multi_upload_shadow_rows": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp]),
        "vp_multi_exchange_halos": (C.c_int, [vp, C.c_int32]),
        "vp_multi_batch_prepare": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32]),
        "vp_multi_rebuild_device": (C.c_int, [vp, C.c_int32]),
        "vp_multi_synchronize": (C.c_int, [vp]),
        "vp_multi_rebuild_batch": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp]),
        "vp_halo_pack": (C.c_int, [vp, C.c_int, vp]),
        "vp_halo_unpack": (C.c_int, [vp, C.c_int, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._vp_signatures = sig
    _lib = lib
    return lib


def _ptr(a):
    """Pointer of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a


class Context:
    """One device-resident world (or z-slab of it) + its rebuild machinery (struct vp_ctx)."""

    def __init__(self, root_bitw, max_bitw, device=0, slab=None, splat_arena_bytes=0, mesh_arena_bytes=0,
                 rle_arena_bytes=0):
        self.lib = load_library()
        cfg = VpConfig()
        cfg.device = device
        cfg.root_bitw = root_bitw
        cfg.max_bitw = (C.c_int32 * 3)(*max_bitw)
        cfg.slab_z0, cfg.slab_z1 = (slab if slab else (0, 0))
        cfg.splat_arena_bytes = splat_arena_bytes
        cfg.mesh_arena_bytes = mesh_arena_bytes
        cfg.rle_arena_bytes = rle_arena_bytes
        h = C.c_void_p()
        rc = self.lib.vp_ctx_create(C.byref(cfg), C.byref(h))
        if rc:
            raise VoxplatError(rc, self.lib.vp_last_error(None).decode())
        self.h = h
        self.root_bitw = root_bitw
        self.R = 1 << root_bitw
        self.N = self.R ** 3
        self.max_bitw = tuple(max_bitw)
        self.n_chunks = 1 << sum(max_bitw)
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.vp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise VoxplatError(rc, self.lib.vp_last_error(self.h).decode())

    # ---- residency ---------------------------------------------------------------------------------
    def upload_chunks_dense(self, ids, dense):
        ids = _u32(ids)
        self._ck(self.lib.vp_upload_chunks_dense(self.h, _ptr(ids), len(ids), _ptr(dense)))

    def upload_chunks_rle(self, ids, words, word_offsets):
        ids = _u32(ids)
        word_offsets = np.ascontiguousarray(word_offsets, dtype=np.uint64)
        assert len(word_offsets) == len(ids) + 1
        self._ck(self.lib.vp_upload_chunks_rle(self.h, _ptr(ids), len(ids), _ptr(words), _ptr(word_offsets)))

    def set_chunks_null(self, ids):
        ids = _u32(ids)
        self._ck(self.lib.vp_set_chunks_null(self.h, _ptr(ids), len(ids)))

    def download_chunks_dense(self, ids):
        ids = _u32(ids)
        out = np.empty((len(ids), self.N), np.uint8)
        self._ck(self.lib.vp_download_chunks_dense(self.h, _ptr(ids), len(ids), _ptr(out)))
        return out

    def encode_chunks_rle(self, ids, cap_words=None):
        ids = _u32(ids)
        offs = np.zeros(len(ids) + 1, np.uint64)
        cap = cap_words if cap_words is not None else len(ids) * (self.N // 8 + 2)
        while True:
            words = np.empty(cap, np.uint32)
            rc = self.lib.vp_encode_chunks_rle(self.h, _ptr(ids), len(ids), _ptr(words), cap, _ptr(offs))
            if rc == VP_ERR_ARENA_FULL and int(offs[-1]) > cap:
                cap = int(offs[-1])
                continue
            self._ck(rc)
            return words[:int(offs[-1])], offs

    def chunks_resident(self, ids):
        """Boolean array: True where the chunk holds voxels on the device (False = null chunk)."""
        ids = _u32(ids)
        out = np.zeros(len(ids), np.uint8)
        self._ck(self.lib.vp_chunks_resident(self.h, _ptr(ids), len(ids), _ptr(out)))
        return out.astype(bool)

    def generate_world(self, seed):
        """Generate the slab's chunks and height-map rows on the device (same world as worldgen.World(seed, ...))."""
        self._ck(self.lib.vp_generate_world(self.h, seed))

    # ---- world file: checkpoint / resume (deadcode.c:320-350 layout) -------------------------------------
    def save_world(self, path):
        """Write the resident world (device RLE encode of every chunk + shadow map); returns the file size."""
        n = C.c_uint64()
        self._ck(self.lib.vp_world_save(self.h, os.fsencode(path), C.byref(n)))
        return int(n.value)

    def load_world(self, path):
        """Replace the resident world by the file's (device RLE decode); geometry must match the context's."""
        self._ck(self.lib.vp_world_load(self.h, os.fsencode(path)))

    def upload_shadow_rows(self, z0, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint16)
        shw = ((1 << self.max_bitw[0]) + (1 << self.max_bitw[1])) << self.root_bitw
        assert rows.size % shw == 0
        self._ck(self.lib.vp_upload_shadow_rows(self.h, z0, z0 + rows.size // shw, _ptr(rows)))

    def upload_shadow_rows_async(self, z0, pinned_rows):
        """No wait for the copy; `pinned_rows` (a pinned torch tensor of uint16 / int16) must outlive the next synchronous call."""
        shw = ((1 << self.max_bitw[0]) + (1 << self.max_bitw[1])) << self.root_bitw
        n = pinned_rows.numel()
        assert n % shw == 0 and pinned_rows.is_pinned()
        self._ck(self.lib.vp_upload_shadow_rows_async(self.h, z0, z0 + n // shw, _ptr(pinned_rows)))

    # ---- flat RLE codec (rle.h:7-8) -----------------------------------------------------------------
    def rle_compress(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(data.size + 1, np.uint32)
        n = C.c_uint32()
        self._ck(self.lib.vp_rle_compress(self.h, _ptr(data), data.size, _ptr(out), out.size, C.byref(n)))
        return out[:n.value].copy()

    def rle_decompress(self, words, cap_bytes):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        out = np.empty(cap_bytes, np.uint8)
        n = C.c_uint32()
        self._ck(self.lib.vp_rle_decompress(self.h, _ptr(words), words.size, _ptr(out), out.size, C.byref(n)))
        return out[:n.value]

    # ---- rebuild ----------------------------------------------------------------------------------
    def rebuild_batch(self, ids, flags=VP_REBUILD_SPLAT, per_chunk_flags=None):
        """Synchronous host-facing rebuild.  Returns (results, splat_bytes, mesh_bytes): results is a
        structured array (RESULT_DTYPE); the byte arrays are views of the context's pinned staging, valid
        until the next rebuild."""
        ids = _u32(ids)
        res = np.zeros(len(ids), RESULT_DTYPE)
        pcf = np.ascontiguousarray(per_chunk_flags, dtype=np.uint8) if per_chunk_flags is not None else None
        sb, mb = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.vp_rebuild_batch(self.h, _ptr(ids), len(ids), flags, _ptr(pcf), _ptr(res), C.byref(sb), C.byref(mb)))
        s_end = int((res["svl_offset"] + res["svl_items_total"].astype(np.uint64) * 2).max()) if len(ids) else 0
        m_end = int(np.maximum(res["vbo_offset"] + res["vbo_items"].astype(np.uint64) * 2,
                               res["ibo_offset"] + res["ibo_items"].astype(np.uint64) * 4).max()) if len(ids) else 0
        splat = np.ctypeslib.as_array(C.cast(sb, C.POINTER(C.c_uint8)), shape=(s_end,)) if s_end and sb.value else np.zeros(0, np.uint8)
        mesh = np.ctypeslib.as_array(C.cast(mb, C.POINTER(C.c_uint8)), shape=(m_end,)) if m_end and mb.value else np.zeros(0, np.uint8)
        return res, splat, mesh

    def rebuild_from_rle(self, ids, words, word_offsets, flags=VP_REBUILD_SPLAT, per_chunk_flags=None, n_blocks=8):
        """Host RLE in -> host buffers out in one pipelined call (see vp_rebuild_from_rle)."""
        ids = _u32(ids)
        word_offsets = np.ascontiguousarray(word_offsets, dtype=np.uint64)
        res = np.zeros(len(ids), RESULT_DTYPE)
        pcf = np.ascontiguousarray(per_chunk_flags, dtype=np.uint8) if per_chunk_flags is not None else None
        sb, mb = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.vp_rebuild_from_rle(self.h, _ptr(ids), len(ids), _ptr(words), _ptr(word_offsets), flags, _ptr(pcf), n_blocks,
                                              _ptr(res), C.byref(sb), C.byref(mb)))
        s_end = int((res["svl_offset"] + res["svl_items_total"].astype(np.uint64) * 2).max()) if len(ids) else 0
        m_end = int(np.maximum(res["vbo_offset"] + res["vbo_items"].astype(np.uint64) * 2,
                               res["ibo_offset"] + res["ibo_items"].astype(np.uint64) * 4).max()) if len(ids) else 0
        splat = np.ctypeslib.as_array(C.cast(sb, C.POINTER(C.c_uint8)), shape=(s_end,)) if s_end and sb.value else np.zeros(0, np.uint8)
        mesh = np.ctypeslib.as_array(C.cast(mb, C.POINTER(C.c_uint8)), shape=(m_end,)) if m_end and mb.value else np.zeros(0, np.uint8)
        return res, splat, mesh

    def batch_prepare(self, ids, flags=VP_REBUILD_SPLAT, per_chunk_flags=None):
        ids = _u32(ids)
        pcf = np.ascontiguousarray(per_chunk_flags, dtype=np.uint8) if per_chunk_flags is not None else None
        self._batch_n = len(ids)
        self._ck(self.lib.vp_batch_prepare(self.h, _ptr(ids), len(ids), _ptr(pcf), flags))

    def rebuild_device(self):
        self._ck(self.lib.vp_rebuild_device(self.h))

    def rebuild_device_part(self, part):
        """part 0: chunks that do not read a ghost row; part 1 (after halo_unpack): the slab's border chunks."""
        self._ck(self.lib.vp_rebuild_device_part(self.h, part))

    def rebuild_device_results(self):
        res = np.zeros(self._batch_n, RESULT_DTYPE)
        sb, mb = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.vp_rebuild_device_results(self.h, _ptr(res), C.byref(sb), C.byref(mb)))
        return res, sb.value, mb.value

    def kernel_ms_history(self, n):
        """(splat_ms[n], mesh_ms[n]) of the last n rebuild_device calls."""
        a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self._ck(self.lib.vp_kernel_ms_history(self.h, n, _ptr(a), _ptr(b)))
        return a, b

    def arena_download(self, which, nbytes):
        out = np.empty(nbytes, np.uint8)
        if nbytes:
            self._ck(self.lib.vp_arena_download(self.h, which, _ptr(out), nbytes))
        return out

    def resize_arenas(self, splat_bytes=0, mesh_bytes=0):
        self._ck(self.lib.vp_ctx_resize_arenas(self.h, splat_bytes, mesh_bytes))

    def set_stream(self, cuda_stream_handle):
        self._ck(self.lib.vp_ctx_set_stream(self.h, C.c_void_p(cuda_stream_handle) if cuda_stream_handle else None))

    def synchronize(self):
        self._ck(self.lib.vp_ctx_synchronize(self.h))

    def kernel_launches(self, reset=False):
        return int(self.lib.vp_kernel_launches(self.h, int(reset)))

    # ---- single-chunk wrappers (mesher.h) -------------------------------------------------------------
    def chunk_make_splatlists(self, chunk_id, cap_items=None):
        cap = cap_items or (self.R + 1) ** 3 * 5
        geom = np.empty(cap, np.int16)
        items = np.zeros(5, np.uint32)
        n = self.lib.vp_chunk_make_splatlists(self.h, chunk_id, _ptr(geom), cap, _ptr(items))
        if n < 0:
            self._ck(int(n))
        return geom[:n].copy(), items

    def chunk_make_mesh(self, chunk_id, cap_faces=None):
        faces = cap_faces or 3 * self.N
        vbo = np.empty(faces * 16, np.int16)
        ibo = np.empty(faces * 6, np.uint32)
        nv, ni = C.c_uint32(), C.c_uint32()
        self._ck(self.lib.vp_chunk_make_mesh(self.h, chunk_id, _ptr(vbo), vbo.size, C.byref(nv), _ptr(ibo), ibo.size, C.byref(ni)))
        return vbo[:nv.value].copy(), ibo[:ni.value].copy()

    # ---- device-side edits (chunkset_edit_sphere) ----------------------------------------------------------
    def edit_sphere(self, x, y, z, radius, voxel):
        """Returns the dirty chunk ids (the reference's list, in its order)."""
        ids = np.zeros(1024, np.uint32)
        n = C.c_uint32()
        self._ck(self.lib.vp_edit_sphere(self.h, x, y, z, radius, voxel, _ptr(ids), ids.size, C.byref(n)))
        return ids[:n.value].copy()

    def raycast(self, origins, vectors, normals=None):
        """chunkset_edit_raycast_until_solid for n rays: returns (voxels[n], coords[n,3], normals[n,3])."""
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        v = np.ascontiguousarray(vectors, dtype=np.float32).reshape(-1, 3)
        n = len(o)
        coords = np.zeros((n, 3), np.uint32)
        nrm = np.zeros((n, 3), np.int8) if normals is None else np.ascontiguousarray(normals, dtype=np.int8).reshape(n, 3).copy()
        vox = np.zeros(n, np.uint8)
        self._ck(self.lib.vp_raycast(self.h, n, _ptr(o), _ptr(v), _ptr(coords), _ptr(nrm), _ptr(vox)))
        return vox, coords, nrm

    def download_shadow_rows(self, z0, z1):
        shw = ((1 << self.max_bitw[0]) + (1 << self.max_bitw[1])) << self.root_bitw
        rows = np.zeros((z1 - z0) * shw, np.uint16)
        self._ck(self.lib.vp_download_shadow_rows(self.h, z0, z1, _ptr(rows)))
        return rows

    # ---- LOD-node aggregation (gfx_update_svl's gather) ------------------------------------------------
    def build_lod_nodes(self, lod, download=True):
        """Returns (nodes structured array, node buffer bytes or None, kernel ms)."""
        nn = 1
        for b in self.max_bitw:
            nn <<= b - min(lod, b)
        nodes = np.zeros(nn, NODE_DTYPE)
        n = C.c_uint32()
        base = C.c_void_p()
        ms = C.c_float()
        self._ck(self.lib.vp_build_lod_nodes(self.h, lod, _ptr(nodes), nn, C.byref(n), C.byref(base) if download else None, C.byref(ms)))
        buf = None
        if download:
            end = int((nodes["offset"] + nodes["items"].astype(np.uint64) * 2)[nodes["items"] > 0].max()) if (nodes["items"] > 0).any() else 0
            buf = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), shape=(end,)) if end else np.zeros(0, np.uint8)
        return nodes, buf, float(ms.value)

    # ---- multi-GPU borders ----------------------------------------------------------------------------
    def halo_plane_bytes(self):
        return int(self.lib.vp_halo_plane_bytes(self.h))

    def border_stream_handle(self):
        """cudaStream_t (as int) on which halo_pack / halo_unpack / rebuild_device_part(1) run."""
        return int(self.lib.vp_ctx_border_stream(self.h) or 0)

    def halo_pack(self, which, device_ptr):
        self._ck(self.lib.vp_halo_pack(self.h, which, C.c_void_p(device_ptr)))

    def halo_unpack(self, which, device_ptr):
        self._ck(self.lib.vp_halo_unpack(self.h, which, C.c_void_p(device_ptr)))


class MultiContext:
    """Several GPUs behind one handle, one host thread (struct vp_multi): z-slabs of chunk rows, border planes pushed
    peer to peer.  `devices` may repeat a device (several slabs on one GPU: used by the tests on 1-GPU boxes)."""

    def __init__(self, root_bitw, max_bitw, devices, splat_arena_bytes=0, mesh_arena_bytes=0, rle_arena_bytes=0):
        self.lib = load_library()
        cfg = VpConfig()
        cfg.root_bitw = root_bitw
        cfg.max_bitw = (C.c_int32 * 3)(*max_bitw)
        cfg.splat_arena_bytes, cfg.mesh_arena_bytes, cfg.rle_arena_bytes = splat_arena_bytes, mesh_arena_bytes, rle_arena_bytes
        devs = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.vp_multi_create(C.byref(cfg), devs, len(devices), C.byref(h))
        if rc:
            raise VoxplatError(rc, self.lib.vp_multi_last_error(None).decode())
        self.h, self.n_dev = h, len(devices)
        self.root_bitw, self.max_bitw = root_bitw, tuple(max_bitw)

    def close(self):
        if getattr(self, "h", None):
            self.lib.vp_multi_destroy(self.h)
            self.h = None

    def _ck(self, rc):
        if rc:
            raise VoxplatError(rc, self.lib.vp_multi_last_error(self.h).decode())

    def upload_chunks_dense(self, ids, dense):
        ids = _u32(ids)
        self._ck(self.lib.vp_multi_upload_chunks_dense(self.h, _ptr(ids), len(ids), _ptr(dense)))

    def upload_chunks_rle(self, ids, words, word_offsets):
        ids = _u32(ids)
        words = np.ascontiguousarray(words, dtype=np.uint32)
        offs = np.ascontiguousarray(word_offsets, dtype=np.uint64)
        self._ck(self.lib.vp_multi_upload_chunks_rle(self.h, _ptr(ids), len(ids), _ptr(words), _ptr(offs)))

    def set_chunks_null(self, ids):
        ids = _u32(ids)
        self._ck(self.lib.vp_multi_set_chunks_null(self.h, _ptr(ids), len(ids)))

    def upload_shadow_rows(self, z0, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint16)
        shw = ((1 << self.max_bitw[0]) + (1 << self.max_bitw[1])) << self.root_bitw
        self._ck(self.lib.vp_multi_upload_shadow_rows(self.h, z0, z0 + rows.size // shw, _ptr(rows)))

    def rebuild_batch(self, ids, flags=VP_REBUILD_SPLAT, per_chunk_flags=None):
        """Returns (results, owner, [splat bytes per device], [mesh bytes per device])."""
        ids = _u32(ids)
        res = np.zeros(len(ids), RESULT_DTYPE)
        owner = np.zeros(len(ids), np.uint8)
        pcf = np.ascontiguousarray(per_chunk_flags, dtype=np.uint8) if per_chunk_flags is not None else None
        sb, mb = (C.c_void_p * self.n_dev)(), (C.c_void_p * self.n_dev)()
        self._ck(self.lib.vp_multi_rebuild_batch(self.h, _ptr(ids), len(ids), flags, _ptr(pcf), _ptr(res), _ptr(owner), sb, mb))
        splat, mesh = [], []
        for d in range(self.n_dev):
            m = owner == d
            s_end = int((res["svl_offset"][m] + res["svl_items_total"][m].astype(np.uint64) * 2).max()) if m.any() else 0
            m_end = int(np.maximum(res["vbo_offset"][m] + res["vbo_items"][m].astype(np.uint64) * 2,
                                   res["ibo_offset"][m] + res["ibo_items"][m].astype(np.uint64) * 4).max()) if m.any() else 0
            splat.append(np.ctypeslib.as_array(C.cast(sb[d], C.POINTER(C.c_uint8)), shape=(s_end,)) if s_end and sb[d] else np.zeros(0, np.uint8))
            mesh.append(np.ctypeslib.as_array(C.cast(mb[d], C.POINTER(C.c_uint8)), shape=(m_end,)) if m_end and mb[d] else np.zeros(0, np.uint8))
        return res, owner, splat, mesh
