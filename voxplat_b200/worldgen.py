"""Binding of the deterministic integer world generator (csrc/vp_worldgen.c): the fixed synthetic input
that the CUDA path, the oracle and the reference all consume."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvpworldgen.so")
_lib = None


class Params(C.Structure):
    _fields_ = [("seed", C.c_uint32), ("root_bitw", C.c_int32), ("bits", C.c_int32 * 3)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s not found -- run `python -m voxplat_b200.build`" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.vpw_height.restype = C.c_int32
        _lib.vpw_gen_chunk.restype = C.c_uint32
    return _lib


def params(seed, root_bitw, max_bitw):
    return Params(seed, root_bitw, (C.c_int32 * 3)(*max_bitw))


def gen_chunks(seed, root_bitw, max_bitw, ids, out=None):
    """Dense voxels of the listed chunks: returns (dense[n, R^3] uint8, solid[n] uint32)."""
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    N = 1 << (3 * root_bitw)
    if out is None:
        out = np.empty((len(ids), N), np.uint8)
    solid = np.zeros(len(ids), np.uint32)
    p = params(seed, root_bitw, max_bitw)
    lib().vpw_gen_chunks(C.byref(p), C.c_void_p(ids.ctypes.data), C.c_uint32(len(ids)),
                         C.c_void_p(out.ctypes.data if isinstance(out, np.ndarray) else out.data_ptr()),
                         C.c_void_p(solid.ctypes.data))
    return out, solid


def shadow_pad(root_bitw, max_bitw):
    """Zero entries that must follow the last shadow row (SURVEY 8a' u3: LOD-l splats sample at +(1<<l), up to 16 rows past the end)."""
    shw = ((1 << max_bitw[0]) + (1 << max_bitw[1])) << root_bitw
    return 17 * shw + 64


def shadow_rows(seed, root_bitw, max_bitw, chunk_ptrs, z0, z1):
    """Shadow-map rows [z0,z1) built from dense chunks by the reference's placement rule in a fixed order.
    chunk_ptrs: sequence of length n_chunks with the address of each chunk's dense bytes (0 = air)."""
    shw = ((1 << max_bitw[0]) + (1 << max_bitw[1])) << root_bitw
    rows = np.zeros((z1 - z0) * shw + 64, np.uint16)
    arr = (C.c_void_p * len(chunk_ptrs))(*[int(p) if p else None for p in chunk_ptrs])
    p = params(seed, root_bitw, max_bitw)
    lib().vpw_shadow_rows(C.byref(p), arr, C.c_uint32(z0), C.c_uint32(z1), C.c_void_p(rows.ctypes.data))
    return rows[:(z1 - z0) * shw]


class World:
    """A whole synthetic world held on the host: dense chunks (None for all-air) + padded shadow map."""

    def __init__(self, seed, root_bitw, max_bitw, dense=None):
        self.seed, self.root_bitw, self.max_bitw = seed, root_bitw, tuple(max_bitw)
        self.R = 1 << root_bitw
        self.N = self.R ** 3
        self.n_chunks = 1 << sum(max_bitw)
        self.dims = tuple((1 << b) * self.R for b in max_bitw)
        ids = np.arange(self.n_chunks, dtype=np.uint32)
        if dense is None:
            self.dense, self.solid = gen_chunks(seed, root_bitw, max_bitw, ids)
        else:
            self.dense = np.ascontiguousarray(dense, dtype=np.uint8).reshape(self.n_chunks, self.N)
            self.solid = (self.dense != 0).sum(axis=1).astype(np.uint32)
        self.shw = self.dims[0] + self.dims[1]
        base = self.dense.ctypes.data
        ptrs = [base + i * self.N if self.solid[i] else 0 for i in range(self.n_chunks)]
        rows = shadow_rows(seed, root_bitw, max_bitw, ptrs, 0, self.dims[2])
        self.shadow = np.zeros(rows.size + shadow_pad(root_bitw, max_bitw), np.uint16)
        self.shadow[:rows.size] = rows

    def nonnull_ids(self):
        return np.nonzero(self.solid)[0].astype(np.uint32)

    def chunk_ptrs(self):
        base = self.dense.ctypes.data
        return [base + i * self.N if self.solid[i] else 0 for i in range(self.n_chunks)]
