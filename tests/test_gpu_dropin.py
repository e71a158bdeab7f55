"""Drop-in test: the reference ENGINE (its ChunkSet state machine, allocator, edit code, compiled unmodified)
runs once with its own chunkset_manage (CPU, OpenMP) and once with the CUDA drop-in
voxplat_b200/host/vp_chunkset_manage.c behind the same entry point; everything published into struct ChunkMD
(svl + svl_items[5], mesh_vbo / mesh_ibo + counts) must be byte-identical, also after brush edits."""
import ctypes as C
import os
import time

import numpy as np
import pytest

import helpers
from voxplat_b200 import worldgen

pytestmark = pytest.mark.gpu
GPU_SO = os.path.join(helpers.ROOT, "oracle", "_ref", "libvoxref_gpu.so")


def load():
    if not os.path.exists(GPU_SO):
        pytest.skip("oracle/_ref/libvoxref_gpu.so not built (needs /root/reference at build time)")
    lib = C.CDLL(GPU_SO)
    lib.vr_world_create.restype = C.c_void_p
    lib.vr_init(C.c_uint64(4 << 30))
    lib.vr_set_scratch_scale(12)
    return lib


def make_set(lib, w, mesh_ids):
    s = C.c_void_p(lib.vr_world_create(w.root_bitw, *w.max_bitw))
    for i in range(w.n_chunks):
        if w.solid[i]:
            lib.vr_world_set_chunk(s, C.c_uint32(i), C.c_void_p(w.dense[i].ctypes.data))
    lib.vr_world_set_shadow(s, helpers.vp(w.shadow), C.c_uint32(w.shadow.size))
    for i in mesh_ids:
        lib.vr_chunk_set_make_mesh(s, C.c_uint32(i), 1)
    return s


def drain(lib, s, n_chunks, manage):
    """Run the dispatcher until nothing is pending; collect what it published (acknowledging like gfx_update_*)."""
    pub = {}
    idle = 0
    for _ in range(400):
        manage(s)
        new = 0
        for i in range(n_chunks):
            svl, vbo, ibo = C.c_void_p(), C.c_void_p(), C.c_void_p()
            items = (C.c_uint32 * 5)()
            tot, nv, ni = C.c_uint32(), C.c_uint32(), C.c_uint32()
            fl = lib.vr_chunk_published(s, C.c_uint32(i), C.byref(svl), items, C.byref(tot), C.byref(vbo), C.byref(nv),
                                        C.byref(ibo), C.byref(ni), 1)
            if fl & 1:
                pub[(i, "svl")] = (list(items), C.string_at(svl, tot.value * 2) if svl.value and tot.value else b"")
                new += 1
            if fl & 2:
                pub[(i, "mesh")] = (nv.value, ni.value, C.string_at(vbo, nv.value * 2) if nv.value else b"",
                                    C.string_at(ibo, ni.value * 4) if ni.value else b"")
                new += 1
        pending = sum(lib.vr_chunk_pending(s, C.c_uint32(i)) for i in range(n_chunks))
        idle = idle + 1 if (new == 0 and pending == 0) else 0
        if idle >= 2:
            return pub
        time.sleep(0.03)
    raise AssertionError("dispatcher did not drain")


def slab_device_lists():
    """"": the dispatcher's own choice (every visible GPU); "0,0": two z-slabs on one device (vp_multi, 1-GPU boxes);
    "0,1": one slab per device, border planes over NVLink peer access (needs 2 GPUs)."""
    import torch
    out = ["", "0,0"]
    if torch.cuda.is_available() and torch.cuda.device_count() >= 2:
        out.append("0,1")
    return out


@pytest.mark.parametrize("devices", ["", "0,0", "0,1"])
def test_chunkset_manage_dropin_matches_reference_dispatcher(devices):
    if devices not in slab_device_lists():
        pytest.skip("needs 2 GPUs")
    if devices:
        os.environ["VP_DEVICES"] = devices                     # read by the drop-in when it attaches to a new ChunkSet
    else:
        os.environ.pop("VP_DEVICES", None)
    lib = load()
    w = worldgen.World(2024, 5, (2, 1, 2))
    mesh_ids = [0, 1, 4, 5]
    a, b = make_set(lib, w, mesh_ids), make_set(lib, w, mesh_ids)
    pa = drain(lib, a, w.n_chunks, lib.vr_manage_cpu)          # the reference's own loop
    pb = drain(lib, b, w.n_chunks, lib.vr_manage)              # same entry point, CUDA behind it
    assert pa.keys() == pb.keys() and len(pa) >= w.n_chunks
    for k in pa:
        assert pa[k] == pb[k], k
    # brush edits (chunkset_edit_sphere, edit.c:179-244): place and remove, crossing chunk borders
    for (x, y, z, r, v) in [(31, 20, 33, 4, 63), (64, 12, 64, 5, 0), (40, 30, 40, 3, 17)]:
        lib.vr_edit_sphere(a, x, y, z, r, v)
        lib.vr_edit_sphere(b, x, y, z, r, v)
    time.sleep(0.12)                                           # the 100 ms per-chunk throttle (chunkset.c:309)
    pa = drain(lib, a, w.n_chunks, lib.vr_manage_cpu)
    pb = drain(lib, b, w.n_chunks, lib.vr_manage)
    assert pa.keys() == pb.keys() and len(pa) >= 4
    for k in pa:
        assert pa[k] == pb[k], k
