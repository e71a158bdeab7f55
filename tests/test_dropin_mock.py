"""The C drop-in dispatcher (voxplat_b200/host/vp_chunkset_manage.c) on CPU: the reference ENGINE compiled unmodified, its
dispatcher swapped for the drop-in, and a host stand-in for vp_multi_* behind it (oracle/mock/vp_multi_mock.c: its own
copy of the world, rebuilt with the oracle restatement -- like the device, it only knows what was uploaded to it).

1. the same comparison as tests/test_gpu_dropin.py -- reference dispatcher vs drop-in, everything published into ChunkMD
   byte-identical, also after brush edits -- so the dispatcher's bookkeeping is checked on every CPU run;
2. what the single-threaded GPU test cannot show: one thread keeps editing the world (chunkset_edit_sphere, the game's edit
   path) while another runs the dispatcher.  An edit that lands between the dispatcher's residency pass and its selection
   pass must not be lost (round 1's version could lose it for good: ADVICE.md); after the edits stop and the dispatcher
   drains, every chunk's published splat list must be the list of the FINAL world."""
import ctypes as C
import os
import threading
import time

import numpy as np
import pytest

import helpers
from voxplat_b200 import worldgen

MOCK_SO = os.path.join(helpers.ROOT, "oracle", "_ref", "libvoxref_mock.so")
pytestmark = pytest.mark.skipif(not os.path.exists(MOCK_SO), reason="oracle/_ref/libvoxref_mock.so not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def lib():
    lib = C.CDLL(MOCK_SO)
    lib.vr_world_create.restype = C.c_void_p
    lib.vr_init(C.c_uint64(2 << 30))
    lib.vr_set_scratch_scale(12)
    return lib


def make_set(lib, w, mesh_ids=()):
    s = C.c_void_p(lib.vr_world_create(w.root_bitw, *w.max_bitw))
    for i in range(w.n_chunks):
        if w.solid[i]:
            lib.vr_world_set_chunk(s, C.c_uint32(i), C.c_void_p(w.dense[i].ctypes.data))
    lib.vr_world_set_shadow(s, helpers.vp(w.shadow), C.c_uint32(w.shadow.size))
    for i in mesh_ids:
        lib.vr_chunk_set_make_mesh(s, C.c_uint32(i), 1)
    return s


def published(lib, s, i, ack, read=True):
    """What chunkset_manage published for chunk i (optionally acknowledged like gfx_update_svl / gfx_update_mesh do).
    read=False only acknowledges: the buffers may be replaced by the dispatcher at any time, and the real consumer reads
    them under ChunkMD.mutex_svl (gfx/vsplat.c:271) -- a test thread that runs BESIDE the dispatcher must not touch them."""
    svl, vbo, ibo = C.c_void_p(), C.c_void_p(), C.c_void_p()
    items = (C.c_uint32 * 5)()
    tot, nv, ni = C.c_uint32(), C.c_uint32(), C.c_uint32()
    fl = lib.vr_chunk_published(s, C.c_uint32(i), C.byref(svl), items, C.byref(tot), C.byref(vbo), C.byref(nv), C.byref(ibo), C.byref(ni), ack)
    if not read:
        return fl, None, None, None
    return fl, list(items), (C.string_at(svl, tot.value * 2) if svl.value and tot.value else b""), \
        (nv.value, ni.value, C.string_at(vbo, nv.value * 2) if vbo.value and nv.value else b"", C.string_at(ibo, ni.value * 4) if ibo.value and ni.value else b"")


def drain(lib, s, n_chunks, manage):
    pub, idle = {}, 0
    for _ in range(600):
        manage(s)
        new = 0
        for i in range(n_chunks):
            fl, items, svl, mesh = published(lib, s, i, 1)
            if fl & 1:
                pub[(i, "svl")] = (items, svl)
                new += 1
            if fl & 2:
                pub[(i, "mesh")] = mesh
                new += 1
        pending = sum(lib.vr_chunk_pending(s, C.c_uint32(i)) for i in range(n_chunks))
        idle = idle + 1 if (new == 0 and pending == 0) else 0
        if idle >= 2:
            return pub
        time.sleep(0.03)
    raise AssertionError("dispatcher did not drain")


def test_dropin_bookkeeping_matches_the_reference_dispatcher(lib):
    w = worldgen.World(2024, 4, (2, 1, 2))
    mesh_ids = [0, 1, 4, 5]
    a, b = make_set(lib, w, mesh_ids), make_set(lib, w, mesh_ids)
    pa = drain(lib, a, w.n_chunks, lib.vr_manage_cpu)          # the reference's own loop
    pb = drain(lib, b, w.n_chunks, lib.vr_manage)              # same entry point, the drop-in behind it
    assert pa.keys() == pb.keys() and len(pa) >= w.n_chunks
    for k in pa:
        assert pa[k] == pb[k], k
    for (x, y, z, r, v) in [(15, 10, 17, 4, 63), (32, 6, 32, 5, 0), (20, 15, 20, 3, 17)]:
        lib.vr_edit_sphere(a, x, y, z, r, v)
        lib.vr_edit_sphere(b, x, y, z, r, v)
    time.sleep(0.12)                                           # the 100 ms per-chunk throttle (chunkset.c:309)
    pa = drain(lib, a, w.n_chunks, lib.vr_manage_cpu)
    pb = drain(lib, b, w.n_chunks, lib.vr_manage)
    assert pa.keys() == pb.keys() and len(pa) >= 4
    for k in pa:
        assert pa[k] == pb[k], k


def test_rle_only_chunks_are_uploaded_as_streams(lib):
    """Every chunk compressed before the first pass (the state of a world that was just loaded: `voxels == NULL`, only
    `rle`): the residency pass sends the streams themselves (vp_multi_upload_chunks_rle, decoded on the other side), null
    chunks as null -- no host decode.  Published geometry must equal the reference dispatcher's."""
    w = worldgen.World(5, 4, (2, 1, 2))
    a, b = make_set(lib, w, [0, 3]), make_set(lib, w, [0, 3])
    lib.vr_world_compress_all(a)
    lib.vr_world_compress_all(b)
    pa = drain(lib, a, w.n_chunks, lib.vr_manage_cpu)
    pb = drain(lib, b, w.n_chunks, lib.vr_manage)
    assert pa.keys() == pb.keys() and len(pa) >= w.n_chunks
    for k in pa:
        assert pa[k] == pb[k], k


def check_published_equals_final_world(lib, s, w, ignore_shadow_bit=False):
    """ignore_shadow_bit: the shadow bit of a splat comes from height-map entry x + y (shadow.h:45-63), which an edit in a
    DIFFERENT chunk of the same z row can change without making this chunk dirty -- the engine itself leaves such splats
    stale until the chunk is rebuilt for another reason.  Positions and colours must always be current."""
    for i in range(w.n_chunks):
        out = np.zeros((w.R + 1) ** 3 * 5, np.int16)
        items = (C.c_uint32 * 5)()
        n = lib.vr_chunk_splat(s, C.c_uint32(i), helpers.vp(out), C.c_uint32(out.size), items)
        fl, got_items, got_svl, _ = published(lib, s, i, 0)
        assert got_items == list(items), i
        want, got = out[:n].copy(), np.frombuffer(got_svl, np.int16).copy()
        if ignore_shadow_bit:
            want[3::4] &= ~np.int16(64)
            got[3::4] &= ~np.int16(64)
        assert np.array_equal(got, want), i


def test_edit_between_the_two_passes_is_not_lost(lib):
    """Deterministic form of the race: the dispatcher's test seam runs an edit exactly between its residency pass (which
    uploads the voxels of dirty chunks) and its selection pass.  The edited chunks must come out with the edit -- round 1
    sampled `dirty` in the first pass and cleared it in the second, so such an edit was rebuilt from the stale device copy
    and then forgotten."""
    w = worldgen.World(99, 4, (2, 1, 2))
    s = make_set(lib, w)
    drain(lib, s, w.n_chunks, lib.vr_manage)
    hook_t = C.CFUNCTYPE(None, C.c_void_p)
    fired = []

    def between(set_ptr):
        if not fired:                                          # once: a sphere across a chunk corner, into chunks that are dirty already
            fired.append(1)
            lib.vr_edit_sphere(C.c_void_p(set_ptr), 16, 9, 16, 4, 63)

    hook = hook_t(between)
    seam = C.c_void_p.in_dll(lib, "vp_chunkset_manage_between_passes")
    lib.vr_edit_sphere(s, 17, 10, 15, 3, 21)                   # makes the same chunks dirty BEFORE the pass: they get uploaded in pass 1
    time.sleep(0.12)
    seam.value = C.cast(hook, C.c_void_p).value
    try:
        lib.vr_manage(s)
    finally:
        seam.value = None
    assert fired
    time.sleep(0.12)
    drain(lib, s, w.n_chunks, lib.vr_manage)
    check_published_equals_final_world(lib, s, w)


def test_no_edit_is_lost_while_the_dispatcher_runs(lib):
    # 32^3 chunks and a bounded number of two-valued edits: the reference's rle_compress overflows its scratch once a chunk
    # has more than N/4 runs (SURVEY 8a' u4), which noisy edits of tiny chunks reach quickly
    w = worldgen.World(77, 5, (2, 1, 2))
    s = make_set(lib, w)
    drain(lib, s, w.n_chunks, lib.vr_manage)
    X, Y, Z = w.dims
    stop = threading.Event()
    edits = [0]

    def editor():                                              # the game thread: brush edits, 400 of them
        rng = np.random.default_rng(3)
        while not stop.is_set() and edits[0] < 400:
            x, y, z = int(rng.integers(0, X)), int(rng.integers(2, Y)), int(rng.integers(0, Z))
            lib.vr_edit_sphere(s, x, y, z, int(rng.integers(1, 4)), int(rng.choice([0, 63])))
            edits[0] += 1
            time.sleep(0.002)

    def consumer():                                            # the GL thread: acknowledges what was published
        while not stop.is_set():
            for i in range(w.n_chunks):
                published(lib, s, i, 1, read=False)
            time.sleep(0.001)

    threads = [threading.Thread(target=editor), threading.Thread(target=consumer)]
    for t in threads:
        t.start()
    t_end = time.time() + 2.5
    passes = 0
    while time.time() < t_end:                                 # the mesher thread: the drop-in dispatcher, back to back
        lib.vr_manage(s)
        passes += 1
    stop.set()
    for t in threads:
        t.join()
    assert edits[0] > 50 and passes > 20
    time.sleep(0.12)
    drain(lib, s, w.n_chunks, lib.vr_manage)
    # every chunk's last published splat list is the list of the final world
    check_published_equals_final_world(lib, s, w, ignore_shadow_bit=True)
