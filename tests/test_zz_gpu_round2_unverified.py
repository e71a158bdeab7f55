"""GPU tests written in round 2 AFTER the round's GPU budget was spent: they compile, their CPU halves (the compiled
reference, the checkers, the harness) run here, but none of them has been executed on a B200 yet.  The file sorts last so
that `pytest -x -m gpu` runs every test that HAS been verified on hardware before it reaches these.

  * chunk 128 / 64 splat + mesh buffers against the COMPILED reference (not only the restatement)
  * device edits at chunk 128, and on two z-slab contexts (edits centred in the other slab / on the cut)
  * vp_raycast against chunkset_edit_raycast_until_solid
  * the flat RLE codec at any length, with the reference's run split at 0xFFFFFF
  * vp_build_lod_nodes against the reference's own gfx_update_svl (compiled unmodified, GL calls captured)
  * rays, node buffers and an edit burst against the committed golden fixtures (no oracle/_ref needed on the box)

tests/test_unverified_gpu_tests_dryrun.py runs every one of these functions on CPU with the oracle-backed stand-in context.
"""
import ctypes as C

import numpy as np
import pytest

import helpers
import voxplat_b200 as vpb
from voxplat_b200 import worldgen
from test_gpu_splat import upload_world
from test_gpu_edit import test_edit_sphere_matches_compiled_reference as _edit_vs_reference

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not helpers.ref_available(), reason="oracle/_ref/libvoxref.so not built")


@needs_ref
@pytest.mark.parametrize("rb,bits", [(4, (2, 1, 3)), (5, (2, 1, 2))])
def test_lod_nodes_match_the_reference_gfx_update_svl(rb, bits):
    """The device gather against the COMPILED reference: its dispatcher publishes every chunk, its gfx_update_svl
    (gfx/vsplat.c:197-338, GL calls captured by oracle/gfx_shim) builds the node buffers."""
    w = worldgen.World(2718, rb, bits)
    r = helpers.RefWorld(w)
    r.run_engine_with_gfx()
    ctx = vpb.Context(rb, bits)
    try:
        upload_world(ctx, w)
        ctx.rebuild_batch(np.arange(w.n_chunks, dtype=np.uint32), vpb.VP_REBUILD_SPLAT)
        for lod in range(5):
            nodes, buf, ms = ctx.build_lod_nodes(lod)
            assert len(nodes) == 1 << sum(b - min(lod, b) for b in bits)
            for node in range(len(nodes)):
                want_n, want = r.node_buffer(lod, node)
                assert nodes["items"][node] == want_n, (lod, node)
                if want_n:
                    off = int(nodes["offset"][node])
                    assert np.array_equal(buf[off:off + want_n * 2].view(np.int16), want), (lod, node)
    finally:
        ctx.close()


@needs_ref
@pytest.mark.parametrize("rb,bits,kind", [(7, (1, 0, 1), "terrain"), (7, (1, 0, 0), "random"), (6, (1, 1, 1), "random")])
def test_large_chunks_against_the_compiled_reference(rb, bits, kind):
    """Chunk 128 (BASELINE config C5) and 64 against the COMPILED reference itself, not only the restatement."""
    w = worldgen.World(4321, rb, bits) if kind == "terrain" else helpers.random_world(4321 + rb, rb, bits, density=0.06, null_frac=0.0)
    r = helpers.RefWorld(w)
    ctx = vpb.Context(rb, bits, splat_arena_bytes=1 << 30, mesh_arena_bytes=2 << 30)
    try:
        nn = w.nonnull_ids()
        ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
        ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
        ids = np.arange(w.n_chunks, dtype=np.uint32)
        res, splat, mesh = ctx.rebuild_batch(ids, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
        for k in range(w.n_chunks):
            g, it = r.splat(k)
            off = int(res["svl_offset"][k])
            assert np.array_equal(res["svl_items"][k], it), k
            assert np.array_equal(splat[off:off + g.size * 2].view(np.int16), g), k
            v, x = r.mesh(k)
            vo, io = int(res["vbo_offset"][k]), int(res["ibo_offset"][k])
            assert res["vbo_items"][k] == v.size and res["ibo_items"][k] == x.size, k
            assert np.array_equal(mesh[vo:vo + v.size * 2].view(np.int16), v), k
            assert np.array_equal(mesh[io:io + x.size * 4].view(np.uint32), x), k
    finally:
        ctx.close()


@needs_ref
def test_edit_sphere_chunk_128_matches_compiled_reference():
    _edit_vs_reference(7, (1, 0, 2))


def test_flat_codec_any_length_and_run_split():
    """rle.h:7 takes any length, and a run stops when its count reaches 0xFFFFFF (rle.c:62): lengths that are not a multiple
    of 16, single bytes, and runs longer than 24 bits -- against the oracle encoder and, where the reference's scratch
    allows it (SURVEY 8a' u4), the compiled reference itself."""
    import ctypes as C
    ctx = vpb.Context(5, (0, 0, 0), rle_arena_bytes=1 << 30)
    try:
        rng = np.random.default_rng(8)
        cases = [np.array([9], np.uint8), np.array([0, 0, 0, 5, 5], np.uint8), np.full(17, 3, np.uint8),
                 np.repeat(rng.integers(0, 256, 123).astype(np.uint8), 7)[:851],
                 np.concatenate([np.zeros(1001, np.uint8), np.full(31, 4, np.uint8)])]
        big = np.zeros(0x2000005, np.uint8)                       # 33.5 M bytes: a run of zeros longer than 0xFFFFFF (split + remainder) ...
        big[0xFFFFFF + 100:0xFFFFFF + 200] = 77                   # ... another value, a second long run of zeros ...
        big[-3:] = 5                                              # ... and a tail that is not a multiple of 16
        cases.append(big)
        big2 = np.zeros(0x2000005, np.uint8)                      # zeros for two full runs and a remainder, across both 0xFFFFF0-byte segments
        big2[-3:] = 5
        cases.append(big2)
        cases.append(np.full(0xFFFFFF * 2, 1, np.uint8))          # exactly two maximal runs, no remainder word
        for d in cases:
            enc = ctx.rle_compress(d)
            want = helpers.rle_encode(d)
            assert np.array_equal(enc, want), d.size
            assert (enc[:-1] & 0xFFFFFF).max() <= 0xFFFFFF and enc[-1] == 0
            assert np.array_equal(ctx.rle_decompress(want, d.size), d), d.size
            if helpers.ref_available() and want.size <= d.size // 4:
                lib = helpers.ref_lib()
                out = np.zeros(want.size + 8, np.uint32)
                k = lib.vr_rle_compress(helpers.vp(d), C.c_uint32(d.size), helpers.vp(out), C.c_uint32(out.size))
                assert k == enc.size and np.array_equal(out[:k], enc), d.size
    finally:
        ctx.close()


@needs_ref
def test_edit_sphere_on_slab_contexts():
    """Two z-slabs, every edit applied to BOTH contexts (each writes the cells and height-map rows it holds): edits
    centred in the other slab, straddling the cut and at the far end of the world must leave every owned chunk and every
    owned height-map row equal to the compiled reference's -- and must not touch memory past a slab's rows."""
    rb, bits = 4, (1, 1, 2)
    w = worldgen.World(77, rb, bits)
    r = helpers.RefWorld(w)
    lib = r.lib
    lib.vr_chunk_voxels.restype = C.c_void_p
    lib.vr_shadow_ptr.restype = C.c_void_p
    per_row, nz, R = 1 << (bits[0] + bits[1]), 1 << bits[2], w.R
    ctxs = [vpb.Context(rb, bits, slab=(0, nz // 2)), vpb.Context(rb, bits, slab=(nz // 2, nz))]
    try:
        for k, ctx in enumerate(ctxs):
            z0, z1 = k * nz // 2, (k + 1) * nz // 2
            own = np.arange(z0 * per_row, z1 * per_row, dtype=np.uint32)
            nn = own[w.solid[own] > 0]
            ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
            ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
        X, Y, Z = w.dims
        cut = nz // 2 * R
        edits = [(5, 9, cut + 10, 4, 63), (20, 12, cut - 12, 5, 63), (9, 10, cut, 6, 63), (9, 14, cut - 1, 3, 0),
                 (3, 8, Z - 2, 5, 63), (12, 9, 1, 4, 63), (16, 11, cut + 2, 4, 17)]
        for (x, y, z, rad, v) in edits:
            lib.vr_edit_sphere(r.set, x, y, z, rad, v)
            for ctx in ctxs:
                ctx.edit_sphere(x, y, z, rad, v)
        want_sh = np.frombuffer(C.string_at(lib.vr_shadow_ptr(r.set), w.shw * Z * 2), np.uint16).reshape(Z, w.shw)
        for k, ctx in enumerate(ctxs):
            z0, z1 = k * nz // 2, (k + 1) * nz // 2
            own = np.arange(z0 * per_row, z1 * per_row, dtype=np.uint32)
            got = ctx.download_chunks_dense(own)
            for j, cid in enumerate(own):
                want = np.frombuffer(C.string_at(lib.vr_chunk_voxels(r.set, C.c_uint32(int(cid))), w.N), np.uint8)
                assert np.array_equal(got[j], want), (k, cid)
            top = min(Z, z1 * R + 17)                 # the slab also keeps the 17 rows of reach past its last chunk row
            rows = ctx.download_shadow_rows(z0 * R, top).reshape(-1, w.shw)
            assert np.array_equal(rows, want_sh[z0 * R:top]), k
    finally:
        for ctx in ctxs:
            ctx.close()


def reference_rays(r, origins, vectors):
    n = len(origins)
    vox, coords, nrm = np.zeros(n, np.uint8), np.zeros((n, 3), np.uint32), np.zeros((n, 3), np.int8)
    for i in range(n):
        o = (C.c_float * 3)(*[float(x) for x in origins[i]])
        v = (C.c_float * 3)(*[float(x) for x in vectors[i]])
        c = (C.c_uint32 * 3)()
        m = (C.c_int8 * 3)()
        vox[i] = r.lib.vr_raycast(r.set, o, v, c, m)
        coords[i] = list(c)
        nrm[i] = list(m)
    return vox, coords, nrm


@needs_ref
@pytest.mark.parametrize("rb,bits,kind", [(5, (2, 1, 2), "terrain"), (4, (2, 1, 2), "random"), (6, (1, 0, 1), "terrain")])
def test_raycast_matches_compiled_reference(rb, bits, kind):
    w = worldgen.World(31, rb, bits) if kind == "terrain" else helpers.random_world(31, rb, bits, density=0.02, null_frac=0.3)
    r = helpers.RefWorld(w)
    rng = np.random.default_rng(12)
    X, Y, Z = w.dims
    n = 400
    o = np.stack([rng.uniform(0, X, n), rng.uniform(0, Y, n), rng.uniform(0, Z, n)], axis=1).astype(np.float32)
    v = rng.normal(size=(n, 3)).astype(np.float32)
    # camera-like rays: from above, looking down at an angle
    o[:100, 1] = Y - 1.5
    v[:100, 1] = -np.abs(v[:100, 1]) - 0.2
    # axis-aligned and planar rays (zero components)
    v[100:110] = [0, -1, 0]
    v[110:120] = [1, 0, 0]
    v[120:130] = [0, 0, -1]
    v[130:150, 2] = 0
    # rays that start outside the world, on both sides
    o[150:170, 0] = X + rng.uniform(1, 20, 20).astype(np.float32)
    v[150:170, 0] = -np.abs(v[150:170, 0]) - 0.1
    o[170:190, 2] = -rng.uniform(1, 20, 20).astype(np.float32)
    v[170:190, 2] = np.abs(v[170:190, 2]) + 0.1
    # rays that leave through the x = 0 face
    o[190:210, 0] = rng.uniform(0, 3, 20).astype(np.float32)
    v[190:210] = [-1, 0.01, 0.02]
    want = reference_rays(r, o, v)
    ctx = vpb.Context(rb, bits)
    try:
        upload_world(ctx, w)
        got = ctx.raycast(o, v)
    finally:
        ctx.close()
    assert want[0].any() and not want[0].all()                 # hits and misses both occur
    for k, name in enumerate(("voxel", "coord", "normal")):
        bad = np.nonzero((got[k] != want[k]).reshape(n, -1).any(axis=1))[0]
        assert len(bad) == 0, (name, bad[:5], got[k][bad[:5]], want[k][bad[:5]])


# ---- the same three rows against the committed golden fixtures (tests/golden/, generated from the compiled reference):
# ---- these do not need oracle/_ref on the box
def test_device_rays_nodes_edits_match_golden():
    from test_golden import load_secondary
    for name in ("rays_terrain_r32", "rays_random_r16"):
        w, z = load_secondary(name)
        ctx = vpb.Context(w.root_bitw, w.max_bitw)
        try:
            upload_world(ctx, w)
            vox, coords, nrm = ctx.raycast(z["origins"], z["vectors"])
        finally:
            ctx.close()
        assert np.array_equal(vox, z["voxels"]) and np.array_equal(coords, z["coords"]) and np.array_equal(nrm, z["normals"]), name
    w, z = load_secondary("nodes_terrain_r16")
    ctx = vpb.Context(w.root_bitw, w.max_bitw)
    try:
        upload_world(ctx, w)
        ctx.rebuild_batch(np.arange(w.n_chunks, dtype=np.uint32), vpb.VP_REBUILD_SPLAT)
        at, by_lod = 0, {}
        for lod, node, n in z["nodes"]:
            if int(lod) not in by_lod:
                by_lod[int(lod)] = ctx.build_lod_nodes(int(lod))
            nodes, buf, _ = by_lod[int(lod)]
            assert nodes["items"][int(node)] == n, (lod, node)
            off = int(nodes["offset"][int(node)])
            assert np.array_equal(buf[off:off + int(n) * 2].view(np.int16), z["data"][at:at + n]), (lod, node)
            at += int(n)
    finally:
        ctx.close()
    w, z = load_secondary("edits_terrain_r32")
    ctx = vpb.Context(w.root_bitw, w.max_bitw)
    try:
        upload_world(ctx, w)
        offs = z["dirty_offsets"]
        for k, (x, y, zz, r, v) in enumerate(z["edits"].tolist()):
            dirty = ctx.edit_sphere(x, y, zz, r, v)
            assert sorted(dirty.tolist()) == z["dirty"][offs[k]:offs[k + 1]].tolist(), k
        assert np.array_equal(ctx.download_chunks_dense(np.arange(w.n_chunks, dtype=np.uint32)), z["dense"])
        assert np.array_equal(ctx.download_shadow_rows(0, w.dims[2]), z["shadow"])
    finally:
        ctx.close()
