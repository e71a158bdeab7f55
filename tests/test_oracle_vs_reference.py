"""Pins the oracle restatement (oracle/vox_oracle.c) to the compiled, unmodified reference
(oracle/_ref/libvoxref.so).  Runs wherever oracle/_ref was built (this container); skipped otherwise."""
import numpy as np
import pytest

import helpers
from voxplat_b200 import worldgen

pytestmark = pytest.mark.skipif(not helpers.ref_available(), reason="oracle/_ref/libvoxref.so not built")

CASES = [
    (2, (1, 0, 0), 0.3), (3, (1, 1, 1), 0.5), (4, (1, 1, 2), 0.1), (4, (2, 0, 1), 0.9), (5, (1, 1, 1), 0.4),
    (6, (1, 0, 1), 0.25), (7, (1, 0, 0), 0.04),          # the chunk sizes BASELINE configs C2-C4 (64) and C5 (128) use
]


@pytest.mark.parametrize("rb,bits,density", CASES)
def test_random_worlds_splat_and_mesh(rb, bits, density):
    w = helpers.random_world(100 + rb, rb, bits, density=density)
    o, r = helpers.OracleWorld(w), helpers.RefWorld(w)
    for cid in range(w.n_chunks):
        ga, ia = r.splat(cid)
        gb, ib = o.splat(cid)
        assert np.array_equal(ia, ib), cid
        assert np.array_equal(ga, gb), cid
        va, xa = r.mesh(cid)
        vb, xb = o.mesh(cid)
        assert np.array_equal(va, vb), cid
        assert np.array_equal(xa, xb), cid


@pytest.mark.parametrize("rb,bits", [(5, (2, 1, 2)), (6, (1, 1, 1)), (7, (1, 0, 1))])
def test_terrain_world_hashes(rb, bits):
    w = worldgen.World(1234, rb, bits)
    o, r = helpers.OracleWorld(w), helpers.RefWorld(w)
    ids = np.arange(w.n_chunks, dtype=np.uint32)
    for mode in (0, 1):
        _, ha, ca = r.rebuild(ids, mode, 4)
        _, hb, cb = o.rebuild(ids, mode, 4)
        assert np.array_equal(ca, cb)
        assert np.array_equal(ha, hb)


def test_rle_codec():
    import ctypes as C
    lib = helpers.ref_lib()
    rng = np.random.default_rng(5)
    for n, p in [(4096, 0.02), (32768, 0.05), (32768, 0.0), (4096, 0.06)]:
        # runs must stay below n/4 words for the reference (its scratch is undersized, SURVEY 8a' u4)
        d = np.repeat(rng.integers(0, 256, n // 8).astype(np.uint8), 8)
        flip = rng.random(n) < p * 0.5
        d[flip] = rng.integers(0, 256, int(flip.sum()))
        out = np.zeros(n + 1, np.uint32)
        k = lib.vr_rle_compress(helpers.vp(d), C.c_uint32(n), helpers.vp(out), C.c_uint32(out.size))
        enc = helpers.rle_encode(d)
        assert k == enc.size and np.array_equal(out[:k], enc)
        assert np.array_equal(helpers.rle_decode(enc, n), d)
        back = np.zeros(n, np.uint8)
        lib.vr_rle_compress(helpers.vp(d), C.c_uint32(n), None, C.c_uint32(0))     # sets the reference's scratch length (u5)
        lib.vr_rle_decompress(helpers.vp(enc), helpers.vp(back), C.c_uint32(n))
        assert np.array_equal(back, d)
