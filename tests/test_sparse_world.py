"""helpers.SparseWorld (the host world of bench.py's parity guard): a few chunk rows of a world must give the checker the
same buffers as the whole world, for every row it declares checkable -- also when the world is a smaller world repeated
along z (bench.py's weak scaling)."""
import numpy as np
import pytest

import helpers
from voxplat_b200 import worldgen


def checkers(w):
    out = [helpers.OracleWorld(w)]
    if helpers.ref_available():
        out.append(helpers.RefWorld(w))
    return out


def test_sparse_rows_match_whole_world():
    rb, bits = 5, (2, 1, 3)
    w = worldgen.World(1234, rb, bits)
    sw = helpers.SparseWorld(1234, rb, bits, [2, 3, 4, 6, 7])
    assert sw.checkable(False) == [2, 3, 6, 7] and sw.checkable(True) == [3, 7]
    per_row = 1 << (bits[0] + bits[1])
    for full, part in zip(checkers(w), checkers(sw)):
        for mesh, rows in ((0, sw.checkable(False)), (1, sw.checkable(True))):
            ids = np.concatenate([np.arange(r * per_row, (r + 1) * per_row, dtype=np.uint32) for r in rows])
            _, h1, c1 = full.rebuild(ids, mesh)
            _, h2, c2 = part.rebuild(ids, mesh)
            assert np.array_equal(h1, h2) and np.array_equal(c1, c2)


def test_sparse_repeated_world():
    rb, base, bits = 4, (2, 1, 2), (2, 1, 4)
    per_row = 1 << (bits[0] + bits[1])
    wb = worldgen.World(7, rb, base)
    # the whole repeated world, built by hand
    ids = np.arange(1 << sum(bits), dtype=np.uint32)
    src = ((ids // per_row) % (1 << base[2])) * per_row + ids % per_row
    w = worldgen.World(7, rb, bits, dense=wb.dense[src])
    sw = helpers.SparseWorld(7, rb, bits, [3, 4, 5], repeat_bits=base)
    rows = np.arange(4 * per_row, 5 * per_row, dtype=np.uint32)
    for mesh in (0, 1):
        _, h1, c1 = helpers.OracleWorld(w).rebuild(rows, mesh)
        _, h2, c2 = helpers.OracleWorld(sw).rebuild(rows, mesh)
        assert np.array_equal(h1, h2) and np.array_equal(c1, c2)


def test_timing_runs_do_not_hash():
    w = worldgen.World(3, 4, (1, 1, 1))
    ids = np.arange(w.n_chunks, dtype=np.uint32)
    for chk in checkers(w):
        _, h, c = chk.rebuild(ids, 0, hashed=False)
        _, h2, c2 = chk.rebuild(ids, 0)
        assert not h.any() and h2.any() and np.array_equal(c, c2)
