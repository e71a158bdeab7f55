"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 4) gloo groups run the same
SlabRebuilder.exchange_halos schedule as the NCCL run, against a host stand-in for the device store."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from voxplat_b200 import slab


class HostStore:
    """Stand-in for the CUDA Context: border planes are deterministic functions of (rank, which)."""

    def __init__(self, rank, nbytes):
        self.rank, self.nbytes = rank, nbytes
        self.ghost = {}
        self.log = []

    def halo_plane_bytes(self):
        return self.nbytes

    @staticmethod
    def plane(rank, which, n):
        return ((np.arange(n) * 7 + rank * 31 + which * 101) % 251).astype(np.uint8)

    def halo_pack(self, which, ptr):
        buf = (np.ctypeslib.as_array((__import__("ctypes").c_uint8 * self.nbytes).from_address(ptr)))
        buf[:] = self.plane(self.rank, which, self.nbytes)

    def halo_unpack(self, which, ptr):
        buf = (np.ctypeslib.as_array((__import__("ctypes").c_uint8 * self.nbytes).from_address(ptr)))
        self.ghost[which] = buf.copy()
        self.log.append(("unpack", which))

    def rebuild_device_part(self, part):
        # the overlapped step: part 0 must be issued before any plane is unpacked, part 1 after all of them
        self.log.append(("part", part))


def _worker(rank, world_size, port, mesh, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    n = 4096
    store = HostStore(rank, n)
    rb = slab.SlabRebuilder(store, rank, world_size, lambda k: torch.empty(k, dtype=torch.uint8), dist=dist)
    ops = rb.exchange_halos(mesh=mesh)
    ok = True
    first = dict(store.ghost)
    store.ghost.clear(); store.log.clear()
    rb.rebuild_step(mesh=mesh)                       # same planes again, with the rebuild parts around the unpack
    ok &= set(store.ghost) == set(first) and all(np.array_equal(store.ghost[k], first[k]) for k in first)
    kinds = [e for e in store.log]
    ok &= kinds[0] == ("part", 0) and kinds[-1] == ("part", 1) and all(k[0] == "unpack" for k in kinds[1:-1])
    # +z halo: plane 0 of the rank above; -z halo (mesh only): plane 1 of the rank below
    if rank < world_size - 1:
        ok &= 0 in store.ghost and np.array_equal(store.ghost[0], HostStore.plane(rank + 1, 0, n))
    else:
        ok &= 0 not in store.ghost
    if mesh and rank > 0:
        ok &= 1 in store.ghost and np.array_equal(store.ghost[1], HostStore.plane(rank - 1, 1, n))
    else:
        ok &= 1 not in store.ghost
    expected_ops = len(slab.halo_schedule(rank, world_size, mesh))
    out[rank] = int(ok and ops == expected_ops)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world_size,mesh", [(2, False), (2, True), (4, True)])
def test_halo_exchange_schedule(world_size, mesh):
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world_size)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world_size, port, mesh, out)) for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * world_size


def test_slab_rows_and_schedule():
    assert slab.slab_rows(32, 1, 0) == (0, 32)
    assert [slab.slab_rows(128, 8, r) for r in (0, 7)] == [(0, 16), (112, 128)]
    assert slab.halo_schedule(0, 1, True) == []
    assert slab.halo_schedule(0, 2, False) == [("recv", 0, 1)]
    assert slab.halo_schedule(1, 2, True) == [("send", 0, 0), ("recv", 1, 0)]
    # every send has a matching recv on the peer
    for ws in (2, 4, 8):
        for mesh in (False, True):
            sends = {(r, p, w) for r in range(ws) for op, w, p in slab.halo_schedule(r, ws, mesh) if op == "send"}
            recvs = {(p, r, w) for r in range(ws) for op, w, p in slab.halo_schedule(r, ws, mesh) if op == "recv"}
            assert sends == recvs


def test_near_camera_rule():
    ids = np.arange(4096)
    near = slab.near_camera_flags(ids, 6, (5, 2, 5))
    assert near[0] and not near[4095] and 200 < near.sum() < 400      # game.c:612-618, camera (64,128,64)
