"""Multi-GPU driver: the world is sharded into z-slabs of chunk rows, one process (rank) per GPU.

New relative to the reference (which is a single process): chunk ids are z-major (chunkset.c:124-126), so a
slab is a contiguous id range and every chunk is rebuilt by exactly one rank.  The only data dependence
between ranks is the 1-voxel border plane: the cull (mesher.c:391-432) needs z-slice 0 of the chunk row
above the slab; mesh AO (mesher.c:119-171) additionally needs z-slice R-1 of the row below.  Those planes
travel with NCCL point-to-point (torch.distributed, NVLink) between slab neighbours -- no other collective
is on the data path.  PyTorch is plumbing here (device buffers for NCCL, streams); pack / unpack and all
compute are kernels of the C-ABI library.
"""
import contextlib

import numpy as np


def slab_rows(nz, world_size, rank):
    """Chunk rows [z0,z1) owned by `rank`: equal contiguous shares (nz and world_size are powers of two)."""
    assert nz % world_size == 0, "chunk rows must divide evenly over the ranks"
    rows = nz // world_size
    return rank * rows, (rank + 1) * rows


def halo_schedule(rank, world_size, mesh):
    """Point-to-point plan of one border exchange as (op, which, peer) tuples, non-periodic chain.
    which 0: z-slice 0 of a slab's FIRST row, flows to the rank below (its +z halo).
    which 1: z-slice R-1 of a slab's LAST row, flows to the rank above (its -z halo; mesh AO only)."""
    plan = []
    if rank > 0:
        plan.append(("send", 0, rank - 1))
        if mesh:
            plan.append(("recv", 1, rank - 1))
    if rank < world_size - 1:
        plan.append(("recv", 0, rank + 1))
        if mesh:
            plan.append(("send", 1, rank + 1))
    return plan


class SlabRebuilder:
    """One rank's slab: owns a Context restricted to its chunk rows and the border exchange.

    `store` is anything with halo_plane_bytes(), halo_pack(which, ptr), halo_unpack(which, ptr) -- the CUDA
    Context in production, a host stand-in in the gloo tests."""

    def __init__(self, store, rank, world_size, make_buffer, dist=None, group=None):
        self.store, self.rank, self.world_size = store, rank, world_size
        self.dist, self.group = dist, group
        n = store.halo_plane_bytes()
        # one buffer per (direction, which); allocated once, reused every exchange
        self.buf = {("send", 0): make_buffer(n), ("recv", 0): make_buffer(n),
                    ("send", 1): make_buffer(n), ("recv", 1): make_buffer(n)}
        # The CUDA store packs / unpacks on its own border stream; the transport must be ordered against that stream,
        # which is what torch.distributed does with the CURRENT stream of the calling thread.
        self._border = None
        handle = store.border_stream_handle() if hasattr(store, "border_stream_handle") else 0
        if handle:
            import torch
            self._border = torch.cuda.ExternalStream(handle)

    def _on_border_stream(self):
        if self._border is None:
            return contextlib.nullcontext()
        import torch
        return torch.cuda.stream(self._border)

    def exchange_begin(self, mesh=False):
        """Pack my border planes and start the swap with the slab neighbours; returns the pending requests.
        NCCL runs the transfer on its own stream, so kernels queued next on the context stream (the rebuild of the
        chunks that do not read a ghost row) overlap it."""
        if self.world_size == 1:
            return None
        plan = halo_schedule(self.rank, self.world_size, mesh)
        with self._on_border_stream():
            for op, which, _ in plan:
                if op == "send":
                    self.store.halo_pack(which, self.buf[(op, which)].data_ptr())
            ops = []
            for op, which, peer in plan:
                fn = self.dist.isend if op == "send" else self.dist.irecv
                ops.append(self.dist.P2POp(fn, self.buf[(op, which)], peer, group=self.group))
            reqs = self.dist.batch_isend_irecv(ops) if ops else []
        return plan, reqs

    def exchange_finish(self, pending):
        """Wait for the swap started by exchange_begin and unpack the received planes into the ghost rows."""
        if pending is None:
            return 0
        plan, reqs = pending
        with self._on_border_stream():
            for req in reqs:
                req.wait()
            for op, which, _ in plan:
                if op == "recv":
                    self.store.halo_unpack(which, self.buf[(op, which)].data_ptr())
        return len(plan)

    def exchange_halos(self, mesh=False):
        """Pack my border planes, swap them with the slab neighbours, unpack into the ghost rows."""
        return self.exchange_finish(self.exchange_begin(mesh))

    def rebuild_step(self, mesh=True):
        """One device-resident rebuild of the prepared batch with the border work -- pack, NCCL transfer, unpack and the
        rebuild of the chunks that read a ghost row (part 1) -- on the store's border stream, beside the chunks that do
        not need it (part 0, context stream).  store = a voxplat_b200.Context."""
        pending = self.exchange_begin(mesh)
        self.store.rebuild_device_part(0)
        self.exchange_finish(pending)
        self.store.rebuild_device_part(1)


def near_camera_flags(ids, root_bitw, max_bitw, camera=(64.0, 128.0, 64.0), radius=512.0):
    """make_mesh rule of the reference's game loop (game.c:612-618): a chunk is meshed when its centre is
    closer than 512 voxels to the LOD origin (initial camera, game.c:270-272)."""
    ids = np.asarray(ids, dtype=np.int64)
    R = 1 << root_bitw
    cx = ids & ((1 << max_bitw[0]) - 1)
    cy = (ids >> max_bitw[0]) & ((1 << max_bitw[1]) - 1)
    cz = ids >> (max_bitw[0] + max_bitw[1])
    d = np.sqrt((cx * R + R / 2 - camera[0]) ** 2 + (cy * R + R / 2 - camera[1]) ** 2 + (cz * R + R / 2 - camera[2]) ** 2)
    return d < radius
