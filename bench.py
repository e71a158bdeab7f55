#!/usr/bin/env python
"""bench.py -- Gvoxel/s culled+meshed of a full-world chunk rebuild (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # the reference's own CPU code on the host cores

A step = one full rebuild of every chunk of the world: cull + 5 LOD splat lists for ALL chunks and the
near-field quad mesh for the chunks within 512 voxels of the initial camera.
N = 1: the default world 2048 x 256 x 2048 voxels, 64^3 chunks.  N > 1: weak scaling, the world grows
along z to 2048 x 256 x (2048 N), one z-slab of 32 chunk rows per GPU, border planes exchanged with NCCL.
Inputs are synthetic (voxplat_b200/csrc/vp_worldgen.c, seed 1234) and, at 0.67 GB of voxels per GPU, larger than L2.

The JSON line also carries
  parity   -- outside the timed region: the buffers of each rank's first, last (and one middle) chunk row, taken from the
              timed device-resident step AND from the end-to-end call, hashed (FNV-1a 64) and compared with the compiled
              reference (oracle/_ref; the oracle port if it did not travel) run on the same world; exit code 1 on a mismatch
  extra    -- the other BASELINE.json configs under the same clock: N = 1: C1 (one chunk), C5 (edit bursts, chunk 32 / 128)
              and the mesh-for-all stress figure; N > 1: C3 (8192 x 512 x 8192, strong scaling); N = 8: C4 (16384 x 512 x
              16384 from RLE streams).  --no-extra skips them.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 1234
ROOT_BITW = 6
BASE_BITS = (5, 2, 5)          # 2048 x 256 x 2048 voxels in 64^3 chunks
METRIC = "Gvoxel/s culled+meshed (full-world chunk rebuild)"
DATA = "synthetic (seeded integer worldgen, seed 1234)"
E2E_BLOCKS = int(os.environ.get("VP_BENCH_E2E_BLOCKS", "8"))     # pipeline depth of the end-to-end call

WORKLOADS = {            # chunk-grid bit widths (64^3 chunks); c2 grows along z with the GPU count (weak scaling)
    "c2": None,           # BASELINE config 2: 2048 x 256 x 2048 per GPU
    "c3": (7, 3, 7),      # BASELINE config 3: 8192 x 512 x 8192, fixed world split over the GPUs (strong scaling)
    "c4": (8, 3, 8),      # BASELINE config 4: 16384 x 512 x 16384 (the README's 128 GB map), 8 GPUs
}


def world_bits(workload, n_gpus):
    extra = int(np.log2(n_gpus))
    assert (1 << extra) == n_gpus, "--gpus must be a power of two"
    if WORKLOADS[workload] is not None:
        return WORKLOADS[workload]
    return (BASE_BITS[0], BASE_BITS[1], BASE_BITS[2] + extra)


def workload_name(workload, bits):
    R = 1 << ROOT_BITW
    name = "%dx%dx%d voxels, chunk %d^3, full rebuild: splat(5 LOD) all chunks + mesh within 512 of camera" % (
        (1 << bits[0]) * R, (1 << bits[1]) * R, (1 << bits[2]) * R, R)
    if workload == "c2" and tuple(bits) != tuple(BASE_BITS):
        name += " (the 2048x256x2048 world repeated %d times along z: identical work per GPU)" % (1 << (bits[2] - BASE_BITS[2]))
    return name


def config_dict(workload, n_gpus):
    """The same dict in both arms (native and --impl reference): everything in it follows from the arguments."""
    bits = world_bits(workload, n_gpus)
    nz = 1 << bits[2]
    return {"workload": workload_name(workload, bits), "baseline_config": workload, "chunks": 1 << sum(bits),
            "parallelism": ("z-slabs of %d chunk rows per GPU, NCCL border planes" % (nz // n_gpus)) if n_gpus > 1 else "single GPU",
            "l2": "inputs larger than L2 (0.67 GB of voxels per GPU at c2, no flush needed)"}


# ---------------------------------------------------------------------------------------------------
# clocks: sample SM clock + throttle reasons DURING the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, device_index):
        self.samples, self.reasons, self.stop_flag, self.thread = [], set(), False, None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self.stop_flag = False
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def pause(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join()
            self.thread = None

    def stop(self):
        self.pause()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# host-side world pieces
# ---------------------------------------------------------------------------------------------------
def generate_slab(workload, bits, z0, z1):
    """Dense chunks of chunk rows [z0, min(z1+1, nz)) and the shadow rows the slab needs."""
    from voxplat_b200 import worldgen
    nx, ny, nz = (1 << b for b in bits)
    per_row = nx * ny
    R = 1 << ROOT_BITW
    N = R ** 3
    zg1 = min(z1 + 1, nz)                                   # one extra row: the shadow reach crosses the border
    ids = np.arange(z0 * per_row, zg1 * per_row, dtype=np.uint32)
    if workload == "c2" and tuple(bits) != tuple(BASE_BITS):
        # weak scaling: the world is the base world (one GPU's 2048x256x2048) repeated along z, so every rank rebuilds
        # exactly the same content -- a larger generated world has other terrain statistics (the generator's edge
        # fall-off is relative to the world size) and the per-GPU work would drift with N
        base_rows = 1 << BASE_BITS[2]
        base_ids = ((ids // per_row) % base_rows) * per_row + ids % per_row
        dense, solid = worldgen.gen_chunks(SEED, ROOT_BITW, BASE_BITS, base_ids.astype(np.uint32))
    else:
        dense, solid = worldgen.gen_chunks(SEED, ROOT_BITW, bits, ids)
    ptrs = [0] * (per_row * nz)
    base = dense.ctypes.data
    for k, cid in enumerate(ids):
        if solid[k]:
            ptrs[int(cid)] = base + k * N
    sz0, sz1 = z0 * R, min(nz * R, z1 * R + 17)
    rows = worldgen.shadow_rows(SEED, ROOT_BITW, bits, ptrs, sz0, sz1)
    n_own = (z1 - z0) * per_row
    return ids[:n_own], dense[:n_own], solid[:n_own], rows, sz0


def algorithmic_bytes(own_ids, solid, res, bits):
    """every input byte read once, every output byte written once."""
    R = 1 << ROOT_BITW
    nx, ny, nz = (1 << b for b in bits)
    nn = solid > 0
    splats = int(res["svl_items_total"].astype(np.int64).sum()) // 4
    faces = int(res["vbo_items"].astype(np.int64).sum()) // 16
    meshed = res["vbo_items"] > 0
    # +x,+y,+z halo faces that exist inside the world (one R^2 plane each)
    ids = own_ids.astype(np.int64)
    cx, cy, cz = ids % nx, (ids // nx) % ny, ids // (nx * ny)
    halo_planes = int(((cx + 1 < nx).astype(np.int64) + (cy + 1 < ny) + (cz + 1 < nz)).sum())
    splat_bytes = int(nn.sum()) * R ** 3 + halo_planes * R * R + 12 * splats
    mesh_bytes = int(meshed.sum()) * (R + 2) ** 3 + 64 * faces
    return splat_bytes, mesh_bytes, splats, faces


def helpers_module():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    return helpers


class stdout_to_stderr:
    """The compiled reference logs to stdout (log.c via printf); keep stdout for the ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        try:
            C.CDLL(None).fflush(None)          # the C library's buffered lines belong to stderr too
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


# ---------------------------------------------------------------------------------------------------
# parity guard (outside every timed region): the checker is the compiled reference, never the thing measured
# ---------------------------------------------------------------------------------------------------
def fnv_chunks(lib_fnv, buf, offs, nbytes, offs2=None, nbytes2=None):
    out = np.zeros(len(offs), np.uint64)
    base = buf.ctypes.data
    for k in range(len(offs)):
        h = lib_fnv(C.c_void_p(base + int(offs[k])), C.c_uint64(int(nbytes[k])), C.c_uint64(0))
        if offs2 is not None:
            h = lib_fnv(C.c_void_p(base + int(offs2[k])), C.c_uint64(int(nbytes2[k])), C.c_uint64(h))
        out[k] = h
    return out


def parity_guard(workload, bits, rank, world_size, z0, z1, own_ids, flags, device_out, e2e_out):
    """Compare this rank's first, last (and one middle) chunk rows with the reference run on the same seeded world.
    device_out = (res, splat bytes, mesh bytes) of the timed device-resident step, e2e_out = (ids, res, splat, mesh) of the
    end-to-end call (non-null chunks only).  Returns [checked chunks, chunks next to another rank's slab, ok, detail]."""
    helpers = helpers_module()
    import voxplat_b200 as vpb
    nx, ny, nz = (1 << b for b in bits)
    per_row = nx * ny
    rows = {z0, z1 - 1}
    if per_row <= 1024 and z1 - z0 > 2:
        rows.add((z0 + z1) // 2)
    rows = sorted(rows)
    need = sorted({r + d for r in rows for d in (-1, 0, 1) if 0 <= r + d < nz})
    with stdout_to_stderr():
        sw = helpers.SparseWorld(SEED, ROOT_BITW, bits, need, repeat_bits=BASE_BITS if workload == "c2" else None)
        kind = "reference" if helpers.ref_available() else "port"
        checker = helpers.RefWorld(sw) if kind == "reference" else helpers.OracleWorld(sw)
        ids = np.concatenate([np.arange(r * per_row, (r + 1) * per_row, dtype=np.uint32) for r in rows])
        k = (ids - own_ids[0]).astype(np.int64)                       # own ids are one contiguous range
        mesh_sel = (flags[k] & vpb.VP_REBUILD_MESH) != 0
        _, want_s, want_c = checker.rebuild(ids, 0)
        want_m = want_mc = None
        if mesh_sel.any():
            _, want_m, want_mc = checker.rebuild(ids[mesh_sel], 1)
    fnv = helpers.oracle_lib().vo_fnv1a
    bad = []

    def compare(tag, pos, res, splat, mesh):
        ok_c = np.array_equal(res["svl_items"][pos], want_c[:, :5])
        got_s = fnv_chunks(fnv, splat, res["svl_offset"][pos], res["svl_items_total"][pos].astype(np.int64) * 2)
        if not ok_c or not np.array_equal(got_s, want_s):
            bad.append("%s: splat buffers of %d chunks differ" % (tag, int((got_s != want_s).sum()) or 1))
        if want_m is not None:
            pm = pos[mesh_sel]
            ok_mc = np.array_equal(res["vbo_items"][pm], want_mc[:, 5]) and np.array_equal(res["ibo_items"][pm], want_mc[:, 6])
            got_m = fnv_chunks(fnv, mesh, res["vbo_offset"][pm], res["vbo_items"][pm].astype(np.int64) * 2,
                               res["ibo_offset"][pm], res["ibo_items"][pm].astype(np.int64) * 4)
            if not ok_mc or not np.array_equal(got_m, want_m):
                bad.append("%s: mesh buffers of %d chunks differ" % (tag, int((got_m != want_m).sum()) or 1))

    compare("device step", k, *device_out)
    if e2e_out is not None:
        e_ids, e_res, e_splat, e_mesh = e2e_out
        # the e2e call takes the non-null chunks only; a null chunk with a visible neighbour face is not part of it
        pos = np.searchsorted(e_ids, ids)
        pos[pos >= len(e_ids)] = 0
        present = e_ids[pos] == ids
        if present.all():
            compare("e2e", pos, e_res, e_splat, e_mesh)
        else:
            sel = np.nonzero(present)[0]
            sub_want_c, sub_want_s = want_c[sel], want_s[sel]
            got_s = fnv_chunks(fnv, e_splat, e_res["svl_offset"][pos[sel]], e_res["svl_items_total"][pos[sel]].astype(np.int64) * 2)
            if not np.array_equal(e_res["svl_items"][pos[sel]], sub_want_c[:, :5]) or not np.array_equal(got_s, sub_want_s):
                bad.append("e2e: splat buffers differ")
    border = 0
    if rank + 1 < world_size:
        border += per_row                                            # the last row reads the +z plane of the rank above
    if rank > 0:
        border += int(mesh_sel[:per_row].sum())                      # meshed chunks of the first row read the plane of the rank below
    return len(ids), border, not bad, "; ".join(bad), kind


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ---------------------------------------------------------------------------------------------------
def reference_world(bits):
    """The same seeded world inside the compiled reference (oracle/_ref) or, if that library did not
    travel, inside the oracle port.  Returns (world, rebuild(ids, mode, threads) -> seconds, kind)."""
    helpers = helpers_module()
    from voxplat_b200 import worldgen
    w = worldgen.World(SEED, ROOT_BITW, bits)
    if helpers.ref_available():
        rw = helpers.RefWorld(w)
        return w, (lambda ids, mode, nt: rw.rebuild(ids, mode, nt, hashed=False)[0]), "reference"
    ow = helpers.OracleWorld(w)
    return w, (lambda ids, mode, nt: ow.rebuild(ids, mode, nt, hashed=False)[0]), "port"


CPU_NOTE = ("per chunk the reference's own call sequence chunkset.c:318-458 with per-thread preallocated scratch (only the (R+1)^3 mask / work "
            "bytes the functions touch are cleared, not the whole 10x scratch of chunkset.c:321-328); outputs are not hashed inside the clock")


def cpu_rebuild_seconds(bits):
    """One full rebuild with every host thread and one with the 4 threads the reference ships with (chunkset.c:251)."""
    from voxplat_b200 import slab
    with stdout_to_stderr():
        w, rebuild, kind = reference_world(bits)
        ids = np.arange(w.n_chunks, dtype=np.uint32)
        near = ids[slab.near_camera_flags(ids, ROOT_BITW, bits)]
        cores = os.cpu_count() or 1
        t = rebuild(ids, 0, cores) + (rebuild(near, 1, cores) if len(near) else 0.0)
        # "as shipped": num_threads(4); a quarter of the world bounds the run
        q = ids[:len(ids) // 4]
        qn = q[slab.near_camera_flags(q, ROOT_BITW, bits)]
        t4 = rebuild(q, 0, 4) + (rebuild(qn, 1, 4) if len(qn) else 0.0)
    desc = "%d of %d chunks (whole world), splat all + mesh %d near, %d threads, OpenMP dynamic; %s" % (len(ids), w.n_chunks, len(near), cores, CPU_NOTE)
    return {"value": len(ids) * w.N / t / 1e9, "unit": "Gvoxel/s", "cores": cores, "kind": kind, "sample": desc, "seconds": t,
            "as_shipped_4_threads": {"value": len(q) * w.N / t4 / 1e9, "unit": "Gvoxel/s", "cores": 4,
                                     "sample": "first %d chunks (a quarter of the world), num_threads(4) as in chunkset.c:251" % len(q)}}


def run_reference(args):
    """Reference arm: build the world once, then time warmup + steps rebuilds of the bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from voxplat_b200 import slab
    # bounded sample: one GPU's share of the workload = the N = 1 world (CPU throughput in Gvoxel/s does not depend on the
    # world's z extent), so the run ends within minutes and needs ~1 GB of host memory
    sample_bits = world_bits("c2", 1)
    with stdout_to_stderr():
        w, rebuild, kind = reference_world(sample_bits)
        n_sample = w.n_chunks
        ids = np.arange(n_sample, dtype=np.uint32)
        near = ids[slab.near_camera_flags(ids, ROOT_BITW, sample_bits)]
        cores = os.cpu_count() or 1
        times = []
        for i in range(args.warmup + args.steps):
            t = rebuild(ids, 0, cores) + (rebuild(near, 1, cores) if len(near) else 0.0)
            if i >= args.warmup:
                times.append(t)
    sec = float(np.mean(times))
    gv = n_sample * w.N / sec / 1e9
    desc = "%d chunks per step = the 2048x256x2048 world (one GPU's share of the workload; splat all + mesh %d near), %d host threads, OpenMP dynamic schedule; %s" % (
        n_sample, len(near), cores, CPU_NOTE)
    line = {"impl": "reference", "metric": METRIC, "value": gv, "unit": "Gvoxel/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak" if args.workload == "c2" else "strong", "vs_baseline": None, "dtype": "u8", "data": DATA,
            "config": config_dict(args.workload, args.gpus),
            "cpu_baseline": {"value": gv, "unit": "Gvoxel/s", "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": gv, "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(device_index):
    """Multi-rank runs: pin this process to the CPUs NVML reports as local to its GPU, so the pinned staging buffers
    (hundreds of MB per step over PCIe) are allocated on that socket instead of all on node 0.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        cpus = [c for c in cpus if c < os.cpu_count()]
        if cpus and len(cpus) < os.cpu_count():
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


class Comm:
    device = "cuda"              # where the few scalars that cross ranks live (tests/test_bench_dryrun.py drives the same code over gloo)

    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert self.world_size == args.gpus, "launch with torchrun --nproc-per-node == --gpus"
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world_size > 1:
            bind_to_gpu_numa_node(self.local_rank)       # before any pinned allocation: first touch decides the NUMA node
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op="sum"):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.device)
        if self.dist:
            self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN}[op])
        return [float(x) for x in t.tolist()]


def pcie_floor_ms(comm, h2d_bytes, d2h_bytes, reps=3):
    """The copies of one e2e step alone (pinned memory, two streams), on all ranks at once: the floor of the e2e step."""
    torch = comm.torch
    dev_out = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=comm.device)
    host_out = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8).pin_memory()
    dev_in = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, device=comm.device)
    host_in = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s1):
            host_out.copy_(dev_out, non_blocking=True)
        with torch.cuda.stream(s2):
            dev_in.copy_(host_in, non_blocking=True)
    once()
    comm.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    comm.barrier()
    return comm.reduce([(time.perf_counter() - t0) / reps * 1e3], "max")[0]


def run_world(args, comm, workload, steps, warmup, with_cpu=False, with_mesh_all=False):
    """One workload on all ranks: device-resident timed steps, the end-to-end call, the parity guard.  Returns the JSON
    dict on rank 0 (None elsewhere)."""
    import voxplat_b200 as vpb
    from voxplat_b200 import slab
    torch, dist, rank, world_size = comm.torch, comm.dist, comm.rank, comm.world_size

    bits = world_bits(workload, world_size)
    nx, ny, nz = (1 << b for b in bits)
    R, N = 1 << ROOT_BITW, 1 << (3 * ROOT_BITW)
    z0, z1 = slab.slab_rows(nz, world_size, rank)
    per_row = nx * ny
    own_vox = (z1 - z0) * per_row * N
    ctx = vpb.Context(ROOT_BITW, bits, device=comm.local_rank, slab=(z0, z1), splat_arena_bytes=max(2 << 30, own_vox // 2),
                      mesh_arena_bytes=2 << 30, rle_arena_bytes=max(1 << 30, own_vox // 4))
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    if workload == "c2":
        own_ids, dense, solid, shadow_rows, sz0 = generate_slab(workload, bits, z0, z1)
        nn = np.nonzero(solid)[0]
        pinned_dense = torch.from_numpy(np.ascontiguousarray(dense[nn])).pin_memory()
        ctx.upload_chunks_dense(own_ids[nn], pinned_dense)
        ctx.upload_shadow_rows(sz0, shadow_rows)
        del dense, pinned_dense
    else:
        # the big fixed worlds (34 / 137 Gvoxel) are generated on the device: same generator, byte for byte
        # (tests/test_gpu_worldgen.py; the parity guard below checks rows of it against the HOST generator's world)
        own_ids = np.arange(z0 * per_row, z1 * per_row, dtype=np.uint32)
        ctx.generate_world(SEED)
        solid = ctx.chunks_resident(own_ids).astype(np.uint32)
        nn = np.nonzero(solid)[0]
        sz0 = z0 * R
        shadow_rows = ctx.download_shadow_rows(sz0, min(nz * R, z1 * R + 17))

    out = None
    with torch.cuda.stream(stream):
        rebuilder = slab.SlabRebuilder(ctx, rank, world_size, lambda n: torch.empty(n, dtype=torch.uint8, device=comm.device), dist=dist)
        near = slab.near_camera_flags(own_ids, ROOT_BITW, bits)
        flags = np.where(near, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH, vpb.VP_REBUILD_SPLAT).astype(np.uint8)
        ctx.batch_prepare(own_ids, per_chunk_flags=flags)

        def step():
            rebuilder.rebuild_step(mesh=True)        # border exchange overlapped with the chunks that do not read a ghost row

        warm = max(warmup, 3)
        for _ in range(warm):
            step()
        comm.barrier()
        launches0 = ctx.kernel_launches()
        sampler = ClockSampler(comm.local_rank)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        sampler.start()
        comm.barrier()
        evs[0].record(stream)
        for i in range(steps):
            step()
            evs[i + 1].record(stream)
        comm.barrier()
        sampler.pause()
        total_ms = comm.reduce([evs[0].elapsed_time(evs[steps])], "max")[0]
        per_step = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(steps)])
        step_med = comm.reduce([float(np.median(per_step))], "max")[0]
        launches = ctx.kernel_launches() - launches0
        res, splat_bytes_out, mesh_bytes_out = ctx.rebuild_device_results()
        hist_n = min(steps, 256)
        splat_ms, mesh_ms = ctx.kernel_ms_history(hist_n)
        # what the timed steps left in the arenas (for the parity guard; downloaded after all timing)
        dev_splat = ctx.arena_download(0, splat_bytes_out)
        dev_mesh = ctx.arena_download(1, mesh_bytes_out)
        # the splat kernels alone (the timed step runs the mesh kernels beside them on a second stream)
        ctx.batch_prepare(own_ids, flags=vpb.VP_REBUILD_SPLAT)
        iso_n = max(3, min(steps, 20))
        for _ in range(iso_n):
            ctx.rebuild_device()
        comm.barrier()
        splat_iso_ms, _ = ctx.kernel_ms_history(iso_n)
        # ... and the mesh kernels of the near-field chunks alone
        # (only the ranks near the camera have any; EVERY rank takes the barrier -- collectives never sit inside a
        # rank-dependent branch in this file)
        mesh_iso_ms = [0.0]
        if near.any():
            ctx.batch_prepare(own_ids[near], flags=vpb.VP_REBUILD_MESH)
            for _ in range(iso_n):
                ctx.rebuild_device()
        comm.barrier()
        if near.any():
            _, mesh_iso_ms = ctx.kernel_ms_history(iso_n)
        mesh_all = None
        if with_mesh_all:
            # stress figure of SURVEY 8(d): the quad mesh of EVERY chunk (the game meshes only the near field)
            ctx.batch_prepare(own_ids, flags=vpb.VP_REBUILD_MESH)
            for _ in range(3):
                ctx.rebuild_device()
            comm.barrier()
            r_all, _, mb_all = ctx.rebuild_device_results()
            _, m_ms = ctx.kernel_ms_history(3)
            faces_all = int(r_all["vbo_items"].astype(np.int64).sum()) // 16
            bytes_all = int((r_all["vbo_items"] > 0).sum()) * (R + 2) ** 3 + 64 * faces_all
            mesh_all = {"kernel": "k_mesh_count + k_mesh_emit, every chunk meshed", "kernel_ms": float(np.mean(m_ms)), "faces": faces_all,
                        "algorithmic_bytes_per_launch": bytes_all, "achieved": bytes_all / max(float(np.mean(m_ms)), 1e-9) / 1e6, "unit": "GB/s"}
        ctx.batch_prepare(own_ids, per_chunk_flags=flags)

        # ---- end to end through the host-facing C ABI: host RLE streams -> H2D -> decode -> rebuild -> D2H ----
        words, offs = ctx.encode_chunks_rle(own_ids[nn])
        pinned_words = torch.from_numpy(words).pin_memory()
        pinned_shadow = torch.from_numpy(np.ascontiguousarray(shadow_rows)).pin_memory()
        e2e_steps = max(3, min(steps, 10))

        # One pipelined call (upload+decode | kernels | download overlap over 8 blocks of chunk rows).  N > 1: the slab's
        # first and last chunk rows are decoded first so that the border planes can be exchanged before the pipeline
        # starts (the pipeline decodes them again with their blocks: 2 of 32 rows).
        nn_flags = np.ascontiguousarray(flags[nn])
        nn_ids = own_ids[nn]
        border = None
        if world_size > 1:
            rows_of = nn_ids // per_row
            sel = np.nonzero((rows_of == z0) | (rows_of == z1 - 1))[0]
            b_words = np.concatenate([words[int(offs[i]):int(offs[i + 1])] for i in sel]) if len(sel) else np.zeros(0, np.uint32)
            b_offs = np.zeros(len(sel) + 1, np.uint64)
            b_offs[1:] = np.cumsum([int(offs[i + 1] - offs[i]) for i in sel])
            border = (np.ascontiguousarray(nn_ids[sel]), torch.from_numpy(b_words).pin_memory(), b_offs)

        def e2e_step():
            ctx.upload_shadow_rows_async(sz0, pinned_shadow)        # travels in front of the step on the context stream
            if border is not None:
                if len(border[0]):
                    ctx.upload_chunks_rle(*border)                  # rle_decompress of the border rows on the device
                rebuilder.exchange_halos(mesh=True)
            return ctx.rebuild_from_rle(nn_ids, pinned_words, offs, per_chunk_flags=nn_flags, n_blocks=E2E_BLOCKS)

        for _ in range(3):
            r_e, sb_e, mb_e = e2e_step()
        comm.barrier()
        l0 = ctx.kernel_launches()
        sampler.start()                                   # the e2e loop is a timed region too
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r_e, sb_e, mb_e = e2e_step()
        comm.barrier()
        e2e_s = comm.reduce([time.perf_counter() - t0], "max")[0]
        clocks = sampler.stop()
        e2e_launches = (ctx.kernel_launches() - l0) // e2e_steps
        h2d = int(words.nbytes + offs.nbytes + pinned_shadow.numel() * 2 + own_ids[nn].nbytes + own_ids.nbytes + flags.nbytes)
        if border is not None:
            h2d += int(border[1].numel() * 4 + border[2].nbytes + border[0].nbytes)
        d2h = int(sb_e.nbytes + mb_e.nbytes + r_e.nbytes)
        floor_ms = pcie_floor_ms(comm, h2d, d2h)

        # ---- parity guard: after every clock has stopped ----
        try:
            chk_n, chk_border, chk_ok, chk_detail, chk_kind = parity_guard(
                workload, bits, rank, world_size, z0, z1, own_ids, flags, (res, dev_splat, dev_mesh),
                (nn_ids, r_e, np.asarray(sb_e), np.asarray(mb_e)))
        except Exception as exc:          # a checker problem on ONE rank must not leave the others waiting in a collective
            chk_n, chk_border, chk_ok, chk_detail, chk_kind = 0, 0, None, "checker failed: %r" % (exc,), "none"
        if not chk_ok:
            print("bench.py: %s on rank %d (%s): %s" % ("PARITY MISMATCH" if chk_ok is False else "parity unchecked", rank, workload, chk_detail),
                  file=sys.stderr, flush=True)

    # ---- reduce the per-rank figures ----
    sb_a, mb_a, splats, faces = algorithmic_bytes(own_ids, solid, res, bits)
    agg = comm.reduce([sb_a, mb_a, splats, faces, h2d, d2h, len(nn), chk_n, chk_border, 1 if chk_ok is False else 0, 1 if chk_ok is None else 0])
    total_vox = float(nx * ny * nz) * N
    ms_per_step = total_ms / steps
    value = total_vox / (ms_per_step * 1e-3) / 1e9
    e2e_ms = e2e_s / e2e_steps * 1e3
    e2e_value = total_vox / (e2e_ms * 1e-3) / 1e9

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
        k_ms = float(np.mean(splat_ms))
        k_iso = float(np.mean(splat_iso_ms))
        achieved = sb_a / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "splat_traffic.json")))
            if tj.get("workload_bits") == list(bits):
                traffic = tj["dram_bytes_per_launch"]
                traffic_src = "profiles/splat_traffic.json: %s (an ncu --set full capture of the same kernels on this world, not measured in this run)" % tj.get("source", "")
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": value, "unit": "Gvoxel/s",
            "n_gpus": world_size, "steps": steps, "warmup": warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak" if workload == "c2" else "strong", "vs_baseline": None, "dtype": "u8",
            "data": DATA,
            "config": config_dict(workload, world_size),
            "workload_stats": {"non_null_chunks": int(agg[6]), "splats": int(agg[2]), "mesh_faces": int(agg[3])},
            "windows": {"per_step_ms_median_max_over_ranks": step_med, "per_step_ms_min": float(per_step.min()), "per_step_ms_max": float(per_step.max()),
                        "note": "CUDA events around every step on rank 0's stream; value uses the whole K-step region"},
            "roofline": {"bound": "hbm", "kernel": "k_splat_count + k_splat_emit (cull + 5 LOD, then splat emission; timed as one pair), rank 0",
                         "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": sb_a, "kernel_ms": k_ms,
                         "kernel_ms_alone": k_iso, "frac_alone": sb_a / (k_iso * 1e-3) / 1e9 / peak,
                         "step": {"algorithmic_bytes": sb_a + mb_a, "achieved": (sb_a + mb_a) / (ms_per_step * 1e-3) / 1e9,
                                  "frac": (sb_a + mb_a) / (ms_per_step * 1e-3) / 1e9 / peak},
                         "mesh_kernel": {"kernel_ms": float(np.mean(mesh_ms)), "kernel_ms_alone": float(np.mean(mesh_iso_ms)), "algorithmic_bytes_per_launch": mb_a,
                                         "achieved": mb_a / max(float(np.mean(mesh_ms)), 1e-9) / 1e6}},
            "e2e": {"value": e2e_value, "unit": "Gvoxel/s", "h2d_bytes_per_step": int(agg[4]), "d2h_bytes_per_step": int(agg[5]),
                    "ms_per_step": e2e_ms, "steps": e2e_steps, "pcie_floor_ms": floor_ms,
                    "pcie_floor_note": "the same H2D + D2H bytes per rank as plain pinned copies on two streams, all ranks at once, max over ranks",
                    "path": ("host RLE streams (pinned) -> vp_rebuild_from_rle (%d blocks pipelined: H2D + device decode | cull/LOD/splat/mesh | D2H) -> pinned host staging" % E2E_BLOCKS
                             if world_size == 1 else "host RLE streams (pinned) -> border rows: vp_upload_chunks_rle + NCCL plane exchange -> vp_rebuild_from_rle (8 blocks pipelined) -> pinned host staging"),
                    "gpu_launches_per_step": int(e2e_launches)},
            "parity": {"chunks": int(agg[7]), "border_chunks": int(agg[8]), "ok": (agg[9] == 0) if agg[10] == 0 else (False if agg[9] else None),
                       "ranks_unchecked": int(agg[10]), "checker": chk_kind,
                       "what": "per rank: first, last (and one middle) chunk row; splat + mesh buffers of the timed device step and of the e2e call, FNV-1a 64 per chunk vs the checker on the same world"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if mesh_all is not None:
            mesh_all["frac"] = mesh_all["achieved"] / peak
            out["roofline"]["mesh_all"] = mesh_all
        if with_cpu:
            out["cpu_baseline"] = cpu_rebuild_seconds(bits)
    ctx.close()
    del ctx
    torch.cuda.empty_cache()
    return out, agg[9] == 0


# ---------------------------------------------------------------------------------------------------
# the small configs of BASELINE.json (N = 1)
# ---------------------------------------------------------------------------------------------------
def run_c1():
    """Config C1: one 64^3 worldgen chunk at chunk coordinate (1,0,1) of the seed-1234 2048x256x2048 world: RLE decode + cull
    + 5 LOD splat lists + mesh + RLE encode, host RLE in -> host buffers out, next to the reference single-threaded."""
    helpers = helpers_module()
    import voxplat_b200 as vpb
    rb, bits = ROOT_BITW, BASE_BITS
    nx, ny = 1 << bits[0], 1 << bits[1]
    cid = (1 * ny + 0) * nx + 1
    with stdout_to_stderr():
        sw = helpers.SparseWorld(SEED, rb, bits, [0, 1, 2])
    ctx = vpb.Context(rb, bits)
    nn = sw.ids[sw.solid[sw.ids] > 0]
    ctx.upload_chunks_dense(nn, np.ascontiguousarray(np.stack([sw.dense[int(i)] for i in nn])))
    ctx.upload_shadow_rows(0, sw.shadow[:sw.shw * 3 * sw.R])
    ids = np.array([cid], np.uint32)
    words, offs = ctx.encode_chunks_rle(ids)
    lat = []
    for _ in range(60):
        t0 = time.perf_counter()
        ctx.upload_chunks_rle(ids, words, offs)                                      # H2D + rle_decompress on the device
        res, splat, mesh = ctx.rebuild_batch(ids, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)   # cull + LOD + splat + mesh, D2H
        enc, _ = ctx.encode_chunks_rle(ids)                                          # rle_compress on the device, D2H
        lat.append((time.perf_counter() - t0) * 1e3)
    out = {"config": "c1", "workload": "one 64^3 chunk (1,0,1): RLE decode + cull + 5 LOD + mesh + RLE encode, host RLE in -> host buffers out",
           "ms": float(np.median(lat[10:])), "ms_min": float(np.min(lat[10:])), "splat_items": int(res["svl_items_total"][0]),
           "mesh_faces": int(res["ibo_items"][0]) // 6, "rle_words": int(words.size)}
    with stdout_to_stderr():
        kind = "reference" if helpers.ref_available() else "port"
        chk = helpers.RefWorld(sw) if kind == "reference" else helpers.OracleWorld(sw)
        g, it = chk.splat(cid)
        v, x = chk.mesh(cid)
        ok = (np.array_equal(res["svl_items"][0], it) and np.array_equal(splat[int(res["svl_offset"][0]):][:g.size * 2].view(np.int16), g)
              and np.array_equal(mesh[int(res["vbo_offset"][0]):][:v.size * 2].view(np.int16), v)
              and np.array_equal(mesh[int(res["ibo_offset"][0]):][:x.size * 4].view(np.uint32), x)
              and np.array_equal(enc, helpers.rle_encode(sw.dense[cid])))
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            chk.splat(cid)
            chk.mesh(cid)
            if kind == "reference":
                buf = np.zeros(sw.N + 1, np.uint32)
                chk.lib.vr_rle_compress(helpers.vp(sw.dense[cid]), C.c_uint32(sw.N), helpers.vp(buf), C.c_uint32(buf.size))
            else:
                helpers.rle_encode(sw.dense[cid])
            ts.append((time.perf_counter() - t0) * 1e3)
    out["parity"] = {"ok": bool(ok), "checker": kind}
    out["cpu_ms_single_thread"] = float(np.median(ts))
    ctx.close()
    return out, bool(ok)


def run_c5(rb, bits, bursts=60):
    """Config C5: bursts of chunkset_edit_sphere (radius 4, alternating place 63 / remove 0, seeded hit points on the surface):
    ms from the edit to all dirty chunks' splat + mesh buffers on the host, the edit applied on the device (vp_edit_sphere).
    Checker: the same bursts through the compiled reference (its own edit code), every dirty chunk compared after the last burst."""
    helpers = helpers_module()
    import voxplat_b200 as vpb
    from voxplat_b200 import worldgen
    with stdout_to_stderr():
        w = worldgen.World(SEED, rb, bits)
    ctx = vpb.Context(rb, bits, mesh_arena_bytes=1 << 30, splat_arena_bytes=1 << 30)
    nn = w.nonnull_ids()
    ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
    ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])
    rng = np.random.default_rng(5)
    X, Y, Z = w.dims
    par = worldgen.params(SEED, rb, bits)
    pts = []
    for b in range(bursts):
        x, z = int(rng.integers(8, X - 8)), int(rng.integers(8, Z - 8))
        pts.append((x, int(worldgen.lib().vpw_height(C.byref(par), x, z)), z, 63 if b % 2 == 0 else 0))
    lat, nd, touched = [], 0, set()
    for (x, h, z, v) in pts:
        t0 = time.perf_counter()
        dirty = ctx.edit_sphere(x, h, z, 4, v)
        res, splat, mesh = ctx.rebuild_batch(dirty, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
        lat.append((time.perf_counter() - t0) * 1e3)
        nd += len(dirty)
        touched.update(int(i) for i in dirty)
    lat = np.array(lat[5:])
    out = {"config": "c5", "workload": "edit bursts: chunkset_edit_sphere r=4 (place 63 / remove 0) -> dirty chunks re-culled + re-meshed, buffers on the host; chunk %d^3, world %dx%dx%d" % (
               1 << rb, X, Y, Z),
           "bursts": bursts, "ms_per_burst": float(np.median(lat)), "ms_per_burst_p95": float(np.percentile(lat, 95)),
           "dirty_chunks_per_burst": nd / bursts, "gvoxel_per_s_dirty_set": nd / bursts * w.N / (float(np.median(lat)) * 1e-3) / 1e9}
    ok = True
    helpers_ok = helpers.ref_available()
    if helpers_ok:
        with stdout_to_stderr():
            ref = helpers.RefWorld(w)
            ts = []
            for (x, h, z, v) in pts:
                t0 = time.perf_counter()
                ref.edit_sphere(x, h, z, 4, v)
                ts.append(time.perf_counter() - t0)
            ids = np.array(sorted(touched), np.uint32)
            t_s, want_s, want_c = ref.rebuild(ids, 0, 4)
            t_m, want_m, want_mc = ref.rebuild(ids, 1, 4)
            res, splat, mesh = ctx.rebuild_batch(ids, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
            fnv = helpers.oracle_lib().vo_fnv1a
            got_s = fnv_chunks(fnv, splat, res["svl_offset"], res["svl_items_total"].astype(np.int64) * 2)
            got_m = fnv_chunks(fnv, mesh, res["vbo_offset"], res["vbo_items"].astype(np.int64) * 2, res["ibo_offset"], res["ibo_items"].astype(np.int64) * 4)
            ok = (np.array_equal(res["svl_items"], want_c[:, :5]) and np.array_equal(got_s[res["svl_items_total"] > 0], want_s[res["svl_items_total"] > 0])
                  and np.array_equal(res["vbo_items"], want_mc[:, 5]) and np.array_equal(got_m[res["vbo_items"] > 0], want_m[res["vbo_items"] > 0]))
        out["cpu_ms_per_burst"] = float(np.median(ts)) * 1e3 + (t_s + t_m) / max(len(ids), 1) * (nd / bursts) * 1e3
        out["cpu_note"] = "reference: chunkset_edit_sphere + its splat and mesh rebuild of the burst's dirty chunks on 4 threads (chunkset.c:251), averaged over the dirty set"
        out["parity"] = {"ok": bool(ok), "checker": "reference", "chunks": int(len(ids)), "what": "every chunk dirtied by the %d bursts, splat + mesh, after the last burst" % bursts}
    ctx.close()
    return out, bool(ok)


def run_native(args, comm=None):
    comm = comm or Comm(args)
    n = comm.world_size
    ok_all = True
    line, ok = run_world(args, comm, args.workload, args.steps, args.warmup, with_cpu=(n == 1 and not args.no_cpu_baseline),
                         with_mesh_all=(n == 1 and args.workload == "c2"))
    ok_all &= ok
    extras = []
    if not args.no_extra and args.workload == "c2":
        if n == 1:
            for name, fn in (("c1", run_c1), ("c5 chunk 32", lambda: run_c5(5, (4, 2, 4))), ("c5 chunk 128", lambda: run_c5(7, (3, 1, 3)))):
                try:
                    r, ok = fn()
                except Exception as exc:
                    print("bench.py: extra workload %s failed: %r" % (name, exc), file=sys.stderr, flush=True)
                    r, ok = {"config": name, "error": repr(exc)}, True
                extras.append(r)
                ok_all &= ok
        else:
            todo = ["c3"] + (["c4"] if n == 8 else [])
            for wl in todo:
                try:
                    r, ok = run_world(args, comm, wl, max(3, min(args.steps, 10)), 3)
                except Exception as exc:      # (the same exception on every rank: sizes are symmetric) keep the headline line
                    print("bench.py: extra workload %s failed on rank %d: %r" % (wl, comm.rank, exc), file=sys.stderr, flush=True)
                    extras.append({"config": config_dict(wl, n), "error": repr(exc)})
                    break
                ok_all &= ok
                if r is not None:
                    extras.append({k: r[k] for k in ("config", "value", "unit", "ms_per_step", "steps", "scaling", "workload_stats", "roofline", "e2e", "parity", "gpu_launches")})
    if comm.rank == 0:
        if extras:
            line["extra"] = extras
        print(json.dumps(line))
    if comm.dist:
        comm.dist.destroy_process_group()
    if not ok_all:
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE.json configs (the `extra` array)")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="BASELINE.json config (default c2, the metric's config)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
