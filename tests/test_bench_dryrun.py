"""bench.py's multi-rank host logic, end to end, on CPU: world_size 2 and 4 gloo groups run bench.run_native unchanged --
slab split, border-plane exchange through voxplat_b200.slab, every barrier and reduction, the parity guard against the
compiled reference, the extra C3 workload, the JSON line -- with tests/hoststore.HostContext (the oracle restatement
behind the Context method surface) standing in for the GPUs and a tiny world standing in for the 2048 x 256 x 2048 one.
A collective that only some ranks reach, a wrong plane, or a wrong border-chunk count fails here, without a GPU."""
import json
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, outdir):
    import contextlib
    import time
    import types
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank),
                      OMP_NUM_THREADS="1")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.stdout = open(os.path.join(outdir, "rank%d.out" % rank), "w")
    import torch
    import torch.distributed as dist
    import bench
    import hoststore
    import voxplat_b200

    # ---- the GPU-shaped pieces bench.py touches, as host objects ----
    class Stream:
        cuda_stream = 0

    class Event:
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3 + 1e-3

    torch.cuda.Stream, torch.cuda.Event = Stream, Event
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.empty_cache = lambda: None
    torch.Tensor.pin_memory = lambda self: self
    voxplat_b200.Context = hoststore.HostContext

    # ---- a world small enough for the oracle: 16^3 chunks, 4 x 2 x 4 of them per rank; "c3" = 4 x 2 x 8 split over the ranks ----
    bench.ROOT_BITW, bench.BASE_BITS = 4, (2, 1, 2)
    bench.WORKLOADS = {"c2": None, "c3": (2, 1, 4), "c4": (1, 1, 4)}          # 16 chunk rows: at least two per rank at world size 8

    comm = bench.Comm.__new__(bench.Comm)
    comm.torch, comm.rank, comm.world_size, comm.local_rank, comm.device = torch, rank, world_size, rank, "cpu"
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    comm.dist = dist
    args = types.SimpleNamespace(gpus=world_size, steps=3, warmup=1, impl="native", no_cpu_baseline=True, no_extra=False, workload="c2")
    bench.run_native(args, comm)
    sys.stdout.flush()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world_size", [2, 4, 8])
def test_bench_multi_rank_flow_on_cpu(world_size, tmp_path):
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world_size, port, str(tmp_path))) for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0, "a rank failed or hung (exit code %r)" % p.exitcode
    lines = [ln for ln in open(os.path.join(str(tmp_path), "rank0.out")).read().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "rank 0 prints exactly one JSON line"
    line = json.loads(lines[0])
    assert line["n_gpus"] == world_size and line["scaling"] == "weak" and line["unit"] == "Gvoxel/s"
    per_row = 8
    # every rank checked its first, last and a middle chunk row; all but the last rank's last row read a plane of the rank above,
    # and (every chunk of this tiny world is meshed) all but the first rank's first row read one of the rank below
    assert line["parity"]["ok"] is True and line["parity"]["ranks_unchecked"] == 0
    assert line["parity"]["chunks"] == 3 * per_row * world_size
    assert line["parity"]["border_chunks"] == 2 * per_row * (world_size - 1)
    for key in ("roofline", "e2e", "clocks", "windows", "gpu_launches", "config", "workload_stats"):
        assert key in line
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    # the other ranks printed nothing
    for r in range(1, world_size):
        assert not [ln for ln in open(os.path.join(str(tmp_path), "rank%d.out" % r)).read().splitlines() if ln.startswith("{")]
    # the extra workloads (fixed worlds split over the ranks: C3, and C4 at world size 8) went through the same machinery
    extra = line["extra"]
    assert [e["config"]["baseline_config"] for e in extra] == (["c3", "c4"] if world_size == 8 else ["c3"])
    for e, rows in zip(extra, (8, 4)):
        assert e["scaling"] == "strong" and e["parity"]["ok"] is True and e["parity"]["ranks_unchecked"] == 0
        assert e["parity"]["border_chunks"] == 2 * rows * (world_size - 1)
        assert e["value"] > 0 and e["e2e"]["value"] > 0


def _worker1(outdir):
    _worker(0, 1, _free_port(), outdir)


def test_bench_single_rank_flow_on_cpu(tmp_path):
    """N = 1: the main line with the mesh-for-all figure, and the extra configs C1 (one chunk) and C5 (edit bursts at two
    chunk sizes), each with its own parity check against the compiled reference."""
    ctx = mp.get_context("spawn")
    p = ctx.Process(target=_worker1, args=(str(tmp_path),))
    p.start()
    p.join(900)
    assert p.exitcode == 0
    lines = [ln for ln in open(os.path.join(str(tmp_path), "rank0.out")).read().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["n_gpus"] == 1 and line["parity"]["ok"] is True and line["parity"]["border_chunks"] == 0
    assert "mesh_all" in line["roofline"] and line["roofline"]["mesh_all"]["faces"] > 0
    extra = line["extra"]
    assert [e["config"] for e in extra] == ["c1", "c5", "c5"]
    for e in extra:
        assert "error" not in e and e["parity"]["ok"] is True


def test_reference_arm_config_equals_native_arm_config():
    sys.path.insert(0, ROOT)
    import bench
    for n in (1, 2, 8):
        assert bench.config_dict("c2", n) == bench.config_dict("c2", n)
        assert set(bench.config_dict("c2", n)) == {"workload", "baseline_config", "chunks", "parallelism", "l2"}
