"""N-rank probe (torchrun): where does a slab step spend its time -- CPU enqueue vs device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import voxplat_b200 as vpb
from voxplat_b200 import slab
import bench
rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
bench.WORKLOAD = "c2"
bits = bench.world_bits(ws)
nz = 1 << bits[2]
z0, z1 = slab.slab_rows(nz, ws, rank)
own_ids, dense, solid, shadow_rows, sz0 = bench.generate_slab(bits, z0, z1)
nn = np.nonzero(solid)[0]
ctx = vpb.Context(bench.ROOT_BITW, bits, device=lr, slab=(z0, z1), splat_arena_bytes=2 << 30, mesh_arena_bytes=2 << 30)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
ctx.upload_chunks_dense(own_ids[nn], np.ascontiguousarray(dense[nn])); ctx.upload_shadow_rows(sz0, shadow_rows)
with torch.cuda.stream(stream):
    rb = slab.SlabRebuilder(ctx, rank, ws, lambda n: torch.empty(n, dtype=torch.uint8, device="cuda"), dist=dist)
    near = slab.near_camera_flags(own_ids, bench.ROOT_BITW, bits)
    flags = np.where(near, 3, 1).astype(np.uint8)
    ctx.batch_prepare(own_ids, per_chunk_flags=flags)
    def run(fn, n=100):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(stream)
        for _ in range(n): fn()
        e1.record(stream); t_cpu = time.perf_counter() - t0
        torch.cuda.synchronize()
        return t_cpu / n * 1e3, e0.elapsed_time(e1) / n
    def seq():
        rb.exchange_halos(mesh=True); ctx.rebuild_device()
    def only_rebuild():
        ctx.rebuild_device()
    def only_exchange():
        rb.exchange_halos(mesh=True)
    for name, fn in [("rebuild only", only_rebuild), ("exchange only", only_exchange), ("exchange + rebuild", seq), ("overlapped step", rb.rebuild_step)]:
        c, g = run(fn)
        if rank == 0: print("%-20s cpu enqueue %.3f ms/step   device %.3f ms/step" % (name, c, g), flush=True)
ctx.close(); dist.destroy_process_group()
