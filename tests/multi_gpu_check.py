"""Run under torchrun (one rank per GPU): slab-sharded rebuild with NCCL border exchange, every rank
checks its own chunks byte for byte against the oracle evaluated on the WHOLE world.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import voxplat_b200 as vpb  # noqa: E402
from voxplat_b200 import slab, worldgen  # noqa: E402


def main():
    rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    failures = 0
    for rb, bits, kind in [(4, (1, 1, 3), "random"), (5, (2, 1, 2), "terrain"), (6, (1, 0, 2), "terrain")]:
        w = worldgen.World(77, rb, bits) if kind == "terrain" else helpers.random_world(77, rb, bits, density=0.4, null_frac=0.2)
        o = helpers.OracleWorld(w)
        nz = 1 << bits[2]
        z0, z1 = slab.slab_rows(nz, ws, rank)
        per_row = 1 << (bits[0] + bits[1])
        own = np.arange(z0 * per_row, z1 * per_row, dtype=np.uint32)
        ctx = vpb.Context(rb, bits, device=lr, slab=(z0, z1), mesh_arena_bytes=max(64 << 20, len(own) * w.N * 90),
                          splat_arena_bytes=max(64 << 20, len(own) * (w.R + 1) ** 3 * 10))
        nn = own[w.solid[own] > 0]
        if len(nn):
            ctx.upload_chunks_dense(nn, np.ascontiguousarray(w.dense[nn]))
        ctx.upload_shadow_rows(0, w.shadow[:w.shw * w.dims[2]])       # rows outside the slab's reach are ignored
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            rbld = slab.SlabRebuilder(ctx, rank, ws, lambda n: torch.empty(n, dtype=torch.uint8, device="cuda"), dist=dist)
            rbld.exchange_halos(mesh=True)
            res, splat, mesh = ctx.rebuild_batch(own, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
            splat, mesh = splat.copy(), mesh.copy()                 # the staging is reused by the later calls
            # the device-resident step with the exchange hidden behind the chunks that do not read a ghost row must
            # give the same buffers (arena offsets may differ: compare per chunk)
            ctx.batch_prepare(own, vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH)
            rbld.rebuild_step(mesh=True)
            res2, sb2, mb2 = ctx.rebuild_device_results()
            splat2, mesh2 = ctx.arena_download(0, sb2), ctx.arena_download(1, mb2)
            # end-to-end form of a slab: border rows decoded first, planes exchanged, then the pipelined call
            if len(nn):
                words, offs = ctx.encode_chunks_rle(nn)
                ctx.set_chunks_null(own)                                # forget the world: everything comes back from RLE
                rows_of = nn // per_row
                sel = np.nonzero((rows_of == z0) | (rows_of == z1 - 1))[0]
                if len(sel):
                    bw = np.concatenate([words[int(offs[i]):int(offs[i + 1])] for i in sel])
                    bo = np.zeros(len(sel) + 1, np.uint64)
                    bo[1:] = np.cumsum([int(offs[i + 1] - offs[i]) for i in sel])
                    ctx.upload_chunks_rle(np.ascontiguousarray(nn[sel]), bw, bo)
            rbld.exchange_halos(mesh=True)
            if len(nn):
                nn_pos = np.searchsorted(own, nn)
                res3, splat3, mesh3 = ctx.rebuild_from_rle(nn, words, offs, flags=vpb.VP_REBUILD_SPLAT | vpb.VP_REBUILD_MESH, n_blocks=4)
                for j, k in enumerate(nn_pos):
                    ok3 = np.array_equal(res3["svl_items"][j], res["svl_items"][k]) and res3["vbo_items"][j] == res["vbo_items"][k]
                    if ok3:
                        a, b, nb = int(res["svl_offset"][k]), int(res3["svl_offset"][j]), int(res["svl_items_total"][k]) * 2
                        ok3 = np.array_equal(splat[a:a + nb], splat3[b:b + nb])
                        a, b, nb = int(res["vbo_offset"][k]), int(res3["vbo_offset"][j]), int(res["vbo_items"][k]) * 2
                        ok3 = ok3 and np.array_equal(mesh[a:a + nb], mesh3[b:b + nb])
                        a, b, nb = int(res["ibo_offset"][k]), int(res3["ibo_offset"][j]), int(res["ibo_items"][k]) * 4
                        ok3 = ok3 and np.array_equal(mesh[a:a + nb], mesh3[b:b + nb])
                    if not ok3:
                        failures += 1
                        print("rank %d: chunk %d of world rb=%d %s: pipelined slab e2e differs" % (rank, nn[j], rb, bits), flush=True)
        for k in range(len(own)):
            same = np.array_equal(res2["svl_items"][k], res["svl_items"][k]) and res2["vbo_items"][k] == res["vbo_items"][k] and res2["ibo_items"][k] == res["ibo_items"][k]
            if same:
                a, b, nb = int(res["svl_offset"][k]), int(res2["svl_offset"][k]), int(res["svl_items_total"][k]) * 2
                same = np.array_equal(splat[a:a + nb], splat2[b:b + nb])
                a, b, nb = int(res["vbo_offset"][k]), int(res2["vbo_offset"][k]), int(res["vbo_items"][k]) * 2
                same = same and np.array_equal(mesh[a:a + nb], mesh2[b:b + nb])
                a, b, nb = int(res["ibo_offset"][k]), int(res2["ibo_offset"][k]), int(res["ibo_items"][k]) * 4
                same = same and np.array_equal(mesh[a:a + nb], mesh2[b:b + nb])
            if not same:
                failures += 1
                print("rank %d: chunk %d of world rb=%d %s: overlapped step differs from rebuild_batch" % (rank, own[k], rb, bits), flush=True)
        for k, cid in enumerate(own):
            g, it = o.splat(int(cid))
            off = int(res["svl_offset"][k])
            v, x = o.mesh(int(cid))
            vo, io = int(res["vbo_offset"][k]), int(res["ibo_offset"][k])
            ok = (np.array_equal(res["svl_items"][k], it) and np.array_equal(splat[off:off + g.size * 2].view(np.int16), g)
                  and res["vbo_items"][k] == v.size and np.array_equal(mesh[vo:vo + v.size * 2].view(np.int16), v)
                  and np.array_equal(mesh[io:io + x.size * 4].view(np.uint32), x))
            if not ok:
                failures += 1
                print("rank %d: chunk %d of world rb=%d %s differs" % (rank, cid, rb, bits), flush=True)
        ctx.close()
    t = torch.tensor([failures], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("MULTI_GPU_CHECK %s (%d ranks, %d failing chunks)" % ("OK" if t.item() == 0 else "FAILED", ws, int(t.item())), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
