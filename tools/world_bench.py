#!/usr/bin/env python
"""BASELINE.json configs[4]: a large world partitioned across the GPUs of one box (X-slabs), generated on the device,
meshed in batches through a reused vertex arena, in the reference's chunk-local mode and in halo mode with the
region-seam exchange over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/world_bench.py --dims 512 32 512 [--batch 32768] [--verify]

Prints one JSON line per mode from rank 0.  Strong scaling (the world is fixed, ranks split it).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = 0xD5E50001
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[512, 32, 512], help="world size in chunks: x y z (y is up)")
    ap.add_argument("--batch", type=int, default=98304, help="chunks per mesh launch (arena is reused between launches)")
    ap.add_argument("--small-output-batch", type=int, default=786432, help="chunks per launch in the halo / greedy modes (their output is 10-20x smaller)")
    ap.add_argument("--arena-gib", type=float, default=8.0)
    ap.add_argument("--quad-modes", action="store_true", help="also run the greedy + indexed modes (chunk-local and halo culling)")
    ap.add_argument("--verify", action="store_true", help="check a sample of seam chunks (and some interior ones) against the oracle, at any world size")
    ap.add_argument("--seam", choices=["c", "nccl", "torch"], default="c", help="c: dsp_seam_* inside the library over peer memory (stores over NVLink, wait inside the "
                    "mesh kernel); nccl: the library's own NCCL fallback transport (ncclSend/ncclRecv from C); torch: pack -> torch.distributed isend/irecv -> "
                    "dsp_set_halo_slabs (round 1's route)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from despawnengine_b200 import capi
    from despawnengine_b200 import partition as part
    from despawnengine_b200.halo import RegionOnDevice, connect_region_seams, connect_region_seams_nccl, exchange_region_halos

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    W = part.Region((0, 0, 0), tuple(args.dims))
    sp = part.SlabPartition(W, world)
    reg = sp.region(rank)
    n = reg.n_chunks
    arena = int(args.arena_gib * (1 << 30)) // 128 * 128
    ctx = capi.Context(local, n, arena)
    ctx.set_uv_table(UV)

    barrier()
    t0 = time.perf_counter()
    dev = RegionOnDevice(ctx, reg, capi.GEN_NOISE, SEED)
    ctx.sync()
    barrier()
    gen_s = reduce(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)

    seam_setup_s = 0.0
    if world > 1 and args.seam in ("c", "nccl"):
        t0 = time.perf_counter()
        if args.seam == "c":
            connect_region_seams(dev, sp, rank, dist)  # once: inboxes, handle swap, peer mapping
        else:
            connect_region_seams_nccl(dev, sp, rank, dist)  # once: communicator, inboxes + send buffers
        seam_setup_s = time.perf_counter() - t0

    out = []
    modes = [("parity", capi.MESH_PARITY), ("halo", capi.MESH_HALO)]
    if args.quad_modes:
        modes += [("greedy_indexed", capi.MESH_GREEDY | capi.MESH_INDEXED), ("halo_greedy_indexed", capi.MESH_HALO | capi.MESH_GREEDY | capi.MESH_INDEXED)]
    for mode_name, mode in modes:
        xchg_s, xchg_bytes = 0.0, 0
        c_seams = bool(mode & capi.MESH_HALO) and world > 1 and args.seam in ("c", "nccl")
        if (mode & capi.MESH_HALO) and world > 1 and args.seam == "torch":
            ctx.clear_halo_slabs()
            exchange_region_halos(dev, sp, rank, dist)  # first exchange pays NCCL's lazy point-to-point setup; time the second
            barrier()
            t0 = time.perf_counter()
            xchg_bytes = exchange_region_halos(dev, sp, rank, dist)
            ctx.sync()
            barrier()
            xchg_s = reduce(time.perf_counter() - t0, dist.ReduceOp.MAX)
        if c_seams:
            # an exchange on its own, for the record (enqueue -> all ranks' pushes complete); in the timed pass below it is
            # enqueued in front of the mesh launches and nobody waits for it on the host
            ctx.seam_exchange_async(); ctx.mesh_range(dev.first_slot, min(n, 1024), mode)  # epoch 1 (warm-up)
            barrier()
            t0 = time.perf_counter()
            ctx.seam_exchange_async()
            ctx.sync()
            barrier()
            xchg_s = reduce(time.perf_counter() - t0, dist.ReduceOp.MAX)
            ctx.mesh_range(dev.first_slot, min(n, 1024), mode)
            xchg_bytes = sp.seam_bytes(rank)
        batch = args.batch if mode == capi.MESH_PARITY else args.small_output_batch
        # warm-up batch (also builds the neighbour table in halo mode)
        ctx.mesh_range(dev.first_slot, min(n, args.batch), mode)
        faces = 0
        vbytes = 0
        barrier()
        t0 = time.perf_counter()
        if c_seams:
            ctx.seam_exchange_async()  # this epoch's border layers: on their way while the interior is meshed
        for b0 in range(0, n, batch):
            m = min(batch, n - b0)
            ctx.mesh_range_async(dev.first_slot + b0, m, mode)
            _, total = ctx.mesh_result(m, entries_ptr=0)  # syncs; raises on arena overflow
            vbytes += total
        barrier()
        mesh_s = reduce(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
        kn, kms, _ = ctx.mesh_kernel_stats()
        vbytes_all = reduce(float(vbytes), dist.ReduceOp.SUM if world > 1 else None)
        line = {"workload": f"configs[4]: {args.dims[0]}x{args.dims[2]}x{args.dims[1]}-chunk world (x,z,y-up), X-slab partition", "mode": mode_name,
                "n_gpus": world, "chunks": W.n_chunks, "chunks_per_gpu": n, "batch": batch,
                "terrain_fill_s": gen_s, "terrain_fill_GBps_per_gpu": n * 8192 / gen_s / 1e9,
                "halo_exchange_s": xchg_s, "halo_bytes_sent_rank0": xchg_bytes, "seam_path": (args.seam if (mode & capi.MESH_HALO) and world > 1 else None),
                "halo_exchange_inside_mesh_wall": bool(c_seams), "seam_setup_s": seam_setup_s,
                "mesh_wall_s": mesh_s, "chunks_per_s": W.n_chunks / mesh_s, "voxels_per_s": W.n_chunks * 4096 / mesh_s,
                "vertex_bytes": vbytes_all, "hbm_algorithmic_GBps_per_gpu": (vbytes_all / world + n * 8224) / mesh_s / 1e9,
                "scaling": "strong", "timing": "wall clock between barriers, max over ranks; one sync per batch"}
        out.append(line)
        if rank == 0:
            print(json.dumps(line), flush=True)

    if args.verify:
        from oracle import loader as orc
        # seam chunks of this rank in halo mode against the extended oracle, neighbours generated on the CPU
        DIRS = [(0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (-1, 0, 0), (1, 0, 0)]
        sample = []
        for t in sp.seam_transfers(rank):
            sample += [tuple(p) for p in t.positions[:: max(1, len(t.positions) // 6)].tolist()]
        slots = dev.slots_of(np.array(sample, dtype=np.int32)) if sample else np.array([], dtype=np.uint32)
        entries, _ = ctx.mesh_chunks(slots, capi.MESH_HALO)
        for p, e in zip(sample, entries):
            ids = orc.generate(orc.GEN_NOISE, SEED, p)
            slabs, present = np.zeros((6, 256), dtype=np.uint16), np.zeros(6, dtype=np.uint8)
            for f, d in enumerate(DIRS):
                q = (p[0] + d[0], p[1] + d[1], p[2] + d[2])
                if all(W.origin[i] <= q[i] < W.origin[i] + W.dims[i] for i in range(3)):
                    v = orc.generate(orc.GEN_NOISE, SEED, q).reshape(16, 16, 16)
                    slabs[f] = [v[:, 0, :], v[:, 15, :], v[15, :, :], v[0, :, :], v[:, :, 15], v[:, :, 0]][f].reshape(256)
                    present[f] = 1
            want = orc.mesh_chunk_halo(p, ids, slabs, present, UV)
            assert ctx.download_vertices(e).tobytes() == want.tobytes(), (rank, p)
        ok = reduce(1.0, dist.ReduceOp.SUM if world > 1 else None)
        if rank == 0:
            print(json.dumps({"verify": "seam chunks byte-identical to the extended oracle", "ranks_ok": ok, "sampled_per_rank": len(sample)}), flush=True)

    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
