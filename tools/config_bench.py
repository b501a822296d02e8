#!/usr/bin/env python
"""Side benches for the BASELINE.json configs that are not the bench.py headline (one GPU):
  config 1  single default chunk (reference terrain), latency of one build_chunk_mesh call through the C ABI
  config 3  worst-case density: 3-D checkerboard (12,288 faces/chunk) and Bernoulli(0.5), 4,096 chunks
  config 4  incremental remesh: 10,000 random place/destroy edits per frame over a 64x64x16-chunk world
Prints one JSON line per config.  python tools/config_bench.py [--frames 20]
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi

UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
PEAK = 6453.1


def kernel_ms(ctx, fn, reps):
    fn(); ctx.sync()
    n0, t0, _ = ctx.mesh_kernel_stats()
    for _ in range(reps):
        fn()
    ctx.sync()
    n1, t1, _ = ctx.mesh_kernel_stats()
    return (t1 - t0) / (n1 - n0)


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x, z ^ (z >> 31)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=20)
    args = ap.parse_args()
    # ---- config 1 -----------------------------------------------------------------------------------------
    with capi.Context(0, 64, 64 << 20) as ctx:
        ctx.set_uv_table(UV)
        slots = ctx.generate_chunks([(0, 0, 0)], capi.GEN_REFERENCE)
        for _ in range(20):
            ctx.mesh_chunks(slots)
        t0 = time.perf_counter()
        for _ in range(200):
            e, _ = ctx.mesh_chunks(slots)
        call_us = (time.perf_counter() - t0) / 200 * 1e6
        kms = kernel_ms(ctx, lambda: ctx.mesh_chunks_async(slots), 50)
        print(json.dumps({"config": 1, "workload": "single default chunk (0,0,0), reference terrain", "faces": int(e[0]["vertex_count"]) // 6,
                          "blocking_call_us": call_us, "kernel_us": kms * 1e3,
                          "note": "one chunk cannot fill 148 SMs: this is launch + one CTA's latency, the batched call is the design point"}), flush=True)
    # ---- the reference's own first frame: settings.json5 distances 10 / 4, camera (2.5, 22.5, -2.5) ---------
    with capi.Context(0, 4096, 512 << 20) as ctx:
        ctx.set_uv_table(UV)
        cam = (2.5, 22.5, -2.5)
        ts = []
        for _ in range(7):
            ctx.scene_reset(); ctx.sync()
            t0 = time.perf_counter()
            st = ctx.scene_update(cam, 10, 4)
            ts.append(time.perf_counter() - t0)
        vis = capi.visible_chunks(cam, 10, 4)
        from oracle import loader as orc  # CPU baseline leg: the restatement of the same frame, one thread like the reference
        cpu1, verts = orc.time_pipeline(vis, orc.GEN_REFERENCE, 0, UV, orc.FAITHFUL, 1, include_generation=True)
        cpun, _ = orc.time_pipeline(vis, orc.GEN_REFERENCE, 0, UV, orc.FAITHFUL, orc.hardware_threads(), include_generation=True)
        print(json.dumps({"config": "reference first frame", "workload": "GameScene::update at the start camera, render distance 10/4: 2,853 chunks generated, 1,268 meshes",
                          "chunks": int(st["n_loaded_total"]), "meshes": int(st["n_meshes"]), "vertices": int(st["vertex_total"]),
                          "scene_update_ms_median": float(np.median(ts)) * 1e3, "scene_update_ms_best": float(np.min(ts)) * 1e3,
                          "cpu_port_ms_1_thread": cpu1 * 1e3, "cpu_port_ms_all_threads": cpun * 1e3, "cpu_threads": orc.hardware_threads(), "cpu_vertices": verts,
                          "note": "wall clock of the blocking dsp_scene_update call (host visible-set + slot bookkeeping + 1 terrain launch + 1 mesh launch + entry table D2H)"}), flush=True)
    # ---- config 3 -----------------------------------------------------------------------------------------
    for name, kind, seed in (("3a checkerboard (12,288 faces/chunk, the maximum)", capi.GEN_CHECKER, 0), ("3b Bernoulli(0.5)", capi.GEN_RANDOM, 0xD5E50003)):
        with capi.Context(0, 4096, 7 << 30) as ctx:
            ctx.set_uv_table(UV)
            first = ctx.generate_region((0, 0, 0), (16, 16, 16), kind, seed)
            entries, total = ctx.mesh_range(first, 4096)
            faces = int(entries["vertex_count"].astype(np.int64).sum()) // 6
            algo = 4096 * (8192 + 32) + 120 * faces
            kms = kernel_ms(ctx, lambda: ctx.mesh_range_async(first, 4096), 10)
            print(json.dumps({"config": 3, "workload": name + ", 4,096 chunks", "faces": faces, "vertex_bytes": total, "kernel_ms": kms,
                              "chunks_per_s": 4096 / (kms * 1e-3), "achieved_GBps": algo / (kms * 1e-3) / 1e9, "frac_of_measured_peak": algo / (kms * 1e-3) / 1e9 / PEAK}), flush=True)
    # ---- config 2 in the quad modes (extensions: greedy merge, indexed layout) ------------------------------------------
    with capi.Context(0, 4096, 1 << 30) as ctx:
        ctx.set_uv_table(UV)
        first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, 0xD5E50001)
        for name, mode in (("unit quads, 6 vertices (reference layout)", capi.MESH_PARITY), ("unit quads, indexed", capi.MESH_INDEXED),
                           ("greedy merge, 6 vertices", capi.MESH_GREEDY), ("greedy merge, indexed", capi.MESH_GREEDY | capi.MESH_INDEXED)):
            entries, total = ctx.mesh_range(first, 4096, mode)
            indexed = bool(mode & capi.MESH_INDEXED)
            quads = int(entries["vertex_count"].astype(np.int64).sum()) // (4 if indexed else 6)
            algo = 4096 * (8192 + 32) + (104 if indexed else 120) * quads
            kms = kernel_ms(ctx, lambda: ctx.mesh_range_async(first, 4096, mode), 20)
            print(json.dumps({"config": 2, "workload": "16x16x16 chunks of noise terrain, " + name, "mode": mode, "quads": quads, "arena_bytes": total,
                              "kernel_ms": kms, "chunks_per_s": 4096 / (kms * 1e-3), "algorithmic_bytes": algo, "achieved_GBps": algo / (kms * 1e-3) / 1e9,
                              "frac_of_measured_peak": algo / (kms * 1e-3) / 1e9 / PEAK}), flush=True)
    # ---- config 4 -----------------------------------------------------------------------------------------
    dims = (64, 16, 64)
    n = dims[0] * dims[1] * dims[2]
    with capi.Context(0, n, 4 << 30) as ctx:
        ctx.set_uv_table(UV)
        first = ctx.generate_region((0, 0, 0), dims, capi.GEN_NOISE, 0xD5E50001)
        ctx.mesh_range(first, 1024)
        state = 0xD5E50004
        frame_ms, dirty_counts, apply_ms, mesh_ms, fused_ms = [], [], [], [], []
        for frame in range(2 * args.frames + 2):
            edits = np.zeros(10000, dtype=capi.EDIT_DTYPE)
            r = np.empty(40000, dtype=np.uint64)
            for i in range(40000):
                state, r[i] = splitmix64(state)
            r = r.reshape(10000, 4)
            edits["wx"] = (r[:, 0] % (dims[0] * 16)).astype(np.int32)
            edits["wy"] = (r[:, 1] % (dims[1] * 16)).astype(np.int32)
            edits["wz"] = (r[:, 2] % (dims[2] * 16)).astype(np.int32)
            edits["block"] = (r[:, 3] & 1).astype(np.uint32)  # 50 % destroy / 50 % place dirt
            ctx.sync()
            t0 = time.perf_counter()
            if frame % 2 == 0:   # fused call: dirty list stays on the device between the edit kernels and the mesh kernel
                dirty, entries, total = ctx.apply_edits_and_remesh(edits)
                t1 = t2 = time.perf_counter()
                fused_ms.append((t2 - t0) * 1e3)
                continue
            dirty = ctx.apply_edits(edits)
            t1 = time.perf_counter()
            entries, total = ctx.mesh_chunks(dirty)   # whole-chunk remesh of the dirty set: the reference's only granularity
            t2 = time.perf_counter()
            if frame >= 2:
                frame_ms.append((t2 - t0) * 1e3); apply_ms.append((t1 - t0) * 1e3); mesh_ms.append((t2 - t1) * 1e3); dirty_counts.append(len(dirty))
        print(json.dumps({"config": 4, "workload": "10,000 random block edits per frame over a 64x64x16-chunk world (65,536 chunks, 512 MiB voxels resident)",
                          "frames": args.frames, "dirty_chunks_per_frame": float(np.mean(dirty_counts)), "frame_ms": float(np.median(frame_ms)),
                          "apply_edits_ms": float(np.median(apply_ms)), "remesh_ms": float(np.median(mesh_ms)),
                          "fused_call_frame_ms": float(np.median(fused_ms[1:])), "edits_per_s_fused": 10000 / (np.median(fused_ms[1:]) * 1e-3),
                          "edits_per_s": 10000 / (np.median(frame_ms) * 1e-3), "dirty_chunks_per_s": float(np.mean(dirty_counts)) / (np.median(frame_ms) * 1e-3),
                          "timing": "host wall clock around dsp_apply_edits + dsp_mesh_chunks (both blocking), edits in pageable host memory"}), flush=True)


if __name__ == "__main__":
    main()
