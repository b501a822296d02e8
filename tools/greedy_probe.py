#!/usr/bin/env python
"""Kernel time of the merged-quad modes on the bench region (16^3 chunks of noise terrain), with the generators' chunk
summaries in place (pre-pass finishes air / uniform chunks) and with them invalidated (every chunk reaches the mesh kernel:
what a host-uploaded world looks like).   python tools/greedy_probe.py [--once]   (--once: one launch per case, for ncu)"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
once = "--once" in sys.argv


def kernel_ms(ctx, fn, reps):
    fn(); ctx.sync()
    n0, t0, _ = ctx.mesh_kernel_stats()
    for _ in range(reps):
        fn()
    ctx.sync()
    n1, t1, _ = ctx.mesh_kernel_stats()
    return (t1 - t0) / (n1 - n0)


def wall_us(ctx, fn, reps):
    import time
    fn(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return (time.perf_counter() - t0) / reps * 1e6


with capi.Context(0, 4096, 1 << 30) as ctx:
    ctx.set_uv_table(UV)
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, 0xD5E50001)
    for label in ("summaries", "no summaries"):
        if label == "no summaries":
            ctx.invalidate_chunks(np.arange(first, first + 4096, dtype=np.uint32))
        out = {"case": label}
        for mname, mode in (("greedy", 4), ("greedy_indexed", 12)):
            e, total = ctx.mesh_range(first, 4096, mode)
            if once:
                continue
            out[mname + "_mesh_kernel_us"] = round(kernel_ms(ctx, lambda: ctx.mesh_range_async(first, 4096, mode), 20) * 1e3, 2)
            out[mname + "_call_us_back_to_back"] = round(wall_us(ctx, lambda: ctx.mesh_range_async(first, 4096, mode), 200), 2)
            out[mname + "_MB"] = round(total / 1e6, 2)
            out[mname + "_quads"] = int(e["vertex_count"].astype(np.int64).sum()) // (4 if mode & 8 else 6)
        print(json.dumps(out), flush=True)
