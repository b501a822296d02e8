#!/usr/bin/env python
"""Per-path cost of the mesh kernels: 4,096 chunks that are all empty / all solid / noise / random, in each mode."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)


def kernel_ms(ctx, fn, reps):
    fn(); ctx.sync()
    n0, t0, _ = ctx.mesh_kernel_stats()
    for _ in range(reps):
        fn()
    ctx.sync()
    n1, t1, _ = ctx.mesh_kernel_stats()
    return (t1 - t0) / (n1 - n0)


cases = [("empty", capi.GEN_REFERENCE, 0, (0, 1, 0)), ("solid", capi.GEN_REFERENCE, 0, (0, -16, 0)), ("noise", capi.GEN_NOISE, 0xD5E50001, (0, 0, 0)),
         ("noise surface band y=6..9 x4", capi.GEN_NOISE, 0xD5E50001, None), ("random", capi.GEN_RANDOM, 0xD5E50003, (0, 0, 0))]
for name, kind, seed, origin in cases:
    with capi.Context(0, 4096, 7 << 30) as ctx:
        ctx.set_uv_table(UV)
        if origin is None:
            first = ctx.generate_region((0, 6, 0), (32, 4, 32), kind, seed)
        else:
            first = ctx.generate_region(origin, (16, 16, 16), kind, seed)
        out = {"case": name}
        for mname, mode in (("parity", 0), ("indexed", 8), ("greedy", 4), ("greedy_indexed", 12), ("halo", 1), ("halo_greedy_indexed", 13)):
            e, total = ctx.mesh_range(first, 4096, mode)
            out[mname + "_us"] = round(kernel_ms(ctx, lambda: ctx.mesh_range_async(first, 4096, mode), 20) * 1e3, 2)
            out[mname + "_MB"] = round(total / 1e6, 1)
        print(json.dumps(out), flush=True)
