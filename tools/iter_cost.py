#!/usr/bin/env python
"""Fixed vs per-iteration cost of the persistent mesh kernels: n = 592 .. 32768 all-air / all-solid chunks."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)


def kernel_us(ctx, first, n, mode, reps=20):
    ctx.mesh_range(first, n, mode)
    n0, t0, _ = ctx.mesh_kernel_stats()
    for _ in range(reps):
        ctx.mesh_range_async(first, n, mode)
    ctx.sync()
    n1, t1, _ = ctx.mesh_kernel_stats()
    return round((t1 - t0) / (n1 - n0) * 1e3, 2)


for name, oy in (("empty", 1), ("solid", -40)):
    with capi.Context(0, 32768, 1 << 30) as ctx:
        ctx.set_uv_table(UV)
        first = ctx.generate_region((0, oy, 0), (32, 32, 32), capi.GEN_REFERENCE, 0)
        for mode, mname in ((0, "parity"), (12, "greedy_indexed")):
            if name == "solid" and mode == 0:
                continue
            print(json.dumps({"case": name, "mode": mname, **{str(n): kernel_us(ctx, first, n, mode) for n in (1, 148, 592, 1184, 2368, 4096, 8192, 16384, 32768)}}), flush=True)
