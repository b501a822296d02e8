#!/usr/bin/env python
"""A/B timing of the parity mesh kernel on the bench region (+ empty / solid / random regions); DSP_LIB_PATH picks the build."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
import torch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def kernel_us(ctx, first, mode, reps=30):
    ctx.mesh_range(first, 4096, mode)
    n0, t0, _ = ctx.mesh_kernel_stats()
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize()
        ctx.mesh_range_async(first, 4096, mode); ctx.sync()
    n1, t1, _ = ctx.mesh_kernel_stats()
    return (t1 - t0) / (n1 - n0) * 1e3


out = {"lib": os.path.basename(capi.LIB_PATH)}
for name, kind, seed, origin in (("noise", capi.GEN_NOISE, 0xD5E50001, (0, 0, 0)), ("empty", capi.GEN_REFERENCE, 0, (0, 1, 0)), ("solid", capi.GEN_REFERENCE, 0, (0, -16, 0)),
                                 ("random", capi.GEN_RANDOM, 0xD5E50003, (0, 0, 0))):
    with capi.Context(0, 4096, 7 << 30) as ctx:
        ctx.set_uv_table(UV)
        first = ctx.generate_region(origin, (16, 16, 16), kind, seed)
        out[name] = round(kernel_us(ctx, first, int(os.environ.get("AB_MODE", "0"))), 2)
print(json.dumps(out))
