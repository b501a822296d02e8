#!/usr/bin/env python
"""Terrain-fill kernels alone: GB/s of voxel bytes written (8,192 per chunk) for each generator, region and list forms."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi

dims = (128, 32, 128)  # 524,288 chunks = 4 GiB of voxels
n = dims[0] * dims[1] * dims[2]
for name, kind, seed in (("reference", capi.GEN_REFERENCE, 0), ("noise", capi.GEN_NOISE, 0xD5E50001), ("random", capi.GEN_RANDOM, 0xD5E50003), ("checker", capi.GEN_CHECKER, 0)):
    ts = []
    for rep in range(4):
        with capi.Context(0, n, 1 << 20) as ctx:
            ctx.sync()
            t0 = time.perf_counter()
            ctx.generate_region((0, -8 if kind == capi.GEN_REFERENCE else 0, 0), dims, kind, seed)
            ctx.sync()
            ts.append(time.perf_counter() - t0)
    t = min(ts[1:])
    print(json.dumps({"generator": name, "chunks": n, "ms": t * 1e3, "GBps_written": n * 8192 / t / 1e9, "frac_of_6453": n * 8192 / t / 1e9 / 6453.1}), flush=True)
