#!/usr/bin/env python
"""Target for ncu: halo-mode mesh of an all-solid 16^3-chunk box (buried chunks dominate)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else capi.MESH_HALO
with capi.Context(0, 4096, 1 << 30) as ctx:
    ctx.set_uv_table(UV)
    first = ctx.generate_region((0, -16, 0), (16, 16, 16), capi.GEN_REFERENCE, 0)
    for _ in range(4):
        ctx.mesh_range(first, 4096, mode)
    n, t, last = ctx.mesh_kernel_stats()
    print("avg us", t / n * 1e3)
