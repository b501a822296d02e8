#!/usr/bin/env python
"""Target for ncu: the terrain fill kernels on a 64x16x64-chunk box (512 MiB of voxels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi
for kind, seed, oy in ((capi.GEN_REFERENCE, 0, -8), (capi.GEN_NOISE, 0xD5E50001, 0), (capi.GEN_RANDOM, 3, 0)):
    with capi.Context(0, 65536, 1 << 20) as ctx:
        ctx.generate_region((0, oy, 0), (64, 16, 64), kind, seed)
        ctx.sync()
