#!/usr/bin/env python
"""Call time of DSP_MESH_ORDERED against the default (completion-order) mode on the bench region, back to back and per call
(library events around the call's kernels).  DSP_LIB_PATH picks the build for A/B runs."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi, synthetic as syn


def kernel_us(ctx, fn, reps=30):
    fn(); ctx.sync()
    n0, t0, _ = ctx.mesh_kernel_stats()
    for _ in range(reps):
        fn()
    ctx.sync()
    n1, t1, _ = ctx.mesh_kernel_stats()
    return (t1 - t0) / (n1 - n0) * 1e3


def wall_us(ctx, fn, reps=200):
    fn(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return (time.perf_counter() - t0) / reps * 1e6


out = {"lib": os.path.basename(capi.LIB_PATH)}
with capi.Context(0, 4 * 4096, 8 << 30) as ctx:
    ctx.set_uv_table(syn.UV_DEFAULT)
    first = [ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, syn.SEED_NOISE)]
    for k in range(1, 4):
        first.append(ctx.clone_region(first[0], (1000 * k, 0, 0)))
    state = {"i": 0}

    def step(mode):
        i = state["i"] = (state["i"] + 1) % 4
        ctx.mesh_range_async(first[i], 4096, mode, arena_offset=i * (2 << 30))

    for name, mode in (("default", 0), ("ordered", capi.MESH_ORDERED)):
        out[name + "_kernels_us"] = round(kernel_us(ctx, lambda: step(mode)), 2)
        out[name + "_call_us_back_to_back"] = round(wall_us(ctx, lambda: step(mode)), 2)
print(json.dumps(out))
