#!/usr/bin/env python
"""End-to-end legs side by side on the bench region: host voxels in -> {entries only, vertices to pinned host memory} x {reference
layout, greedy+indexed}.  DSP_TRACE=1 prints the host-side timeline of each call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi, synthetic as syn
N = 4096
with capi.Context(0, N, 1 << 30) as ctx:
    ctx.set_uv_table(syn.UV_DEFAULT)
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, syn.SEED_NOISE)
    pos = syn.box_positions((0, 0, 0), (16, 16, 16))
    vox = torch.empty((N, 4096), dtype=torch.int16).pin_memory()
    hv = vox.numpy().view(np.uint16)
    for i in range(N):
        hv[i] = ctx.download_chunk(first + i)[0]
    ent = np.zeros(N, dtype=capi.MESH_ENTRY_DTYPE)
    host = torch.empty(448 << 20, dtype=torch.uint8).pin_memory()
    for name, mode in (("parity", 0), ("greedy_indexed", 12)):
        for what in ("entries only", "vertices to host"):
            fn = (lambda: ctx.mesh_host_chunks_ptr(N, pos, vox.data_ptr(), ent, mode)) if what == "entries only" else \
                 (lambda: ctx.mesh_host_chunks_to_host_ptr(N, pos, vox.data_ptr(), ent, host.data_ptr(), 448 << 20, mode))
            for _ in range(3):
                fn()
            t0 = time.perf_counter()
            for _ in range(20):
                fn()
            print(f"{name:15s} {what:17s} {(time.perf_counter() - t0) / 20 * 1e3:7.3f} ms/call", flush=True)
