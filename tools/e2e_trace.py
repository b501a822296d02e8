#!/usr/bin/env python
"""DSP_TRACE=1 python tools/e2e_trace.py: host-side timeline of dsp_mesh_host_chunks on the bench region."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
N = 4096
with capi.Context(0, N, 1 << 30) as ctx:
    ctx.set_uv_table(UV)
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, 0xD5E50001)
    pos = np.array([(a, b, c) for c in range(16) for b in range(16) for a in range(16)], dtype=np.int32)
    host = torch.empty((N, 4096), dtype=torch.int16).pin_memory()
    hv = host.numpy().view(np.uint16)
    for i in range(N):
        hv[i] = ctx.download_chunk(first + i)[0]
    ent = torch.empty(N * 16, dtype=torch.uint8).pin_memory()
    ent_np = ent.numpy().view(capi.MESH_ENTRY_DTYPE)
    os.environ.pop("DSP_TRACE", None)
    for _ in range(3):
        ctx.mesh_host_chunks_ptr(N, pos, host.data_ptr(), ent_np, 0)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); ctx.mesh_host_chunks_ptr(N, pos, host.data_ptr(), ent_np, 0); ts.append(time.perf_counter() - t0)
    print("untraced step us:", [round(t * 1e6) for t in ts])
    os.environ["DSP_TRACE"] = "1"
    ctx.mesh_host_chunks_ptr(N, pos, host.data_ptr(), ent_np, 0)
    # the same chunks with their palette lengths: kernel pull (dsp_mesh_host_chunks_ex)
    os.environ.pop("DSP_TRACE", None)
    pal = np.where((hv == 0).all(axis=1), 1, 3).astype(np.uint32)
    for _ in range(3):
        ctx.mesh_host_chunks_ex_ptr(N, pos, host.data_ptr(), pal, ent_np)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); ctx.mesh_host_chunks_ex_ptr(N, pos, host.data_ptr(), pal, ent_np); ts.append(time.perf_counter() - t0)
    print("untraced step us, kernel pull with palette lengths:", [round(t * 1e6) for t in ts])
    os.environ["DSP_TRACE"] = "1"
    ctx.mesh_host_chunks_ex_ptr(N, pos, host.data_ptr(), pal, ent_np)
