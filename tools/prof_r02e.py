#!/usr/bin/env python
"""Target for ncu: three calls of dsp_mesh_host_chunks_ex on the bench region through the packed upload (unpack_host_chunks_kernel +
mesh_chunks_bump_kernel per piece), then one call with DSP_MESH_GREEDY | DSP_MESH_INDEXED."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
N = 4096
pos = np.array([(a, b, c) for c in range(16) for b in range(16) for a in range(16)], dtype=np.int32)
host = torch.empty((N, 4096), dtype=torch.int16).pin_memory()
hv = host.numpy().view(np.uint16)
with capi.Context(0, N, 1 << 30) as ctx:
    ctx.set_uv_table(UV)
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, 0xD5E50001)
    for i in range(N):
        hv[i] = ctx.download_chunk(first + i)[0]
pal = np.where((hv == 0).all(axis=1), 1, 3).astype(np.uint32)
ent = np.zeros(N, dtype=capi.MESH_ENTRY_DTYPE)
with capi.Context(0, N, 1 << 30) as ctx:
    ctx.set_uv_table(UV)
    for _ in range(3):
        ctx.mesh_host_chunks_ex_ptr(N, pos, host.data_ptr(), pal, ent, 0)
    print(ctx.host_upload_stats(), int(ent["vertex_count"].astype(np.int64).sum()))
    ctx.mesh_host_chunks_ex_ptr(N, pos, host.data_ptr(), pal, ent, capi.MESH_GREEDY | capi.MESH_INDEXED)
