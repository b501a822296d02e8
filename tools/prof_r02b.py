#!/usr/bin/env python
"""ncu target, round 2 second session: the kernels that changed.
  1. parity mesh of the bench region with the generator's chunk summaries (mesh_chunks_bump_kernel), 3 launches over rotating copies
  2. DSP_MESH_ORDERED: count_chunks_scan_kernel + mesh_chunks_bump_kernel at preset offsets, 2 calls
  3. merged quads, indexed layout (classify_greedy_kernel + mesh_chunks_quads_kernel with direct stores), 2 calls
  4. parity mesh with the summaries invalidated (every tile read), 2 launches
  5. dsp_mesh_host_chunks_ex on pinned voxels with palette lengths: stage_host_chunks_kernel + the mesh kernel pulling host tiles, 2 + 2 launches
python tools/prof_r02b.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi, synthetic as syn

SL = 448 << 20
with capi.Context(0, 4096 * 4, 4 * SL) as ctx:
    ctx.set_uv_table(syn.UV_DEFAULT)
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, syn.SEED_NOISE)
    firsts = [first] + [ctx.clone_region(first, (0, 4096 * k, 0)) for k in range(1, 4)]
    for k in range(3):
        ctx.mesh_range(firsts[k], 4096, capi.MESH_PARITY, k * SL)
    for k in range(2):
        ctx.mesh_range(firsts[k + 1], 4096, capi.MESH_ORDERED, (k + 1) * SL)
    for k in range(2):
        ctx.mesh_range(firsts[k], 4096, capi.MESH_GREEDY | capi.MESH_INDEXED, k * SL)
    ctx.invalidate_chunks(np.arange(firsts[2], firsts[2] + 8192, dtype=np.uint32))
    for k in (2, 3):
        ctx.mesh_range(firsts[k], 4096, capi.MESH_PARITY, k * SL)
    # 5. host-resident chunks through dsp_mesh_host_chunks_ex: staging kernel + mesh kernel pulling its tiles out of pinned host memory, two launches
    import torch
    pos = syn.box_positions((0, 0, 0), (16, 16, 16))
    vox = torch.empty((4096, 4096), dtype=torch.int16).pin_memory()
    hv = vox.numpy().view(np.uint16)
    for i in range(4096):
        hv[i] = ctx.download_chunk(first + i)[0]
    pal = np.where((hv == 0).all(axis=1), 1, 3).astype(np.uint32)
    ent = np.zeros(4096, dtype=capi.MESH_ENTRY_DTYPE)
    ctx.mesh_host_chunks_ex_ptr(4096, pos, vox.data_ptr(), pal, ent)
print("done")
