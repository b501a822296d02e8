#!/usr/bin/env python
"""ncu target of round 2: one pass over the kernels whose numbers profiles/README.md quotes, few launches each.
  1. headline: parity mesh of the bench region (mesh_chunks_bump_kernel), 3 launches over rotating copies
  2. merged quads with summaries (classify_greedy_kernel + mesh_chunks_quads_kernel), 2 calls; then without (every chunk in the kernel), 2 calls
  3. halo mode on a 64 x 16 x 64-chunk noise world (classify_chunks_kernel + halo mesh kernel), 2 calls
  4. a region seam between two contexts on this GPU: seam_push_kernel + halo mesh with the in-kernel wait
python tools/prof_r02.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi, synthetic as syn

with capi.Context(0, 4096 * 4, 4 * (448 << 20)) as ctx:
    ctx.set_uv_table(syn.UV_DEFAULT)
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, syn.SEED_NOISE)
    firsts = [first] + [ctx.clone_region(first, (0, 4096 * k, 0)) for k in range(1, 4)]
    for k in range(3):
        ctx.mesh_range(firsts[k], 4096, capi.MESH_PARITY, k * (448 << 20))
    for k in range(2):
        ctx.mesh_range(firsts[k], 4096, capi.MESH_GREEDY | capi.MESH_INDEXED, k * (448 << 20))
    ctx.invalidate_chunks(np.arange(firsts[2], firsts[2] + 8192, dtype=np.uint32))
    for k in (2, 3):
        ctx.mesh_range(firsts[k], 4096, capi.MESH_GREEDY | capi.MESH_INDEXED, k * (448 << 20))
with capi.Context(0, 65536, 2 << 30) as ctx:
    ctx.set_uv_table(syn.UV_DEFAULT)
    first = ctx.generate_region((0, 0, 0), (64, 16, 64), capi.GEN_NOISE, syn.SEED_NOISE)
    for _ in range(2):
        ctx.mesh_range(first, 65536, capi.MESH_HALO)
with capi.Context(0, 32768, 1 << 30) as a, capi.Context(0, 32768, 1 << 30) as b:
    for c in (a, b):
        c.set_uv_table(syn.UV_DEFAULT)
    fa = a.generate_region((0, 0, 0), (32, 16, 64), capi.GEN_NOISE, syn.SEED_NOISE)
    fb = b.generate_region((32, 0, 0), (32, 16, 64), capi.GEN_NOISE, syn.SEED_NOISE)
    sa = np.array([fa + 31 + 32 * (y + 16 * z) for z in range(64) for y in range(16)], dtype=np.uint32)
    sb = np.array([fb + 0 + 32 * (y + 16 * z) for z in range(64) for y in range(16)], dtype=np.uint32)
    ia, ha = a.seam_create(capi.FACE_LEFT, sa)
    ib, hb = b.seam_create(capi.FACE_RIGHT, sb)
    a.seam_connect(ia, hb); b.seam_connect(ib, ha)
    for _ in range(2):
        a.seam_exchange_async(); b.seam_exchange_async()
        a.mesh_range_async(fa, 32768, capi.MESH_HALO); b.mesh_range_async(fb, 32768, capi.MESH_HALO)
        a.sync(); b.sync()
print("done")
