#!/usr/bin/env python
"""Small run of every kernel (written as a compute-sanitizer target; the tool is closed on this GPU pool, so races are hunted by
tests/test_gpu_stress.py instead): terrain kinds, parity / halo / ordered / quad modes (incl. chunks with more
than one record window), edits on both addressing paths, scene update, border slabs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
with capi.Context(0, 512, 256 << 20) as ctx:
    ctx.set_uv_table(UV)
    firsts = [ctx.generate_region((0, 6, 0), (3, 3, 3), capi.GEN_NOISE, 0xD5E50001), ctx.generate_region((8, 0, 0), (2, 2, 2), capi.GEN_RANDOM, 3),
              ctx.generate_region((16, 0, 0), (2, 1, 2), capi.GEN_CHECKER, 0), ctx.generate_region((24, -1, 0), (2, 3, 2), capi.GEN_REFERENCE, 0)]
    sizes = [27, 8, 4, 12]
    off = 0
    for mode in (0, capi.MESH_HALO, capi.MESH_ORDERED, capi.MESH_GREEDY, capi.MESH_INDEXED, capi.MESH_GREEDY | capi.MESH_INDEXED, capi.MESH_GREEDY | capi.MESH_HALO):
        for f, n in zip(firsts, sizes):
            e, total = ctx.mesh_range(f, n, mode, 0)
    rng = np.random.default_rng(0)
    edits = np.zeros(500, dtype=capi.EDIT_DTYPE)
    edits["wx"], edits["wy"], edits["wz"] = rng.integers(0, 48, 500), rng.integers(96, 144, 500), rng.integers(0, 48, 500)
    edits["block"] = rng.integers(0, 3, 500)
    ctx.apply_edits_and_remesh(edits)
    extra = ctx.generate_chunks([(100, 0, 0), (101, 0, 0)], capi.GEN_NOISE, 1)
    edits["wx"] += 1600; edits["wy"] -= 96
    ctx.apply_edits_and_remesh(edits)
    import torch
    buf = torch.empty(2 * 256, dtype=torch.int16, device="cuda")
    ctx.pack_border_slabs(extra, capi.FACE_LEFT, buf.data_ptr()); ctx.sync()
    ctx.set_halo_slabs(extra, capi.FACE_RIGHT, buf.data_ptr())
    ctx.mesh_chunks(extra, capi.MESH_HALO)
with capi.Context(0, 512, 64 << 20) as ctx:
    ctx.set_uv_table(UV)
    for cam in ((2.5, 22.5, -2.5), (20.0, 22.5, -2.5), (300.0, 5.0, 9.0)):
        ctx.scene_update(cam, 2, 1)
print("sanitize target done")
