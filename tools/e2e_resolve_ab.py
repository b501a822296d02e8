#!/usr/bin/env python
"""dsp_mesh_host_chunks_ex on the bench region (packed upload, pinned voxels, palette lengths, entry table out): wall time per call with
the positions resolving to box-region slots (bench.py's situation) and to map slots (a context that only ever saw host chunks).
DSP_LIB_PATH / DSP_E2E_PACK_PIECES select what is compared; DSP_TRACE_ONE=1 adds one traced call."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi, synthetic as syn
N = 4096
pos = syn.box_positions((0, 0, 0), (16, 16, 16))
vox = torch.empty((N, 4096), dtype=torch.int16).pin_memory()
hv = vox.numpy().view(np.uint16)
ent = torch.empty(N * 16, dtype=torch.uint8).pin_memory()
ent_np = ent.numpy().view(capi.MESH_ENTRY_DTYPE)
out = {"lib": os.path.basename(capi.LIB_PATH), "pieces": os.environ.get("DSP_E2E_PACK_PIECES", "default")}


def timed(ctx, pal, label):
    fn = lambda: ctx.mesh_host_chunks_ex_ptr(N, pos, vox.data_ptr(), pal, ent_np)
    for _ in range(5):
        fn()
    ts = []
    for _ in range(60):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    out[label + "_us_median"] = round(float(np.median(ts)) * 1e6, 1)
    out[label + "_us_min"] = round(float(np.min(ts)) * 1e6, 1)
    out[label + "_faces"] = int(ent_np["vertex_count"].astype(np.int64).sum()) // 6
    if os.environ.get("DSP_TRACE_ONE"):
        os.environ["DSP_TRACE"] = "1"
        for _ in range(int(os.environ["DSP_TRACE_ONE"])):
            fn()
            print("[trace] ----", file=sys.stderr, flush=True)
        os.environ.pop("DSP_TRACE")


with capi.Context(0, N, 1 << 30) as ctx:
    ctx.set_uv_table(syn.UV_DEFAULT)
    ctx.set_kernel_timing(False)
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, syn.SEED_NOISE)
    for i in range(N):
        hv[i] = ctx.download_chunk(first + i)[0]
    pal = np.where((hv == 0).all(axis=1), 1, 3).astype(np.uint32)
    timed(ctx, pal, "box_slots")
with capi.Context(0, N, 1 << 30) as ctx:
    ctx.set_uv_table(syn.UV_DEFAULT)
    ctx.set_kernel_timing(False)
    timed(ctx, pal, "map_slots")
print(json.dumps(out))
