#!/usr/bin/env python
"""dsp_mesh_host_chunks_ex on the bench region, pinned host voxels, entry table out: copy-engine pipeline (DSP_E2E_PULL=0) against the
kernel pulling its tiles itself, without and with the reference's palette lengths (all-air chunks are then never fetched)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi, synthetic as syn
N = 4096
pos = syn.box_positions((0, 0, 0), (16, 16, 16))
vox = torch.empty((N, 4096), dtype=torch.int16).pin_memory()
hv = vox.numpy().view(np.uint16)
with capi.Context(0, N, 1 << 30) as ctx:
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, syn.SEED_NOISE)
    for i in range(N):
        hv[i] = ctx.download_chunk(first + i)[0]
pal = np.where((hv == 0).all(axis=1), 1, 3).astype(np.uint32)
ent = torch.empty(N * 16, dtype=torch.uint8).pin_memory()
ent_np = ent.numpy().view(capi.MESH_ENTRY_DTYPE)
out = {"air_chunks": int((pal == 1).sum())}
cases = [("copy_engine", "0", None, "8"), ("pull", "2", None, "8"), ("pull_palette_lens", "1", pal, "8"), ("copy_engine_again", "0", None, "8")]
cases += [("pull_palette_lens_one_stream", "1", pal, "8")] + [(f"pull_palette_lens_div{d}", "1", pal, str(d)) for d in (4, 6, 12, 16)] + [("pull_palette_lens_again", "1", pal, "8")]
host = torch.empty(448 << 20, dtype=torch.uint8).pin_memory()
for label, pull, pl, div in cases:
    os.environ["DSP_E2E_PULL"] = pull
    os.environ["DSP_E2E_PULL_DIV"] = div
    os.environ["DSP_E2E_PULL_PIECES"] = "1" if "one_stream" in label else "2"
    with capi.Context(0, N, 1 << 30) as ctx:
        ctx.set_uv_table(syn.UV_DEFAULT)
        fn = lambda: ctx.mesh_host_chunks_ex_ptr(N, pos, vox.data_ptr(), pl, ent_np)
        for _ in range(3):
            total = fn()
        ts = []
        for _ in range(30):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        out[label + "_ms"] = round(float(np.median(ts)) * 1e3, 4)
        out[label + "_ms_min"] = round(float(np.min(ts)) * 1e3, 4)
        out[label + "_total_bytes"] = total
        if label in ("copy_engine", "pull_palette_lens"):  # the merged + indexed leg with its vertices delivered to pinned host memory
            gi = capi.MESH_GREEDY | capi.MESH_INDEXED
            fn2 = lambda: ctx.mesh_host_chunks_ex_ptr(N, pos, vox.data_ptr(), pl, ent_np, gi, 0, host.data_ptr(), 448 << 20)
            for _ in range(3):
                fn2()
            ts = []
            for _ in range(20):
                t0 = time.perf_counter(); fn2(); ts.append(time.perf_counter() - t0)
            out[label + "_greedy_indexed_to_host_ms"] = round(float(np.median(ts)) * 1e3, 4)
print(json.dumps(out))
