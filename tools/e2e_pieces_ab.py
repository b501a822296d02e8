#!/usr/bin/env python
"""Same-process A/B of the packed upload's piece policy: two contexts side by side -- pieces cut by content (default) and three
pieces fixed up front (DSP_E2E_PACK_PIECES=3, the former default) -- called alternately on the bench region; medians of 80 calls each.
Second pair: the same with DSP_MESH_GREEDY | DSP_MESH_INDEXED.  Third: a dense region (checkerboard chunks), fourth: a sparse one."""
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
with capi.Context(0, N, 64 << 20) as c0:
    first = c0.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, syn.SEED_NOISE)
    for i in range(N):
        hv[i] = c0.download_chunk(first + i)[0]
noise = hv.copy()


def make(env):
    for k, v in env.items():
        os.environ[k] = v
    ctx = capi.Context(0, N, 3 << 30)
    for k in env:
        os.environ.pop(k)
    ctx.set_uv_table(syn.UV_DEFAULT)
    ctx.set_kernel_timing(False)
    return ctx


A, B = make({}), make({"DSP_E2E_PACK_PIECES": "3"})
out = {}


def run(label, pal, mode):
    fa = lambda: A.mesh_host_chunks_ex_ptr(N, pos, vox.data_ptr(), pal, ent_np, mode)
    fb = lambda: B.mesh_host_chunks_ex_ptr(N, pos, vox.data_ptr(), pal, ent_np, mode)
    for _ in range(5):
        ta = fa(); tb = fb()
    assert ta == tb, (ta, tb)
    sa, sb = [], []
    for _ in range(80):
        t0 = time.perf_counter(); fa(); t1 = time.perf_counter(); fb(); t2 = time.perf_counter()
        sa.append(t1 - t0); sb.append(t2 - t1)
    out[label] = {"adaptive_us": round(float(np.median(sa)) * 1e6, 1), "three_fixed_us": round(float(np.median(sb)) * 1e6, 1),
                  "adaptive_min_us": round(float(np.min(sa)) * 1e6, 1), "three_fixed_min_us": round(float(np.min(sb)) * 1e6, 1), "bytes": int(ta)}


pal = np.where((hv == 0).all(axis=1), 1, 3).astype(np.uint32)
run("noise region", pal, capi.MESH_PARITY)
run("noise region, no palette lengths", None, capi.MESH_PARITY)
run("noise region, merged + indexed", pal, capi.MESH_GREEDY | capi.MESH_INDEXED)
idx = np.arange(4096)
hv[:] = ((((idx & 15) + ((idx >> 4) & 15) + (idx >> 8)) & 1) == 0).astype(np.uint16)[None, :]   # every chunk a 3-D checkerboard (1.47 MB of vertices each)
hv[::2] = 0                                                                                       # ... every other chunk air (the arena holds 3 GiB)
run("checkerboard chunks, half of them air", np.where((hv == 0).all(axis=1), 1, 2).astype(np.uint32), capi.MESH_PARITY)
hv[:] = 0
hv[::9, :256] = 1                                                                                 # one chunk in nine holds a layer
run("sparse: one chunk in nine", np.where((hv == 0).all(axis=1), 1, 2).astype(np.uint32), capi.MESH_PARITY)
hv[:] = noise
A.close(); B.close()
print(json.dumps(out, indent=1))
