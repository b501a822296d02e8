#!/usr/bin/env python
"""Packed upload: call time of dsp_mesh_host_chunks_ex on the bench region against the number of host worker threads and the share of
chunks the host packs (DSP_HOST_THREADS x DSP_E2E_PACK_KEEP; the others cross the bus as they are).   python tools/e2e_pack_sweep.py"""
import os, sys, time
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
ent = torch.empty(N * 16, dtype=torch.uint8).pin_memory()
ent_np = ent.numpy().view(capi.MESH_ENTRY_DTYPE)
threads = [int(t) for t in sys.argv[1].split(",")] if len(sys.argv) > 1 else [3, 7, 11, 15]
for t in threads:
    row = []
    for keep in ("pull", 4, 6, 8, 10, 12, 14, 16, 0):
        os.environ["DSP_HOST_THREADS"] = str(t)
        os.environ["DSP_E2E_PACK"] = "0" if keep == "pull" else "1"
        os.environ["DSP_E2E_PACK_KEEP"] = "0" if keep == "pull" else str(keep)
        with capi.Context(0, N, 1 << 30) as ctx:
            ctx.set_uv_table(UV)
            ctx.set_kernel_timing(False)
            for _ in range(5):
                ctx.mesh_host_chunks_ex_ptr(N, pos, host.data_ptr(), pal, ent_np, 0)
            ts = []
            for _ in range(30):
                t0 = time.perf_counter(); ctx.mesh_host_chunks_ex_ptr(N, pos, host.data_ptr(), pal, ent_np, 0); ts.append(time.perf_counter() - t0)
            st = ctx.host_upload_stats()
            row.append(f"{'auto' if keep == 0 else keep}: {np.median(ts) * 1e6:.0f} us" + (f" (raw {st['raw']})" if keep == 0 else ""))
    print(f"workers {t:2d}  " + "   ".join(row), flush=True)
