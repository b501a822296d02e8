#!/usr/bin/env python
"""Does rotating over R copies of the bench region (footprint >> L2, no flush, steps back to back, ONE event pair) agree with
flushed per-step events?  Prints us per step for both methods."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
N, R, K = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 4, 200
with capi.Context(0, N * R, (R * 400 << 20)) as ctx:
    ctx.set_uv_table(UV)
    firsts = [ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, 0xD5E50001) for _ in range(R)]
    _, total = ctx.mesh_range(firsts[0], N)
    slice_bytes = (total + (1 << 20)) & ~127
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    ctx.set_kernel_timing(False)
    for k in range(10):
        ctx.mesh_range_async(firsts[k % R], N, 0, (k % R) * slice_bytes)
    ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for k in range(K):
        ctx.mesh_range_async(firsts[k % R], N, 0, (k % R) * slice_bytes)
    b.record(stream)
    torch.cuda.synchronize()
    print(f"rotating over {R} copies, back to back: {a.elapsed_time(b) / K * 1e3:.2f} us per step")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        ev[k][0].record(stream)
        ctx.mesh_range_async(firsts[0], N, 0, 0)
        ev[k][1].record(stream)
        with torch.cuda.stream(stream):
            flush.zero_()
    torch.cuda.synchronize()
    print(f"same copy, L2 flushed, per-step events:  {np.mean([x.elapsed_time(y) for x, y in ev]) * 1e3:.2f} us per step")
