#!/usr/bin/env python
"""Halo-mode mesh of a world WITHOUT generator summaries (every chunk invalidated: the state of a world that came from the host or
was edited everywhere): the border-mask refresh of the first call also leaves the chunk summaries, so later calls are classified
like a generated world.  DSP_LIB_PATH picks the build for A/B runs."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi, synthetic as syn

dims = (64, 16, 64)
n = dims[0] * dims[1] * dims[2]
out = {"lib": os.path.basename(capi.LIB_PATH), "chunks": n}
with capi.Context(0, n, 4 << 30) as ctx:
    ctx.set_uv_table(syn.UV_DEFAULT)
    first = ctx.generate_region((0, 0, 0), dims, capi.GEN_NOISE, syn.SEED_NOISE)
    e_ref, t_ref = ctx.mesh_range(first, n, capi.MESH_HALO)
    for label in ("generated", "invalidated"):
        if label == "invalidated":
            ctx.invalidate_chunks(np.arange(first, first + n, dtype=np.uint32))
        ts = []
        for k in range(4):
            ctx.sync(); t0 = time.perf_counter()
            e, t = ctx.mesh_range(first, n, capi.MESH_HALO)
            ts.append(round((time.perf_counter() - t0) * 1e3, 3))
            assert t == t_ref and np.array_equal(e["vertex_count"], e_ref["vertex_count"])
        out[label + "_call_ms"] = ts
print(json.dumps(out))
