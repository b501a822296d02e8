"""Debug: per-CTA timeline of one merged-quad launch on the bench region (globaltimer stamps of thread 0).
Needs the -DDSP_PHASE_TIMERS build:  python tools/build_timers.py && DSP_LIB_PATH=despawnengine_b200/lib/libdespawn_timers.so python tools/quad_timeline.py [mode]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 12
L = capi.lib()
L.dsp_debug_timeline.restype = C.c_int
L.dsp_debug_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64]
with capi.Context(0, 4096, 1 << 30) as ctx:
    ctx.set_uv_table(UV)
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, 0xD5E50001)
    for case in ("summaries", "no summaries"):
        if case == "no summaries":
            ctx.invalidate_chunks(np.arange(first, first + 4096, dtype=np.uint32))
        for _ in range(5):
            ctx.mesh_range(first, 4096, mode)
        out = np.zeros(1024 * 16, dtype=np.uint64)
        for rep in range(3):
            L.dsp_debug_timeline(ctx._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size)  # clear
            ctx.mesh_range(first, 4096, mode)
            L.dsp_debug_timeline(ctx._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size)
            t = out.reshape(1024, 16).astype(np.int64)
            t = t[t[:, 0] > 0]
            t0 = t[:, 0].min()
            rel = lambda c: (c - t0) / 1e3
            n, tot, last = ctx.mesh_kernel_stats()
            nch = t[:, 10]
            q = lambda a: " ".join(f"{v:6.1f}" for v in np.percentile(a, [0, 10, 50, 90, 100]))
            print(f"[{case}] rep {rep}: {len(t)} CTAs, call {last*1e3:.1f} us; chunks/CTA histogram {np.bincount(nch, minlength=5)[:8].tolist()}")
            print(f"  (us since first CTA start; min p10 p50 p90 max)")
            print(f"  CTA start          {q(rel(t[:, 0]))}")
            print(f"  prologue done      {q(rel(t[:, 1]))}")
            has = t[:, 11] > 0
            print(f"  first tile in      {q(rel(t[has, 11]))}")
            for k in range(4):
                sel = nch > k
                if sel.any():
                    print(f"  iteration {k} done   {q(rel(t[sel, 2 + k]))}   ({sel.sum()} CTAs)")
            print(f"  exit               {q(rel(t[:, 8]))}")
            persm = np.bincount(t[:, 9], weights=nch, minlength=148)
            print(f"  iterations per SM: min {persm.min():.0f} max {persm.max():.0f}")
