"""Debug: where a CTA of the merged-quad kernel spends its cycles (thread 0 and thread 255), per surface chunk.
Needs the -DDSP_PHASE_TIMERS build:  python tools/build_timers.py && DSP_LIB_PATH=despawnengine_b200/lib/libdespawn_timers.so python tools/quad_phase_timers.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
names = {0: "tile wait (loop top .. mbarrier)", 1: "occupancy masks + barrier 1", 9: "deferred store of previous chunk", 2: "face/eq masks, shuffles, stores", 10: "barrier 2",
         3: "continued runs, seg scan, count", 11: "barrier 3 (scan)", 4: "records", 13: "bulk_wait_read + barrier 4 (records)", 12: "extent pass + barrier", 5: "tile loop: geometry, stores, inter-tile waits",
         14: "barrier 5 (staging)", 6: "store issue / loop end (+ whole uniform/air iterations)", 8: "prologue (first tile, uniform list)", 7: "exit"}
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 4
with capi.Context(0, 4096, 1 << 30) as ctx:
    ctx.set_uv_table(UV)
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, 0xD5E50001)
    for _ in range(3):
        ctx.mesh_range(first, 4096, mode)
    ctx.debug_phase_cycles()
    reps = 10
    for _ in range(reps):
        ctx.mesh_range(first, 4096, mode)
    ph = ctx.debug_phase_cycles().astype(np.float64)
    n, tot_ms, last = ctx.mesh_kernel_stats()
    ctas = ph[15] / reps
    for who, base in (("thread 0", 0), ("thread 255", 16)):
        tot = sum(ph[base + k] for k in names)
        print(f"{who}: {ctas:.0f} CTAs, kernel {last*1e3:.1f} us; cycles per CTA per launch: {tot/reps/ctas:.0f}")
        for k, nm in names.items():
            print(f"  {nm:58s} {100*ph[base+k]/tot:5.1f} %   {ph[base+k]/reps/ctas:8.0f} cycles/CTA")
