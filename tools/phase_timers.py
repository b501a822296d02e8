"""Debug: prints the per-phase cycle breakdown of the single-pass mesh kernel (needs the DSP_PHASE_TIMERS build:
DSP_LIB_PATH=despawnengine_b200/lib/libdespawn_timers.so python tools/phase_timers.py)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
names = ["loop top", "records", "wait tile(next)", "count(next)", "look-back (warp0) / idle", "barrier B1", "emission", "ticket atomic", "barrier B2", "drain at exit", "sweep 1 (precount only)"]
with capi.Context(0, 4096, 1 << 30) as ctx:
    ctx.set_uv_table(UV)
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, 0xD5E50001)
    for _ in range(3):
        ctx.mesh_range(first, 4096)
    ctx.debug_phase_cycles()
    reps = 10
    for _ in range(reps):
        ctx.mesh_range(first, 4096)
    ph = ctx.debug_phase_cycles().astype(np.float64)
    n, tot_ms, last = ctx.mesh_kernel_stats()
    ctas = ph[15] / reps
    tot = ph[:11].sum()
    print(f"variant {os.environ.get('DSP_MESH_VARIANT','default')}: {ctas:.0f} CTAs, kernel {last*1e3:.1f} us; thread-0 cycles per CTA per launch: {tot/reps/ctas:.0f}")
    for k, nm in enumerate(names):
        print(f"  {nm:28s} {100*ph[k]/tot:5.1f} %   {ph[k]/reps/4096:8.0f} cycles/chunk")
