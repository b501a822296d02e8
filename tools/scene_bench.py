#!/usr/bin/env python
"""dsp_scene_update at growing render distances: the first frame (everything newly visible) and single-chunk crossings along +X
(only the shell changes), wall clock of the blocking call.  Noise terrain so that the meshes are non-trivial.
    python tools/scene_bench.py [--dist 10 4 --dist 32 8 ...]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import capi, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--dist", type=int, nargs=2, action="append")
ap.add_argument("--crossings", type=int, default=12)
args = ap.parse_args()
for h, v in (args.dist or [(10, 4), (32, 8), (64, 8)]):
    cam = np.array([8.0, 120.0, 8.0], dtype=np.float32)
    n_vis = len(capi.visible_chunks(cam, h, v))
    with capi.Context(0, int(n_vis * 1.3) + 4096, min(64 << 30, max(1 << 30, n_vis * 300_000)) // 128 * 128) as ctx:
        ctx.set_uv_table(syn.UV_DEFAULT)
        t0 = time.perf_counter()
        st = ctx.scene_update(cam, h, v, capi.GEN_NOISE, syn.SEED_NOISE)
        first_ms = (time.perf_counter() - t0) * 1e3
        ts, loaded, unloaded = [], [], []
        for k in range(args.crossings):
            cam[0] += 16.0
            t0 = time.perf_counter()
            st = ctx.scene_update(cam, h, v, capi.GEN_NOISE, syn.SEED_NOISE)
            ts.append((time.perf_counter() - t0) * 1e3)
            loaded.append(int(st["n_loaded_now"])); unloaded.append(int(st["n_unloaded_now"]))
        t0 = time.perf_counter()
        pos, ent = ctx.scene_draw_list()
        draw_ms = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"lib": os.path.basename(capi.LIB_PATH), "render_distance": [h, v], "visible_chunks": n_vis, "first_frame_ms": round(first_ms, 3),
                          "crossing_ms_median": round(float(np.median(ts)), 3), "crossing_ms_min": round(float(np.min(ts)), 3),
                          "chunks_loaded_per_crossing": int(np.median(loaded)), "chunks_unloaded_per_crossing": int(np.median(unloaded)),
                          "meshes": int(st["n_meshes"]), "vertices": int(st["vertex_total"]), "draw_list_ms": round(draw_ms, 3)}), flush=True)
