#!/usr/bin/env python
"""bench.py -- chunks meshed/s on BASELINE.json's configs[1] (16x16x16-chunk region of seeded noise
terrain, 4,096 chunks = 256^3 voxels per GPU), device-resident (`value`), end to end from host buffers
through the C ABI (`e2e`), against the kernel's HBM roofline (`roofline`) and with the CPU
restatement of the reference timed beside it (`cpu_baseline`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by the driver through torch.distributed.run (one rank per GPU, NCCL); started
without it, this script re-executes itself under torch.distributed.run.  Weak scaling: every rank
meshes its own 16^3-chunk region (regions abut along X); chunks are independent in the reference's
mesher (chunk_mesh.rs:46-51 never looks outside the chunk), so there is no data-path collective.

A "step" is one pass of the hot path over the batch: the fused cull + scan + emit kernel over all
4,096 resident chunks of the rank (one launch), writing the full vertex stream into the HBM arena.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

# The contract is ONE JSON line on stdout.  Libraries below us write there too (NCCL prints "NCCL version ..." to fd 1 when
# a communicator is created), so fd 1 is pointed at stderr for the whole run and the line goes to the saved descriptor.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict) -> None:
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0xD5E50001
DIMS = (16, 16, 16)
N_CHUNKS = DIMS[0] * DIMS[1] * DIMS[2]
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
WORKLOAD = "configs[1]: 16x16x16-chunk region (256^3 voxels) of seeded integer value-noise terrain, reference cull semantics, unit quads"
METRIC = "chunks meshed/sec"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE mesh_chunks_bump_kernel launch of this exact workload, from the
# `ncu --set full` capture summarised in profiles/r01_final_bump_ncu_summary.csv (33.65 MB read + 337.92 MB written; the
# remaining ~59 MB of the 397 MB vertex stream was still dirty in the 126 MB L2 when the kernel ended).  Below the
# algorithmic 430.5 MB, i.e. no wasted re-reads.
NCU_TRAFFIC_BYTES = 33_651_968 + 337_916_928
NCU_TRAFFIC_SOURCE = "profiles/r01_final_bump_ncu_summary.csv (ncu --set full, one launch, same workload)"


def region_positions(origin):
    return np.array([(origin[0] + a, origin[1] + b, origin[2] + c) for c in range(DIMS[2]) for b in range(DIMS[1]) for a in range(DIMS[0])],
                    dtype=np.int32)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.power = [], set(), []
        self.stop_flag = threading.Event()
        self.sm_max = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, note=""):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "note": "NVML unavailable or no samples " + note}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "power_w_max": round(max(self.power), 1), "note": note}


def reference_arm(args, rank):
    """The reference's own CPU path.  Its Rust sources cannot be built in this image (no rustc/cargo), so this
    arm times the C++ restatement of build_chunk_mesh (oracle/, kind = "port") on all host threads."""
    if rank != 0:
        return
    from oracle import loader as orc
    threads = orc.hardware_threads()
    pos = region_positions((0, 0, 0))
    w = orc.Workload(pos, orc.GEN_NOISE, SEED, UV)
    for _ in range(max(1, min(args.warmup, 3))):
        w.mesh(orc.FAITHFUL, threads)
    secs, verts = [], 0
    budget_t0 = time.perf_counter()
    steps_done = 0
    for _ in range(args.steps):
        s, verts = w.mesh(orc.FAITHFUL, threads)
        secs.append(s)
        steps_done += 1
        if time.perf_counter() - budget_t0 > 150:  # keep the whole run within a few minutes
            break
    total = float(np.sum(secs))
    value = N_CHUNKS * steps_done / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": args.gpus, "steps": steps_done,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / steps_done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16 voxels -> f32 vertices", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chunks_per_step": N_CHUNKS, "vertices_per_step": verts,
                   "note": "C++ restatement of build_chunk_mesh (faithful flavour: per-voxel string-keyed lookups, unreserved Vec) over "
                           "host-resident chunks + one memcpy per chunk for the Vulkan upload; the Rust reference itself cannot be built here"},
        "voxels_per_s": value * 4096,
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": threads, "kind": "port",
                         "sample": f"all {N_CHUNKS} chunks of the workload per step, {steps_done} steps"},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (default min(steps, 40))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ordered", action="store_true", help="use DSP_MESH_ORDERED (slots packed in call order); default is completion order")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        port = os.environ.get("MASTER_PORT", "29577")
        os.dup2(_REAL_STDOUT.fileno(), 1)  # the re-executed ranks redirect for themselves
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                  "--master-addr", "127.0.0.1", "--master-port", port, os.path.abspath(__file__)] + sys.argv[1:])
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the mesh path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    from despawnengine_b200 import build as dsp_build
    from despawnengine_b200 import capi
    if local_rank == 0 and not os.path.exists(capi.LIB_PATH):
        dsp_build.build()
    if distributed:
        dist.barrier()

    def barrier():
        if distributed:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # COPIES identical copies of the region (same positions, same seed, different chunk slots) and as many arena slices: the
    # timed steps rotate over them, so that a step's input (32 MiB of voxels) and output (397 MB of vertices) were last
    # touched COPIES steps earlier -- 1.7 GB of traffic ago, against 126 MB of L2.  No flush is needed then, the steps run back
    # to back on the stream and ONE event pair times all of them (kernel launch latency overlaps the previous step, as it does
    # in a frame loop).
    COPIES, SLICE = 4, 448 << 20
    ctx = capi.Context(local_rank, N_CHUNKS * COPIES, COPIES * SLICE)
    ctx.set_uv_table(UV)
    origin = (DIMS[0] * rank, 0, 0)
    first = ctx.generate_region(origin, DIMS, capi.GEN_NOISE, SEED)
    # identical voxels at other places of the world (a position holds one chunk): same faces, same bytes per step
    firsts = [first] + [ctx.clone_region(first, (origin[0], 4096 * k, 0)) for k in range(1, COPIES)]
    mode = capi.MESH_ORDERED if args.ordered else capi.MESH_PARITY
    entries, total_bytes = ctx.mesh_range(first, N_CHUNKS, mode)
    assert total_bytes <= SLICE
    faces = int(entries["vertex_count"].astype(np.int64).sum()) // 6
    # algorithmic bytes per launch: voxels read once + vertices written once + position + entry
    algo_bytes = N_CHUNKS * (8192 + 16 + 16) + 120 * faces

    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def step(k: int):
        ctx.mesh_range_async(firsts[k % COPIES], N_CHUNKS, mode, (k % COPIES) * SLICE)

    # ---- device-resident leg: `value` -----------------------------------------------------------------
    for k in range(max(args.warmup, COPIES)):
        step(k)
    ctx.sync()
    # A step is exactly one launch of the mesh kernel (it resets its own control block: no memset in front); the library's own
    # per-launch event pair is switched off for the timed region and back on for the side measurements below.
    ctx.set_kernel_timing(False)
    launches0 = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    wall0 = time.perf_counter()
    ev0.record(stream)
    for k in range(args.steps):
        step(k)
    ev1.record(stream)
    torch.cuda.synchronize()
    wall1 = time.perf_counter()
    barrier()
    step_ms = [ev0.elapsed_time(ev1) / args.steps] * args.steps
    launches = ctx.kernel_launches - launches0
    if not args.ordered:
        assert launches == args.steps, (launches, args.steps)  # one kernel launch per step, nothing else of ours on the stream
    ctx.set_kernel_timing(True)
    # clocks: if the timed region was too short for NVML to see it, keep the same load running ~1.2 s more
    note = "sampled during the timed region"
    if len(sampler.samples) < 10:
        t_end = time.perf_counter() + 1.2
        while time.perf_counter() < t_end:
            for _ in range(20):
                ctx.mesh_range_async(first, N_CHUNKS, mode)
            ctx.sync()
        note = "timed region too short for NVML; sampled over the timed region plus ~1.2 s of the same step repeated"
    sampler.stop_flag.set()
    sampler.join()
    _ = ctx.mesh_kernel_stats()

    dev_ms_total = max_over_ranks(float(np.sum(step_ms)))
    value = world * N_CHUNKS * args.steps / (dev_ms_total * 1e-3)
    kernel_ms_avg = float(np.mean(step_ms))  # this rank's launches; `value` uses the max over ranks of the sum
    peak, peak_src = load_peaks()
    achieved = algo_bytes / (kernel_ms_avg * 1e-3) / 1e9

    # ---- end-to-end leg: host voxel buffers -> dsp_load_chunks -> dsp_mesh_chunks -> entries on host ----------
    e2e_steps = args.e2e_steps or min(args.steps, 40)
    pos = region_positions(origin)
    host_vox = torch.empty((N_CHUNKS, 4096), dtype=torch.int16).pin_memory()
    hv = host_vox.numpy().view(np.uint16)
    for i in range(N_CHUNKS):  # setup, untimed: host copies of the chunks, as a host caller would hold them (Chunk.blocks)
        hv[i] = ctx.download_chunk(first + i)[0]
    slots_out = np.empty(N_CHUNKS, dtype=np.uint32)
    entries_host = torch.empty(N_CHUNKS * 16, dtype=torch.uint8).pin_memory()
    entries_np = entries_host.numpy().view(capi.MESH_ENTRY_DTYPE)
    h2d = N_CHUNKS * (8192 + 16)
    d2h = N_CHUNKS * 16 + 32

    def e2e_step():
        # one C-ABI call: H2D of 32 MiB voxels + positions in sub-batches, each meshed as soon as it has landed, then the
        # entry table D2H; blocking
        ctx.mesh_host_chunks_ptr(N_CHUNKS, pos, host_vox.data_ptr(), entries_np, mode)

    for _ in range(3):
        e2e_step()
    # PCIe floor for context: the same 32 MiB pinned host buffer copied to the device, nothing else
    dev_tmp = torch.empty((N_CHUNKS, 4096), dtype=torch.int16, device="cuda")
    dev_tmp.copy_(host_vox, non_blocking=True)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(5):
        dev_tmp.copy_(host_vox, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_only_ms = c0.elapsed_time(c1) / 5
    del dev_tmp
    barrier()
    torch.cuda.synchronize()
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s_ev.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    e_ev.record(stream)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    barrier()
    assert int(entries_np["vertex_count"].astype(np.int64).sum()) == faces * 6  # the result really came back
    e2e_ms = max_over_ranks(max(s_ev.elapsed_time(e_ev), (t1 - t0) * 1e3))
    e2e_value = world * N_CHUNKS * e2e_steps / (e2e_ms * 1e-3)

    # ---- full read-back variant (vertices copied to pinned host memory as well), reported separately -----
    rb_host = torch.empty(total_bytes, dtype=torch.uint8).pin_memory()
    ctx.download_arena_into(0, total_bytes, rb_host.data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rb_steps = max(2, min(5, e2e_steps))
    for _ in range(rb_steps):
        e2e_step()
        ctx.download_arena_into(0, total_bytes, rb_host.data_ptr())
    torch.cuda.synchronize()
    rb_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    rb_value = world * N_CHUNKS * rb_steps / (rb_ms * 1e-3)

    # ---- the reference's own call pattern, end to end: chunk POSITIONS in (World::load_chunk generates, scene_game.rs:56-64),
    #      terrain fill + mesh on the device, entry table out.  Reported beside `e2e`, which carries host voxels.
    gp_steps = max(3, min(20, e2e_steps))
    slots_region = np.arange(first, first + N_CHUNKS, dtype=np.uint32)

    def positions_step():
        ctx.regenerate_chunks(pos, capi.GEN_NOISE, SEED)    # 48 KiB of positions H2D, voxels never exist on the host (the generator runs every step)
        ctx.mesh_chunks(slots_region, mode)                  # blocking: entry table D2H

    for _ in range(2):
        positions_step()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(gp_steps):
        positions_step()
    torch.cuda.synchronize()
    gp_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    gp_value = world * N_CHUNKS * gp_steps / (gp_ms * 1e-3)

    # ---- extension modes on the same region (kernel time from the library's own events; L2 not flushed here) ----------
    modes_out = {}
    for mname, m in (("indexed", capi.MESH_INDEXED), ("greedy", capi.MESH_GREEDY), ("greedy_indexed", capi.MESH_GREEDY | capi.MESH_INDEXED)):
        e_m, tot_m = ctx.mesh_range(first, N_CHUNKS, m)
        quads = int(e_m["vertex_count"].astype(np.int64).sum()) // (4 if m & capi.MESH_INDEXED else 6)
        ctx.sync()
        k0, t0ms, _ = ctx.mesh_kernel_stats()
        for _ in range(10):
            ctx.mesh_range_async(first, N_CHUNKS, m)
        ctx.sync()
        k1, t1ms, _ = ctx.mesh_kernel_stats()
        kms = (t1ms - t0ms) / max(1, k1 - k0)
        ab = N_CHUNKS * (8192 + 32) + (104 if m & capi.MESH_INDEXED else 120) * quads
        modes_out[mname] = {"quads": quads, "arena_bytes": int(tot_m), "kernel_ms": kms, "chunks_per_s_per_gpu": N_CHUNKS / (kms * 1e-3),
                            "algorithmic_GBps": ab / (kms * 1e-3) / 1e9}

    # ---- CPU baseline beside it (rank 0, N = 1 only) --------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import loader as orc
        threads = orc.hardware_threads()
        w = orc.Workload(pos, orc.GEN_NOISE, SEED, UV)
        w.mesh(orc.FAITHFUL, threads)
        secs, reps, t_start = 0.0, 0, time.perf_counter()
        while time.perf_counter() - t_start < 10.0 and reps < 200:
            s, _ = w.mesh(orc.FAITHFUL, threads)
            secs += s
            reps += 1
        one, _ = w.mesh(orc.FAITHFUL, 1)
        fair, _ = w.mesh(orc.FAIR, threads)
        cpu = {"value": N_CHUNKS * reps / secs, "unit": "chunks/s", "cores": threads, "kind": "port",
               "sample": f"build_chunk_mesh restatement (faithful flavour) over all {N_CHUNKS} chunks of the workload, {reps} passes (~10 s), host-resident chunks",
               "one_thread_chunks_per_s": N_CHUNKS / one, "fair_flavour_chunks_per_s": N_CHUNKS / fair,
               "note": "C++ restatement of the reference mesher; the Rust reference cannot be built in this image (no rustc)"}

    faces_all = sum_over_ranks(float(faces))
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16 voxels -> f32 vertices", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chunks_per_step_per_gpu": N_CHUNKS, "faces_per_step": int(faces_all),
                       "vertex_bytes_per_step_per_gpu": int(total_bytes), "seed": hex(SEED), "slot_order": "call order (DSP_MESH_ORDERED)" if args.ordered else "completion order", "parallelism": f"region-per-gpu x{world}",
                       "l2": "not flushed: the steps rotate over 4 identical copies of the region and 4 arena slices, so a step's 32 MiB of voxels and 397 MB of "
                             "vertices were last touched 1.7 GB of traffic earlier (L2 is 126 MB)",
                       "timing": "one CUDA event pair on the context's stream around all K steps (each step = one kernel launch, back to back); max over ranks"},
            "voxels_per_s": value * 4096,
            "faces_per_s": world * faces * args.steps / (dev_ms_total * 1e-3),
            "gpu_launches": int(launches),
            "wall_s_timed_region": wall1 - wall0,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if args.ordered else NCU_TRAFFIC_BYTES, "traffic_source": NCU_TRAFFIC_SOURCE,
                         "kernel": "mesh_chunks_kernel (look-back, DSP_MESH_ORDERED)" if args.ordered else "mesh_chunks_bump_kernel", "kernel_ms_avg": kernel_ms_avg,
                         "algorithmic_bytes_per_launch": int(algo_bytes), "peak_source": peak_src,
                         "frac_of_8TBs_datasheet": achieved / 8000.0},
            "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps, "h2d_copy_alone_ms": h2d_only_ms,
                    "what": "dsp_mesh_host_chunks(host u16 voxels, pinned; copy/mesh pipelined in 512-chunk sub-batches) -> dsp_mesh_entry table on the host; vertices stay in HBM "
                            "(the reference's result is a device buffer too, chunk_mesh.rs:111-124)"},
            "e2e_vertex_readback": {"value": rb_value, "unit": "chunks/s", "d2h_bytes_per_step": int(total_bytes) + d2h, "steps": rb_steps,
                                    "what": "same, plus the whole vertex stream copied to pinned host memory"},
            "e2e_from_positions": {"value": gp_value, "unit": "chunks/s", "h2d_bytes_per_step": N_CHUNKS * 16, "d2h_bytes_per_step": d2h, "steps": gp_steps,
                                   "what": "dsp_generate_chunks(host positions) + dsp_mesh_chunks -> entry table on the host: the reference's load loop "
                                           "(World::load_chunk generates the chunk, scene_game.rs:56-64), voxels never cross PCIe"},
            "modes_per_gpu": modes_out,
            "clocks": sampler.summary(note),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    ctx.close()
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
