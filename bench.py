#!/usr/bin/env python
"""bench.py -- chunks meshed/s on BASELINE.json's configs[1] (16x16x16-chunk region of seeded noise
terrain, 4,096 chunks = 256^3 voxels per GPU), device-resident (`value`), end to end from host buffers
through the C ABI (`e2e`), against the kernel's HBM roofline (`roofline`) and with the CPU
restatement of the reference timed beside it (`cpu_baseline`).  The other BASELINE.json configs ride along as
sub-records under `configs` (cfg3a/cfg3b worst-case density, cfg4 incremental remesh, cfg5 the 512x512x32 world split
over the ranks with the seam exchange), each with a sampled check against the oracle at its full size.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--skip-configs]

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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE mesh_chunks_bump_kernel launch of this exact workload, from the round-2
# `ncu --set full` capture summarised in profiles/r02_kernels_ncu_summary.csv (launch1: 33.66 MB read + 340.96 MB written; the
# remaining ~56 MB of the 397 MB vertex stream was still dirty in the 126 MB L2 when the kernel ended).  Below the algorithmic
# 430.5 MB, i.e. no wasted re-reads.
NCU_TRAFFIC_BYTES = 10_231_040 + 338_869_248
NCU_TRAFFIC_SOURCE = "profiles/r02b_kernels_ncu_summary.csv (ncu --set full, tools/prof_r02b.py, launch1 of mesh_chunks_bump_kernel, same workload; the rest of the 396.8 MB vertex stream is still dirty in L2 at kernel exit)"


def workload_config(n_gpus: int) -> dict:
    """What defines the workload -- identical, key for key and value for value, in both arms (`--impl reference` and this one)."""
    return {"workload": WORKLOAD, "region_dims_chunks": list(DIMS), "chunks_per_step_per_gpu": N_CHUNKS, "voxels_per_chunk": 4096,
            "terrain": "integer value-noise heightmap, 3 octaves, ids 1 (body) / 2 (surface) (DESIGN.md section 5)", "seed": hex(SEED),
            "mesh_mode": "reference semantics: chunk-local cull, one unit quad of 6 vertices (120 B) per visible face, stream order of chunk_mesh.rs",
            "uv_table": "template:dirt -> [0,0]-[0.5,1], template:engine -> [0.5,0]-[1,1] (SURVEY.md 8d config 1)",
            "parallelism": f"region-per-gpu x{n_gpus}"}


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

    def __init__(self, index: int, period_s: float = float(os.environ.get("DSP_BENCH_SAMPLE_PERIOD", "0.02"))):
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
        "config": workload_config(args.gpus),
        "vertices_per_step": verts,
        "note": "C++ restatement of build_chunk_mesh (faithful flavour: per-voxel string-keyed lookups, unreserved Vec) over host-resident chunks + one "
                "memcpy per chunk for the Vulkan upload, on all host threads of rank 0; the Rust reference itself cannot be built here (no rustc)",
        "voxels_per_s": value * 4096,
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": threads, "kind": "port",
                         "sample": f"all {N_CHUNKS} chunks of the workload per step, {steps_done} steps"},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ======================================================================================================================
# side configs (BASELINE.json configs[2..4]); every record carries a sampled check against the oracle at the config's full size
# ======================================================================================================================
def _events(torch, stream):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def config3(torch, capi, device, peak):
    """configs[2]: worst-case density, 4,096 chunks on one GPU: (a) 3-D checkerboard = 12,288 faces per chunk, the maximum;
    (b) Bernoulli(0.5) from a 32-bit hash.  Output per launch is 6.0 / 3.2 GB (>> L2), so launches are timed back to back."""
    from oracle import loader as orc
    out = {}
    for key, name, kind, seed in (("cfg3a", "3-D checkerboard (12,288 faces/chunk, the per-chunk maximum)", capi.GEN_CHECKER, 0),
                                  ("cfg3b", "Bernoulli(0.5) density from hash3(seed, wx, wy, wz)", capi.GEN_RANDOM, 0xD5E50003)):
        with capi.Context(device, N_CHUNKS, 6 << 30) as ctx:
            ctx.set_uv_table(UV)
            stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", device))
            first = ctx.generate_region((0, 0, 0), DIMS, kind, seed)
            entries, total = ctx.mesh_range(first, N_CHUNKS)
            faces = int(entries["vertex_count"].astype(np.int64).sum()) // 6
            # sampled check at full size: 24 chunks spread over the region, byte for byte against the oracle
            pos = region_positions((0, 0, 0))
            sample = np.linspace(0, N_CHUNKS - 1, 24).astype(np.int64)
            for i in sample:
                ids = orc.generate(kind, seed, pos[i])
                assert np.array_equal(ctx.download_chunk(first + int(i))[0], ids), (key, "terrain", int(i))
                assert ctx.download_vertices(entries[i]).tobytes() == orc.mesh_chunk(pos[i], ids, UV, flavour=orc.FAIR).tobytes(), (key, "mesh", int(i))
            ctx.set_kernel_timing(False)
            for _ in range(3):
                ctx.mesh_range_async(first, N_CHUNKS)
            ctx.sync()
            reps = 10
            e0, e1 = _events(torch, stream)
            e0.record(stream)
            for _ in range(reps):
                ctx.mesh_range_async(first, N_CHUNKS)
            e1.record(stream)
            ctx.sync()
            ms = e0.elapsed_time(e1) / reps
            algo = N_CHUNKS * (8192 + 32) + 120 * faces
            out[key] = {"workload": f"configs[2]: {name}, 4,096 chunks (16x16x16), reference cull semantics", "faces": faces, "vertex_bytes": int(total),
                        "kernel_ms": ms, "chunks_per_s": N_CHUNKS / (ms * 1e-3), "algorithmic_bytes": int(algo), "algorithmic_GBps": algo / (ms * 1e-3) / 1e9,
                        "frac_of_measured_peak": algo / (ms * 1e-3) / 1e9 / peak, "launches_timed": reps,
                        "verified": f"{len(sample)} chunks spread over the region: voxels and vertex streams byte-identical to the oracle"}
    return out


def config4(torch, capi, device, frames=20):
    """configs[3]: 10,000 random block place/destroy edits per frame (SplitMix64, seed 0xD5E50004) over a 64x64x16-chunk world
    (65,536 chunks, 512 MiB of voxels resident), one dsp_apply_edits_and_remesh call per frame: addressing, last-edit-wins,
    voxel writes, dirty list and the whole-chunk remesh of every dirty chunk (the reference's only granularity)."""
    from despawnengine_b200 import synthetic as syn
    from oracle import loader as orc
    dims = (64, 16, 64)
    n = dims[0] * dims[1] * dims[2]
    with capi.Context(device, n, 4 << 30) as ctx:
        ctx.set_uv_table(UV)
        first = ctx.generate_region((0, 0, 0), dims, capi.GEN_NOISE, SEED)
        ctx.mesh_range(first, 1024)
        state, all_edits, frame_ms, dirty_n = syn.SEED_EDITS, [], [], []
        dirty = entries = None
        for frame in range(frames + 2):
            state, edits = syn.edit_frame(state, 10000, dims)
            all_edits.append(edits)
            ctx.sync()
            t0 = time.perf_counter()
            dirty, entries, total = ctx.apply_edits_and_remesh(edits)
            t1 = time.perf_counter()
            if frame >= 2:
                frame_ms.append((t1 - t0) * 1e3)
                dirty_n.append(len(dirty))
        # check: a sample of the LAST frame's dirty chunks against the oracle on the voxels after ALL frames' edits
        hist = np.concatenate(all_edits)
        pick = np.linspace(0, len(dirty) - 1, 32).astype(np.int64)
        for k in pick:
            slot = int(dirty[k]) - first
            p = (slot % dims[0], (slot // dims[0]) % dims[1], slot // (dims[0] * dims[1]))
            ids = syn.replay_edits_on_chunk(orc.generate(orc.GEN_NOISE, SEED, p), p, hist)
            assert np.array_equal(ctx.download_chunk(int(dirty[k]))[0], ids), ("cfg4 voxels", p)
            assert ctx.download_vertices(entries[k]).tobytes() == orc.mesh_chunk(p, ids, UV, flavour=orc.FAIR).tobytes(), ("cfg4 mesh", p)
        ms = float(np.median(frame_ms))
        return {"workload": "configs[3]: 10,000 random place/destroy edits per frame (SplitMix64 seed 0xD5E50004) over a 64x64x16-chunk world (65,536 chunks)",
                "frames": frames, "frame_ms": ms, "frame_ms_best": float(np.min(frame_ms)), "edits_per_s": 10000 / (ms * 1e-3),
                "dirty_chunks_per_frame": float(np.mean(dirty_n)), "dirty_chunks_per_s": float(np.mean(dirty_n)) / (ms * 1e-3),
                "timing": "host wall clock around the blocking dsp_apply_edits_and_remesh call: edits H2D, device edit pipeline, mesh of the dirty list, "
                          "dirty list + entry table D2H",
                "verified": f"{len(pick)} dirty chunks of the last frame: voxels (all {frames + 2} frames of edits replayed in order) and vertex streams byte-identical to the oracle"}


def config5(torch, dist, capi, rank, world, local_rank, peak):
    """configs[4]: the 512x512x32-chunk world (8,388,608 chunks, 64 GiB of voxels) cut into one X-slab per GPU (strong scaling),
    generated on the devices, meshed (a) with the reference's chunk-local cull, vertices streamed through a reused 8 GiB arena,
    and (b) with cross-chunk culling (DSP_MESH_HALO) where region seams are exchanged inside the library (dsp_seam_*: peer stores
    over NVLink, the wait for them inside the mesh kernel behind the interior chunks)."""
    from despawnengine_b200 import partition as part
    from despawnengine_b200.halo import RegionOnDevice, connect_region_seams
    from oracle import loader as orc
    W = part.Region((0, 0, 0), (512, 32, 512))
    sp = part.SlabPartition(W, world)
    reg = sp.region(rank)
    n = reg.n_chunks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def red(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    with capi.Context(local_rank, n, 8 << 30) as ctx:
        ctx.set_uv_table(UV)
        barrier()
        t0 = time.perf_counter()
        dev = RegionOnDevice(ctx, reg, capi.GEN_NOISE, SEED)
        ctx.sync()
        barrier()
        gen_s = red(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
        if world > 1:
            connect_region_seams(dev, sp, rank, dist)
        res = {}
        # (chunks per launch sized for ~5 GB of the 8 GiB arena: parity 98,304 chunks at ~50 KB each, halo 786,432 at ~6 KB each; every launch
        #  is followed by a host sync that reads its total)
        for mode_name, mode, batch in (("parity", capi.MESH_PARITY, 98304), ("halo", capi.MESH_HALO, 786432)):
            xchg_ms = 0.0
            if mode == capi.MESH_HALO and world > 1:
                ctx.seam_exchange_async(); ctx.mesh_range(dev.first_slot, min(n, 4096), mode)  # epoch 1: warm-up
                barrier()
                t0 = time.perf_counter()
                ctx.seam_exchange_async()  # an exchange alone, for the record: enqueue -> every rank's pushes have landed
                ctx.sync()
                barrier()
                xchg_ms = red(time.perf_counter() - t0, dist.ReduceOp.MAX) * 1e3
            ctx.mesh_range(dev.first_slot, min(n, batch), mode)  # warm-up (builds the neighbour table in halo mode)
            vbytes = 0
            barrier()
            t0 = time.perf_counter()
            if mode == capi.MESH_HALO and world > 1:
                ctx.seam_exchange_async()  # this pass's border layers travel while the interior is meshed
            for b0 in range(0, n, batch):
                m = min(batch, n - b0)
                ctx.mesh_range_async(dev.first_slot + b0, m, mode)
                _, total = ctx.mesh_result(m, entries_ptr=0)  # syncs; raises on arena overflow or a seam time-out
                vbytes += total
            barrier()
            secs = red(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
            vb_all = red(float(vbytes), dist.ReduceOp.SUM if world > 1 else None)
            res[mode_name] = {"ms": secs * 1e3, "chunks_per_s": W.n_chunks / secs, "vertex_bytes": vb_all,
                              "algorithmic_GBps_per_gpu": (vb_all / world + n * 8224) / secs / 1e9,
                              "frac_of_measured_peak": (vb_all / world + n * 8224) / secs / 1e9 / peak, "chunks_per_launch": batch}
            if mode == capi.MESH_HALO:
                res[mode_name].update({"halo_exchange_ms_alone": xchg_ms, "halo_bytes_sent_per_seam_direction": 512 * 32 * 512 if world > 1 else 0,
                                       "seam_path": "dsp_seam_* (inside the library, overlapped with the interior of the region)" if world > 1 else None})
        # sampled check at full dimensions, in halo mode: seam chunks (N > 1) and chunks spread over the slab, neighbours generated on the CPU
        DIRS = [(0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (-1, 0, 0), (1, 0, 0)]
        sample = []
        for t in sp.seam_transfers(rank):
            sample += [tuple(p) for p in t.positions[:: max(1, len(t.positions) // 8)].tolist()]
        allp = reg.positions()
        surf = allp[(allp[:, 1] >= 1) & (allp[:, 1] <= 14)]  # the noise surface lives in chunk Y 1..14
        sample += [tuple(p) for p in surf[:: max(1, len(surf) // 16)].tolist()]
        slots = dev.slots_of(np.array(sample, dtype=np.int32))
        e_h, _ = ctx.mesh_chunks(slots, capi.MESH_HALO)
        e_p, _ = ctx.mesh_chunks(slots, capi.MESH_PARITY, arena_offset=1 << 30)
        for p, eh, ep in zip(sample, e_h, e_p):
            ids = orc.generate(orc.GEN_NOISE, SEED, p)
            slabs, present = np.zeros((6, 256), dtype=np.uint16), np.zeros(6, dtype=np.uint8)
            for f, d in enumerate(DIRS):
                q = (p[0] + d[0], p[1] + d[1], p[2] + d[2])
                if all(W.origin[i] <= q[i] < W.origin[i] + W.dims[i] for i in range(3)):
                    v = orc.generate(orc.GEN_NOISE, SEED, q).reshape(16, 16, 16)
                    slabs[f] = [v[:, 0, :], v[:, 15, :], v[15, :, :], v[0, :, :], v[:, :, 15], v[:, :, 0]][f].reshape(256)
                    present[f] = 1
            assert ctx.download_vertices(eh).tobytes() == orc.mesh_chunk_halo(p, ids, slabs, present, UV).tobytes(), ("cfg5 halo", rank, p)
            assert ctx.download_vertices(ep).tobytes() == orc.mesh_chunk(p, ids, UV, flavour=orc.FAIR).tobytes(), ("cfg5 parity", rank, p)
        checked = red(float(len(sample)), dist.ReduceOp.SUM if world > 1 else None)
        return {"workload": "configs[4]: 512x512x32-chunk world (x, z, y-up; 8,388,608 chunks, 64 GiB of voxels), X-slab per GPU, generated on the devices",
                "n_gpus": world, "chunks": W.n_chunks, "chunks_per_gpu": n, "scaling": "strong", "terrain_fill_ms": gen_s * 1e3,
                "terrain_fill_GBps_per_gpu": n * 8192 / gen_s / 1e9, "parity": res["parity"], "halo": res["halo"],
                "timing": "wall clock between barriers, max over ranks; one host sync per launch batch",
                "verified": f"{int(checked)} chunks at full world dimensions (seam chunks of every rank + chunks spread over the slabs): halo-mode streams byte-identical "
                            "to the extended oracle with CPU-generated neighbours, parity-mode streams to the reference restatement"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end legs (default min(steps, 40))")
    ap.add_argument("--sustain-s", type=float, default=1.2, help="length of the continuous run of the rotating step that contains the K timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="headline only: skip the cfg3a/cfg3b/cfg4/cfg5 sub-records")
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
    # Host side of the end-to-end legs: keep this rank's thread -- and with it the pinned buffers it allocates below -- on the CPUs
    # (NUMA node) the GPU hangs off, so that eight ranks do not all pull their uploads across one socket's memory.
    affinity = "not set"
    if os.environ.get("DSP_BENCH_NO_AFFINITY") is None:
        try:
            import pynvml
            pynvml.nvmlInit()
            # the CPU set of every GPU of this job; ranks whose GPUs share a set (one NUMA node) take equal slices of it, so that each
            # rank's host threads -- the library's upload packers are as many as the CPUs the calling thread may run on, minus one --
            # have cores of their own
            sets = []
            for g in range(world):
                pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(g))
                sets.append(tuple(sorted(os.sched_getaffinity(0))))
            mine = sets[local_rank]
            sharers = [g for g in range(world) if sets[g] == mine]
            k, idx = len(mine) // len(sharers), sharers.index(local_rank)
            os.sched_setaffinity(0, set(mine[idx * k:(idx + 1) * k]) if k >= 1 else set(mine))
            affinity = f"{len(os.sched_getaffinity(0))} of the {len(mine)} CPUs local to GPU {local_rank} (nvmlDeviceSetCpuAffinity; {len(sharers)} rank(s) share them)"
        except Exception as e:  # pragma: no cover
            affinity = f"not set ({type(e).__name__})"
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

    peak, peak_src = load_peaks()
    # COPIES copies of the region with identical voxels (dsp_clone_region: same faces, same bytes, other chunk slots and other
    # places of the world) and as many arena slices: the steps rotate over them, so that a step's input (32 MiB of voxels) and
    # output (397 MB of vertices) were last touched COPIES steps earlier -- 1.7 GB of traffic ago, against 126 MB of L2.  No flush
    # is needed then and the steps run back to back on the stream (launch latency overlaps the previous step, as in a frame loop).
    COPIES, SLICE = 4, 448 << 20
    ctx = capi.Context(local_rank, N_CHUNKS * COPIES, COPIES * SLICE)
    ctx.set_uv_table(UV)
    origin = (DIMS[0] * rank, 0, 0)
    first = ctx.generate_region(origin, DIMS, capi.GEN_NOISE, SEED)
    firsts = [first] + [ctx.clone_region(first, (origin[0], 4096 * k, 0)) for k in range(1, COPIES)]
    mode = capi.MESH_PARITY
    entries, total_bytes = ctx.mesh_range(first, N_CHUNKS, mode)
    assert total_bytes <= SLICE
    faces = int(entries["vertex_count"].astype(np.int64).sum()) // 6
    # algorithmic bytes per launch: voxels read once + vertices written once + position + entry
    algo_bytes = N_CHUNKS * (8192 + 16 + 16) + 120 * faces

    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def step(k: int):
        ctx.mesh_range_async(firsts[k % COPIES], N_CHUNKS, mode, (k % COPIES) * SLICE)

    # ---- device-resident leg: `value` -----------------------------------------------------------------
    # A step is exactly one launch of the mesh kernel (it resets its own control block: no memset in front).  The K timed steps
    # sit INSIDE a continuous run of the same rotating step that lasts >= --sustain-s seconds: [lead-in][ev0 K steps ev1][tail],
    # all enqueued back to back, one event pair around the K steps and one around the whole run.  NVML is sampled every 20 ms
    # over the whole run, so the clocks / power / throttle reasons reported are those of the load the K steps ran under.
    ctx.set_kernel_timing(False)
    for k in range(max(args.warmup, COPIES)):
        step(k)
    ctx.sync()
    q0, q1 = _events(torch, stream)
    q0.record(stream)
    for k in range(64):
        step(k)
    q1.record(stream)
    ctx.sync()
    est_ms = q0.elapsed_time(q1) / 64
    n_side = max(COPIES, int(np.ceil(0.5 * args.sustain_s * 1e3 / est_ms)))
    launches0 = ctx.kernel_launches
    ev_a, ev_z = _events(torch, stream)
    ev0, ev1 = _events(torch, stream)
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    wall0 = time.perf_counter()
    k = 0
    ev_a.record(stream)
    for _ in range(n_side):
        step(k); k += 1
    ev0.record(stream)
    for _ in range(args.steps):
        step(k); k += 1
    ev1.record(stream)
    for _ in range(n_side):
        step(k); k += 1
    ev_z.record(stream)
    torch.cuda.synchronize()
    wall1 = time.perf_counter()
    sampler.stop_flag.set()
    sampler.join()
    barrier()
    launches_total = ctx.kernel_launches - launches0
    assert launches_total == args.steps + 2 * n_side, (launches_total, args.steps, n_side)  # one launch per step, nothing else of ours on the stream
    timed_ms = ev0.elapsed_time(ev1)
    sustained_ms = ev_a.elapsed_time(ev_z)
    dev_ms_total = max_over_ranks(timed_ms)
    value = world * N_CHUNKS * args.steps / (dev_ms_total * 1e-3)
    kernel_ms_avg = timed_ms / args.steps  # this rank; `value` uses the max over ranks
    achieved = algo_bytes / (kernel_ms_avg * 1e-3) / 1e9
    sus_total = max_over_ranks(sustained_ms)
    sus_steps = args.steps + 2 * n_side
    sustained = {"seconds": sus_total * 1e-3, "steps": sus_steps, "ms_per_step": sus_total / sus_steps,
                 "chunks_per_s": world * N_CHUNKS * sus_steps / (sus_total * 1e-3),
                 "algorithmic_GBps_per_gpu": algo_bytes / (sustained_ms / sus_steps * 1e-3) / 1e9,
                 "frac_of_measured_peak": algo_bytes / (sustained_ms / sus_steps * 1e-3) / 1e9 / peak,
                 "what": "the whole continuous run of the rotating step (lead-in + the K timed steps + tail), one event pair, max over ranks"}
    # burst figure beside it: the same K steps after the GPU has been idle (what round 1 reported)
    time.sleep(0.3)
    b0, b1 = _events(torch, stream)
    barrier()
    b0.record(stream)
    for kk in range(args.steps):
        step(kk)
    b1.record(stream)
    ctx.sync()
    burst_ms = max_over_ranks(b0.elapsed_time(b1))
    burst = {"ms_per_step": burst_ms / args.steps, "chunks_per_s": world * N_CHUNKS * args.steps / (burst_ms * 1e-3),
             "what": "the K steps alone, enqueued on an idle GPU (clocks not yet settled under load)"}
    # per-step figures as SURVEY.md 8(d) words them: every step bracketed by its own CUDA event pair, median and best of 30
    # (an event between two launches keeps the next step's CTAs from taking their places early, so these sit a little above `value`)
    pairs = [_events(torch, stream) for _ in range(30)]
    barrier()
    for kk, (ea, eb) in enumerate(pairs):
        ea.record(stream)
        step(kk)
        eb.record(stream)
    ctx.sync()
    per = sorted(ea.elapsed_time(eb) for ea, eb in pairs)
    per_step = {"median_ms": max_over_ranks(per[len(per) // 2]), "best_ms": max_over_ranks(per[0]), "steps": len(per),
                "what": "each of 30 steps bracketed by its own CUDA event pair on the context's stream (max over ranks of the per-rank median / best)"}
    ctx.set_kernel_timing(True)

    # ---- other kernels on the same region (kernel time from the library's own events, 12 calls back to back;
    #      before the end-to-end legs, whose uploads reset the chunk summaries of the first copy) ----------
    modes_out = {}
    for mname, m in (("ordered", capi.MESH_ORDERED), ("indexed", capi.MESH_INDEXED), ("greedy", capi.MESH_GREEDY), ("greedy_indexed", capi.MESH_GREEDY | capi.MESH_INDEXED)):
        e_m, tot_m = ctx.mesh_range(first, N_CHUNKS, m)
        per = 4 if m & capi.MESH_INDEXED else 6
        quads = int(e_m["vertex_count"].astype(np.int64).sum()) // per
        ctx.sync()
        k0, t0ms, _ = ctx.mesh_kernel_stats()
        for j in range(12):
            ctx.mesh_range_async(firsts[j % COPIES], N_CHUNKS, m, (j % COPIES) * SLICE)
        ctx.sync()
        k1, t1ms, _ = ctx.mesh_kernel_stats()
        kms = (t1ms - t0ms) / max(1, k1 - k0)
        ab = N_CHUNKS * (8192 + 32) + (104 if m & capi.MESH_INDEXED else 120) * quads
        modes_out[mname] = {"quads": quads, "arena_bytes": int(tot_m), "kernel_ms": kms, "chunks_per_s_per_gpu": N_CHUNKS / (kms * 1e-3),
                            "algorithmic_GBps": ab / (kms * 1e-3) / 1e9, "frac_of_measured_peak": ab / (kms * 1e-3) / 1e9 / peak}

    # ---- end-to-end legs: host voxel buffers -> C ABI -> results on the host ---------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 40)
    pos = region_positions(origin)
    host_vox = torch.empty((N_CHUNKS, 4096), dtype=torch.int16).pin_memory()
    hv = host_vox.numpy().view(np.uint16)
    for i in range(N_CHUNKS):  # setup, untimed: host copies of the chunks, as a host caller would hold them (Chunk.blocks)
        hv[i] = ctx.download_chunk(first + i)[0]
    # chunks whose tile the parity kernel does not read: all air (ended at the ticket) / one solid block id (rebuilt from the id), by the generator's chunk summaries
    n_air = int((hv == 0).all(axis=1).sum())
    n_uniform = int(((hv == hv[:, :1]).all(axis=1) & (hv[:, 0] != 0)).sum())
    # Chunk::palette.len() of every chunk as the reference's World would hold it (chunk.rs:11,58-65): 1 for a chunk that has never held
    # anything but air, else air + the block ids present.  Part of the host-side chunk, like `blocks`; the CPU arm's restatement uses it too
    # (its build_chunk_mesh returns None on palette.len() == 1 without looking at a voxel, chunk_mesh.rs:18-20).
    palette_lens = np.array([1 + len(np.setdiff1d(np.unique(hv[i]), [0])) for i in range(N_CHUNKS)], dtype=np.uint32)
    assert int((palette_lens == 1).sum()) == n_air
    entries_host = torch.empty(N_CHUNKS * 16, dtype=torch.uint8).pin_memory()
    entries_np = entries_host.numpy().view(capi.MESH_ENTRY_DTYPE)
    h2d = N_CHUNKS * (8192 + 16)                        # every tile + position (calls without palette lengths)
    h2d_pal = (N_CHUNKS - n_air) * 8192 + N_CHUNKS * 21   # tiles of the chunks that hold something + position, slot, flag per chunk
    d2h_entries = N_CHUNKS * 16 + 32
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

    def timed_calls(fn, steps):
        for _ in range(3):
            fn()
        barrier()
        torch.cuda.synchronize()
        s_ev, e_ev = _events(torch, stream)
        t0 = time.perf_counter()
        s_ev.record(stream)
        for _ in range(steps):
            fn()
        e_ev.record(stream)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        ms = max_over_ranks(max(s_ev.elapsed_time(e_ev), (t1 - t0) * 1e3))
        return world * N_CHUNKS * steps / (ms * 1e-3), ms / steps

    # (1) `e2e`: host chunks (voxels + palette lengths) in, entry table out, vertices stay in HBM (the reference's result is a device
    #     buffer too).  dsp_mesh_host_chunks_ex, packed upload: host worker threads reduce every chunk to its smallest lossless form
    #     (nothing for air / one-id chunks, 2 bits per voxel for most of the rest), a kernel expands them on the device, the mesh kernel follows.
    e2e_value, e2e_ms = timed_calls(lambda: ctx.mesh_host_chunks_ex_ptr(N_CHUNKS, pos, host_vox.data_ptr(), palette_lens, entries_np, mode), e2e_steps)
    assert int(entries_np["vertex_count"].astype(np.int64).sum()) == faces * 6  # the result really came back
    up = ctx.host_upload_stats()
    packed = sum(up[k] for k in capi.PACK_FORMS) == N_CHUNKS  # (not with DSP_E2E_PACK=0 or on a one-CPU box: then this leg is the kernel pull)
    assert not packed or up["air"] == n_air, up  # all-air chunks by their palette length
    # (1b) the same chunks without the palette lengths (dsp_mesh_host_chunks): the packer reads the all-air chunks too
    nopal_value, nopal_ms = timed_calls(lambda: ctx.mesh_host_chunks_ptr(N_CHUNKS, pos, host_vox.data_ptr(), entries_np, mode), e2e_steps)
    assert int(entries_np["vertex_count"].astype(np.int64).sum()) == faces * 6
    up_nopal = ctx.host_upload_stats()
    # (1c) the former paths, for comparison (a second context created under DSP_E2E_PACK=0): kernel pull with palette lengths (the
    #      previous `e2e`), copy-engine pipeline without them
    os.environ["DSP_E2E_PACK"] = "0"
    ctx2 = capi.Context(local_rank, N_CHUNKS, SLICE)
    os.environ.pop("DSP_E2E_PACK")
    ctx2.set_uv_table(UV)
    ctx2.set_kernel_timing(False)
    pull_value, pull_ms = timed_calls(lambda: ctx2.mesh_host_chunks_ex_ptr(N_CHUNKS, pos, host_vox.data_ptr(), palette_lens, entries_np, mode), e2e_steps)
    assert int(entries_np["vertex_count"].astype(np.int64).sum()) == faces * 6 and sum(ctx2.host_upload_stats().values()) == 0
    ce_value, ce_ms = timed_calls(lambda: ctx2.mesh_host_chunks_ptr(N_CHUNKS, pos, host_vox.data_ptr(), entries_np, mode), e2e_steps)
    ctx2.close()
    # (2) vertices delivered to pinned HOST memory (a consumer that is not on this GPU), reference layout: PCIe D2H bound
    host_verts = torch.empty(SLICE, dtype=torch.uint8).pin_memory()
    rb_steps = max(3, min(10, e2e_steps))
    rb_value, rb_ms = timed_calls(lambda: ctx.mesh_host_chunks_to_host_ptr(N_CHUNKS, pos, host_vox.data_ptr(), entries_np, host_verts.data_ptr(), SLICE, mode), rb_steps)
    probe = entries_np[N_CHUNKS // 2 + 7]
    assert host_verts.numpy()[int(probe["offset_bytes"]):int(probe["offset_bytes"]) + 20 * int(probe["vertex_count"])].tobytes() == ctx.download_vertices(probe).tobytes()
    # (3) the same with greedy merge + indexed layout: ~15x fewer bytes cross PCIe on the way out
    gi = capi.MESH_GREEDY | capi.MESH_INDEXED
    gi_value, gi_ms = timed_calls(lambda: ctx.mesh_host_chunks_to_host_ptr(N_CHUNKS, pos, host_vox.data_ptr(), entries_np, host_verts.data_ptr(), SLICE, gi), e2e_steps)
    gi_bytes = ctx.mesh_host_chunks_to_host_ptr(N_CHUNKS, pos, host_vox.data_ptr(), entries_np, host_verts.data_ptr(), SLICE, gi)
    gi_quads = int(entries_np["vertex_count"].astype(np.int64).sum()) // 4
    del host_verts

    # ---- the reference's own call pattern, end to end: chunk POSITIONS in (World::load_chunk generates, scene_game.rs:56-64),
    #      terrain fill + mesh on the device, entry table out.  Reported beside `e2e`, which carries host voxels.
    gp_steps = max(3, min(20, e2e_steps))
    slots_region = np.arange(first, first + N_CHUNKS, dtype=np.uint32)

    def positions_step():
        ctx.regenerate_chunks(pos, capi.GEN_NOISE, SEED)    # 48 KiB of positions H2D, voxels never exist on the host (the generator runs every step)
        ctx.mesh_chunks(slots_region, mode)                  # blocking: entry table D2H

    gp_value, gp_ms = timed_calls(positions_step, gp_steps)

    # ---- the headline step on a world WITHOUT chunk summaries (what a host-uploaded or freshly edited world looks like): every tile
    #      is read, all-air chunks cost a CTA iteration.  Same rotating step, 64 launches back to back, one event pair.
    ctx.set_kernel_timing(False)
    for f in firsts:
        ctx.invalidate_chunks(np.arange(f, f + N_CHUNKS, dtype=np.uint32))
    for k in range(COPIES):
        step(k)
    ctx.sync()
    n0, n1 = _events(torch, stream)
    barrier()
    n0.record(stream)
    for k in range(64):
        step(k)
    n1.record(stream)
    ctx.sync()
    nosum_ms = max_over_ranks(n0.elapsed_time(n1)) / 64
    no_summaries = {"ms_per_step": nosum_ms, "chunks_per_s": world * N_CHUNKS / (nosum_ms * 1e-3), "algorithmic_GBps_per_gpu": algo_bytes / (nosum_ms * 1e-3) / 1e9,
                    "frac_of_measured_peak": algo_bytes / (nosum_ms * 1e-3) / 1e9 / peak,
                    "what": "the same step after dsp_invalidate_chunks on every chunk (summaries unknown, as after a host upload or an edit): all 4,096 tiles are read"}
    ctx.set_kernel_timing(True)

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
    ctx.close()

    # ---- the other BASELINE.json configs --------------------------------------------------------------------------------------
    configs = {}
    if not args.skip_configs:
        if world == 1:  # single-GPU configs: reported by the N = 1 run
            configs.update(config3(torch, capi, local_rank, peak))
            configs["cfg4"] = config4(torch, capi, local_rank)
        configs["cfg5"] = config5(torch, dist, capi, rank, world, local_rank, peak)

    faces_all = sum_over_ranks(float(faces))
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16 voxels -> f32 vertices", "data": "synthetic",
            "config": workload_config(world),
            "method": {"l2": "not flushed: the steps rotate over 4 copies of the region with identical voxels (dsp_clone_region) and 4 arena slices, so a step's 32 MiB of "
                             "voxels and 397 MB of vertices were last touched 1.7 GB of traffic earlier (L2 is 126 MB)",
                       "timing": f"the K timed steps (one kernel launch each, back to back, one CUDA event pair on the context's stream) sit inside a continuous "
                                 f"{sus_total * 1e-3:.2f} s run of the same rotating step ({n_side} steps before, {n_side} after); NVML sampled every 20 ms over the whole run; max over ranks",
                       "slot_order": "completion order (atomic slot claim); DSP_MESH_ORDERED (count + scan pass, then the same kernel at preset offsets) is under modes_per_gpu.ordered"},
            "faces_per_step": int(faces_all), "vertex_bytes_per_step_per_gpu": int(total_bytes),
            "voxels_per_s": value * 4096,
            "faces_per_s": faces_all * args.steps / (dev_ms_total * 1e-3),
            "gpu_launches": int(args.steps),
            "gpu_launches_whole_sustained_run": int(launches_total),
            "wall_s_sustained_run": wall1 - wall0,
            "sustained": sustained, "burst": burst, "per_step_events": per_step, "no_summaries": no_summaries,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES, "traffic_source": NCU_TRAFFIC_SOURCE,
                         "kernel": "mesh_chunks_bump_kernel", "kernel_ms_avg": kernel_ms_avg,
                         "algorithmic_bytes_per_launch": int(algo_bytes), "peak_source": peak_src,
                         "bytes_moved_per_launch": int(algo_bytes - 8192 * (n_air + n_uniform)),
                         "bytes_moved_note": f"the generator's chunk summaries let the kernel skip the tile of {n_air} all-air and {n_uniform} one-solid-id chunks "
                                             f"({8192 * (n_air + n_uniform)} B of the algorithmic bytes are not read); `no_summaries` is the same step with every tile read",
                         "moved_GBps": (algo_bytes - 8192 * (n_air + n_uniform)) / (kernel_ms_avg * 1e-3) / 1e9,
                         "frac_of_8TBs_datasheet": achieved / 8000.0},
            "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": up["device_read_bytes"] if packed else h2d_pal, "d2h_bytes_per_step": d2h_entries, "steps": e2e_steps,
                    "ms_per_step": e2e_ms, "h2d_copy_alone_ms": h2d_only_ms, "host_voxel_bytes_per_step": N_CHUNKS * 8192, "host_threads": up["host_threads"],
                    "path": ("packed upload" + (f", shared with the bus: {up['raw']} chunks fetched as they are ({up['host_threads']} host worker threads per rank)" if up["raw"] else ""))
                            if packed else "kernel pull (see e2e_kernel_pull)",
                    "chunks_by_form": {k: up[k] for k in capi.PACK_FORMS},
                    "what": "dsp_mesh_host_chunks_ex(host u16 voxels in pinned memory + Chunk::palette.len() per chunk) -> dsp_mesh_entry table on the host, PACKED UPLOAD: "
                            "host worker threads (host_threads; this rank's share of the CPUs local to the GPU) reduce every chunk to the smallest lossless form that holds it -- "
                            "nothing for all-air (by palette length, chunk_mesh.rs:18-20, which the CPU arm takes too) and one-id chunks, 2 / 4 / 8 bits per voxel where the ids "
                            "fit, raw otherwise -- in pinned staging; unpack_host_chunks_kernel reads that over PCIe (h2d_bytes_per_step: what the device actually read) and "
                            "expands it into the voxel store in six pieces, ONE mesh launch started behind the first piece takes the chunks as the pieces are released; chunks are resident afterwards bit for bit; vertices stay in HBM "
                            "(the reference's result is a device buffer too, chunk_mesh.rs:111-124; dsp_arena_export_fd hands that buffer to the renderer)"},
            "e2e_without_palette_lens": {"value": nopal_value, "unit": "chunks/s", "h2d_bytes_per_step": up_nopal["device_read_bytes"] if packed else h2d, "d2h_bytes_per_step": d2h_entries,
                                         "steps": e2e_steps, "ms_per_step": nopal_ms,
                                         "what": "dsp_mesh_host_chunks: the same chunks without their palette lengths -- packed upload too, the packer reads the all-air chunks as well"},
            "e2e_kernel_pull": {"value": pull_value, "unit": "chunks/s", "h2d_bytes_per_step": h2d_pal, "d2h_bytes_per_step": d2h_entries, "steps": e2e_steps, "ms_per_step": pull_ms,
                                "what": "DSP_E2E_PACK=0, with palette lengths (the `e2e` of the earlier sessions of this round): the mesh kernel pulls each non-air tile over PCIe itself "
                                        "(1-D bulk / TMA copies from pinned host memory), no host threads"},
            "e2e_copy_engine": {"value": ce_value, "unit": "chunks/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_entries, "steps": e2e_steps, "ms_per_step": ce_ms,
                                "what": "DSP_E2E_PACK=0, without palette lengths (round 1's `e2e`): all 4,096 tiles cross PCIe, copy/mesh pipelined in ~1,024-chunk sub-batches on the copy engine"},
            "e2e_vertices_to_host": {"value": rb_value, "unit": "chunks/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(total_bytes) + d2h_entries, "steps": rb_steps, "ms_per_step": rb_ms,
                                     "what": "dsp_mesh_host_chunks_to_host, reference layout (6 vertices per face): vertices end in pinned host memory; upload, meshing and download "
                                             "overlap on three streams; bound by PCIe device-to-host (397 MB per step)"},
            "e2e_vertices_to_host_greedy_indexed": {"value": gi_value, "unit": "chunks/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(gi_bytes) + d2h_entries, "steps": e2e_steps,
                                                    "ms_per_step": gi_ms, "quads": gi_quads,
                                                    "what": "the same call with DSP_MESH_GREEDY | DSP_MESH_INDEXED (extension: merged quads, 4 vertices + 6 indices): vertices and indices "
                                                            "end in pinned host memory; back at the upload's pace"},
            "e2e_from_positions": {"value": gp_value, "unit": "chunks/s", "h2d_bytes_per_step": N_CHUNKS * 16, "d2h_bytes_per_step": d2h_entries, "steps": gp_steps, "ms_per_step": gp_ms,
                                   "what": "dsp_regenerate_chunks(host positions) + dsp_mesh_chunks -> entry table on the host: the reference's load loop "
                                           "(World::load_chunk generates the chunk, scene_game.rs:56-64), voxels never cross PCIe"},
            "modes_per_gpu": modes_out,
            "host_affinity_rank0": affinity,
            "configs": configs,
            "clocks": sampler.summary("sampled every 20 ms over the continuous run that contains the timed steps"),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
