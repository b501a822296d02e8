"""Builds libdespawn_b200.so in-tree with nvcc for sm_100a (B200).  No JIT cache, no torch arch list:
the .so sits next to the sources so that it travels with the repo snapshot to the GPU box.

    python -m despawnengine_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdespawn_b200.so")
SOURCES = [os.path.join(CSRC, "dsp_api.cu"), os.path.join(CSRC, "host_pack.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("dsp_common.cuh", "mesh_kernels.cuh", "quad_kernels.cuh", "terrain_kernels.cuh", "dsp_scene.inl", "dsp_seam.inl", "dsp_export.inl", "host_pack.hpp")] + [
    os.path.join(ROOT, "include", "despawn_b200.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",  # the reference is rustc output: one rounding per f32 op, never a fused multiply-add
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
