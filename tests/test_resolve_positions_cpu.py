"""Position -> slot resolution of the batch calls (csrc/dsp_api.cu: acquire_slot, resolve_positions), on the CPU: the harness includes
the library source as one translation unit, builds contexts by hand (no CUDA call) and holds the run-continuing loop to the
entry-by-entry path and to the plain lookup.  The GPU twin through the C ABI is tests/test_gpu_slot_resolve.py."""
import os
import subprocess

from despawnengine_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "resolve_positions_harness.cu")


def test_resolve_positions_matches_entry_by_entry_lookup(tmp_path):
    exe = str(tmp_path / "resolve_positions_harness")
    flags = [f for f in build.NVCC_FLAGS if f not in ("-shared", "-lineinfo")]
    subprocess.run([build.nvcc()] + flags + ["-o", exe, SRC, os.path.join(build.CSRC, "host_pack.cpp"), "-lpthread"], check=True, capture_output=True, cwd=os.path.dirname(SRC))
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:]
    assert r.stdout.count("ok   ") == 39
