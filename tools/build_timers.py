"""Builds despawnengine_b200/lib/libdespawn_timers.so: the library with -DDSP_PHASE_TIMERS (debug only, git-ignored)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from despawnengine_b200 import build as b
out = os.path.join(b.LIB_DIR, "libdespawn_timers.so")
subprocess.run([b.nvcc()] + b.NVCC_FLAGS + ["-DDSP_PHASE_TIMERS", "-o", out] + b.SOURCES, check=True)
print(out)
