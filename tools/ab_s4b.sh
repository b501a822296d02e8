#!/bin/bash
# merged-quad calls: library at HEAD against the working tree (with / without the early trigger in the pre-pass), alternating, in one call
python -m pytest tests/test_gpu_quads.py -m gpu -x -q 2>&1 | tail -1
for rep in 1 2; do
for v in "tools/ab_libs/head.so 1" "despawnengine_b200/lib/libdespawn_b200.so 1" "despawnengine_b200/lib/libdespawn_b200.so 0"; do
  set -- $v
  echo "== $1 DSP_CLASSIFY_PDL=$2"; DSP_CLASSIFY_PDL=$2 DSP_LIB_PATH=$PWD/$1 python tools/greedy_probe.py | head -1 | cut -c1-330
done; done
