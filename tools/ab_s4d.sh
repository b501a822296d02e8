#!/bin/bash
for rep in 1 2; do
for lib in despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/ctas4.so; do
for v in 0 1; do
  echo "== $lib DSP_GREEDY_SELF=$v"; DSP_LIB_PATH=$PWD/$lib DSP_GREEDY_SELF=$v python tools/greedy_probe.py | head -1 | cut -c1-330
done; done; done 2>&1 | grep -v -E "Traceback|File|print|BrokenPipe"
