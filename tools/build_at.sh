#!/bin/bash
# tools/build_at.sh <commit> <name>: the library as of <commit> -> tools/ab_libs/<name>.so (base of same-call A/B runs)
set -e
c=$1; name=$2; d=$(mktemp -d)
git archive "$c" despawnengine_b200/csrc include | tar -x -C "$d"
mkdir -p tools/ab_libs
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared -cudart static \
  -o tools/ab_libs/$name.so "$d/despawnengine_b200/csrc/dsp_api.cu" $(ls "$d/despawnengine_b200/csrc/host_pack.cpp" 2>/dev/null) 2>&1 | grep -v -E "warning|Remark|^$|\^|declared but never" | head -5
rm -rf "$d"; ls -la tools/ab_libs/$name.so
