#!/bin/bash
# fifth session: resolve_positions (run-continuing position -> slot loop) against base.so = the library at 59d0c34, pieces 2 / 3 / 4
python -m pytest tests/test_gpu_slot_resolve.py tests/test_gpu_packed_upload.py tests/test_gpu_host_pull.py tests/test_gpu_semantics.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2 3; do
for lib in tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so; do
for pc in 3 2 4; do
  DSP_LIB_PATH=$PWD/$lib DSP_E2E_PACK_PIECES=$pc python tools/e2e_resolve_ab.py 2>&1 | tail -1
done; done; done
echo "== trace, new library, 3 pieces"
DSP_TRACE_ONE=2 python tools/e2e_resolve_ab.py 2>&1 | grep -E "trace|lib"
