#!/bin/bash
out=gpurun_out/ab_r02d; mkdir -p $out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_semantics.py tests/test_gpu_export.py tests/test_gpu_fullsize.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
for lib in tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so; do
  DSP_LIB_PATH=$PWD/$lib python tools/ordered_probe.py
done 2>&1 | tee $out/ordered_ab.txt
for lib in tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so; do echo "== $lib"; DSP_LIB_PATH=$PWD/$lib python tools/e2e_modes.py; done 2>&1 | tee $out/e2e_ab.txt
for h in 0 128; do DSP_E2E_HEAD=$h DSP_TRACE=1 python tools/e2e_trace.py 2>&1 | tail -12; done | tee $out/e2e_trace.txt
