#!/bin/bash
# one gpurun call: GPU tests on the current build, then A/B of the two-pass ordered mode and of the small first e2e sub-batch
out=gpurun_out/ab_r02c; mkdir -p $out
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_stress.py > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
python -m pytest tests/test_gpu_stress.py -m gpu -x -q > $out/stress.log 2>&1; echo "stress rc=$?" >> $out/stress.log; tail -2 $out/stress.log
for lib in tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so; do
  DSP_LIB_PATH=$PWD/$lib python tools/ordered_probe.py
done 2>&1 | tee $out/ordered_ab.txt
for h in 0 128 0 128 64 256; do echo "== DSP_E2E_HEAD=$h"; DSP_E2E_HEAD=$h python tools/e2e_modes.py; done 2>&1 | tee $out/e2e_head_ab.txt
DSP_TRACE=1 python tools/e2e_trace.py > $out/e2e_trace.txt 2>&1; tail -25 $out/e2e_trace.txt
echo "== greedy"; for lib in tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so; do DSP_LIB_PATH=$PWD/$lib python tools/greedy_probe.py; done 2>&1 | tee $out/greedy_ab.txt
