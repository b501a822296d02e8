#!/bin/bash
out=gpurun_out/ab_r02f; mkdir -p $out
python -m pytest tests/test_gpu_quads.py tests/test_gpu_stress.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -3 $out/pytest.log
for lib in tools/ab_libs/q4.so despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/q6.so tools/ab_libs/q4.so despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/q6.so; do echo "== $lib"; DSP_LIB_PATH=$PWD/$lib python tools/greedy_probe.py; done 2>&1 | tee $out/greedy_ctas_ab.txt
for lib in tools/ab_libs/q4.so despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/q6.so; do echo "== $lib"; DSP_LIB_PATH=$PWD/$lib python tools/world_bench.py --dims 512 32 512 --quad-modes 2>&1 | tail -4; done 2>&1 | tee $out/world_quads_ab.txt
