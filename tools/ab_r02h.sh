#!/bin/bash
out=gpurun_out/ab_r02h; mkdir -p $out
python -m pytest tests/test_gpu_quads.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py tests/test_scene.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -3 $out/pytest.log
for lib in tools/ab_libs/base3.so despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/base3.so despawnengine_b200/lib/libdespawn_b200.so; do echo "== $lib"; DSP_LIB_PATH=$PWD/$lib python tools/greedy_probe.py; done 2>&1 | tee $out/greedy_ab.txt
