#!/bin/bash
out=gpurun_out/ab_r02e; mkdir -p $out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_semantics.py tests/test_gpu_export.py tests/test_gpu_stress.py tests/test_gpu_seams.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
for lib in tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so; do
  DSP_LIB_PATH=$PWD/$lib python tools/ordered_probe.py
done 2>&1 | tee $out/ordered_ab.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'count_chunks' --csv --log-file $out/count_launches.csv python tools/prof_r02b.py > /dev/null 2>&1; grep count_chunks $out/count_launches.csv | cut -d, -f5,15- | head
