#!/bin/bash
out=gpurun_out/r02e; mkdir -p $out
python tools/prof_r02e.py > $out/prof.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'unpack_host|mesh_chunks' -c 14 -o $out/r02e_kernels -f python tools/prof_r02e.py >> $out/prof.log 2>&1; tail -2 $out/prof.log
python bench.py --steps 3 --warmup 3 --skip-configs --no-cpu-baseline --sustain-s 0.01 > $out/bench_short.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches.csv python bench.py --steps 3 --warmup 3 --skip-configs --no-cpu-baseline --sustain-s 0.01 > $out/ncu_launches.log 2>&1; tail -c 200 $out/ncu_launches.log
