#!/bin/bash
# eight concurrent copies of tools/e2e_pack_sweep.py, one per GPU, each on its own slice of the CPUs (what bench.py --gpus 8 does per rank)
n=$(nproc); per=$((n / 8))
for g in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$g taskset -c $((g * per))-$((g * per + per - 1)) python tools/e2e_pack_sweep.py $((per - 1)) > gpurun_out/s4_pack_sweep_n8_$g.txt 2>&1 &
done
wait
echo "CPUs $n, $per per GPU"; cat gpurun_out/s4_pack_sweep_n8_0.txt gpurun_out/s4_pack_sweep_n8_5.txt
