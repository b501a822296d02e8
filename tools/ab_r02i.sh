#!/bin/bash
out=gpurun_out/ab_r02i; mkdir -p $out
DSP_LIB_PATH=$PWD/tools/ab_libs/early.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_summaries.py -m gpu -x -q 2>&1 | tail -1
for lib in despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/early.so despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/early.so; do
  DSP_LIB_PATH=$PWD/$lib python bench.py --steps 200 --warmup 10 --skip-configs --no-cpu-baseline --e2e-steps 5 > $out/bench.json 2>$out/bench.err
  python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("$lib", "ms/step", round(d["ms_per_step"],5), "frac", round(d["roofline"]["frac"],4), "nosum", round(d["no_summaries"]["ms_per_step"],5), "e2e", round(d["e2e"]["ms_per_step"],4))
PY
done 2>&1 | tee $out/bench_ab.txt
