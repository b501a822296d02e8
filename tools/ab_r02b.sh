#!/bin/bash
# A/B inside one gpurun call: chunk-summary skipping in the parity kernel, direct stores in the merged-quad kernel (base.so = HEAD before them)
out=gpurun_out/ab_r02b; mkdir -p $out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_quads.py tests/test_gpu_semantics.py tests/test_gpu_fullsize.py tests/test_gpu_edits.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
for s in 0 1 3 0 3; do
  DSP_SUMMARY_SKIP=$s python bench.py --steps 200 --warmup 10 --skip-configs --no-cpu-baseline --e2e-steps 5 > $out/bench_skip$s.json 2>$out/bench_skip$s.err
  python - <<PY
import json
d=json.load(open("$out/bench_skip$s.json"))
print("skip=$s", "ms/step", round(d["ms_per_step"],5), "frac", round(d["roofline"]["frac"],4), "sust", round(d["sustained"]["ms_per_step"],5))
PY
done
for lib in tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so; do
  echo "== $lib"; DSP_LIB_PATH=$PWD/$lib python tools/greedy_probe.py
done 2>&1 | tee $out/greedy_ab.txt
