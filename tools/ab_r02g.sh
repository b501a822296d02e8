#!/bin/bash
# record phase with rows handed out by y-level: A/B against the previous commit (base2.so) inside one call
out=gpurun_out/ab_r02g; mkdir -p $out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_summaries.py tests/test_gpu_fullsize.py tests/test_gpu_stress.py tests/test_gpu_edits.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -3 $out/pytest.log
for lib in tools/ab_libs/base2.so despawnengine_b200/lib/libdespawn_b200.so tools/ab_libs/base2.so despawnengine_b200/lib/libdespawn_b200.so; do
  DSP_LIB_PATH=$PWD/$lib python bench.py --steps 200 --warmup 10 --skip-configs --no-cpu-baseline --e2e-steps 5 > $out/bench.json 2>$out/bench.err
  python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("$lib", "ms/step", round(d["ms_per_step"],5), "frac", round(d["roofline"]["frac"],4), "nosum", round(d["no_summaries"]["ms_per_step"],5), {k:round(v["kernel_ms"]*1e3,1) for k,v in d["modes_per_gpu"].items()})
PY
done 2>&1 | tee $out/bench_ab.txt
for lib in tools/ab_libs/base2.so despawnengine_b200/lib/libdespawn_b200.so; do echo "== $lib"; DSP_LIB_PATH=$PWD/$lib python tools/world_bench.py --dims 512 32 512 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: continue
    if 'mode' in d: print(d['mode'], round(d['mesh_wall_s']*1e3,2),'ms', round(d['hbm_algorithmic_GBps_per_gpu']),'GB/s algorithmic')
"; done 2>&1 | tee $out/world_ab.txt
for lib in tools/ab_libs/base2.so despawnengine_b200/lib/libdespawn_b200.so; do echo "== $lib"; DSP_LIB_PATH=$PWD/$lib python tools/config_bench.py 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: continue
    print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ('config','name','kernel_ms','ms','frame_ms','chunks_per_s','algorithmic_GBps','frac','mode')})
"; done 2>&1 | tee $out/config_ab.txt
