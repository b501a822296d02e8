#!/bin/bash
# one gpurun call: full GPU suite, bench line, ordered A/B, then (each only after its command exited 0 without ncu) the launch list and one --set full capture
out=gpurun_out/r02b; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -3 $out/pytest.log
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$out/bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"].get("moved_GBps"), "nosum", d["no_summaries"]["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print({k:round(v["kernel_ms"]*1e3,2) for k,v in d["modes_per_gpu"].items()})
print({k:(v.get("ms") or v.get("kernel_ms") or v.get("frame_ms")) for k,v in d["configs"].items() if isinstance(v,dict)})
PY
for lib in tools/ab_libs/base.so despawnengine_b200/lib/libdespawn_b200.so; do DSP_LIB_PATH=$PWD/$lib python tools/ordered_probe.py; done 2>&1 | tee $out/ordered_ab.txt
python tools/prof_r02b.py > $out/prof.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'mesh_chunks|count_chunks|classify' -o $out/r02b_kernels -f python tools/prof_r02b.py >> $out/prof.log 2>&1; tail -2 $out/prof.log
python bench.py --steps 3 --warmup 3 --skip-configs --no-cpu-baseline --sustain-s 0.01 > $out/bench_short.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv python bench.py --steps 3 --warmup 3 --skip-configs --no-cpu-baseline --sustain-s 0.01 > $out/ncu_launches.log 2>&1; tail -1 $out/ncu_launches.log
