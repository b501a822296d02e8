#!/bin/bash
out=gpurun_out/r02c; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -3 $out/pytest.log
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$out/bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "nopal", d["e2e_without_palette_lens"]["ms_per_step"])
print({k:round(v["kernel_ms"]*1e3,2) for k,v in d["modes_per_gpu"].items()})
for k in ("e2e_vertices_to_host","e2e_vertices_to_host_greedy_indexed","e2e_from_positions"): print(k, round(d[k]["ms_per_step"],4), round(d[k]["value"]))
print("cpu", d.get("cpu_baseline",{}).get("value"))
PY
python bench.py --impl reference --steps 5 --warmup 3 > $out/bench_reference_arm.json 2>$out/ref.err; tail -c 400 $out/bench_reference_arm.json
