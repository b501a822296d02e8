#!/bin/bash
# fifth session, final build (persistent mesh launch beside the packed upload on by default): GPU test suite, smoke(), default bench line
out=gpurun_out/r02g; mkdir -p $out
timeout 400 python -m pytest tests -m gpu -x -q > $out/tests.txt 2>&1; tail -2 $out/tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 100 python bench.py --impl reference > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err; tail -c 200 $out/bench_reference_arm.json
timeout 200 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; tail -c 300 $out/bench_n1.json
