#!/bin/bash
# fifth session, final build: the GPU test suite, smoke(), the default bench line and the reference arm on one box
out=gpurun_out/r02f; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/tests.txt 2>&1; tail -2 $out/tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --impl reference > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err; tail -c 300 $out/bench_reference_arm.json
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; tail -c 400 $out/bench_n1.json
