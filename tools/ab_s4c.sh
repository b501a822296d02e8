#!/bin/bash
# merged-quad calls: pre-pass (DSP_GREEDY_SELF=0) against the self-classifying prologue, alternating, in one call
python -m pytest tests/test_gpu_quads.py tests/test_gpu_summaries.py tests/test_gpu_host_pull.py tests/test_gpu_stress.py tests/test_gpu_semantics.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
for v in 0 1; do
  echo "== DSP_GREEDY_SELF=$v"; DSP_GREEDY_SELF=$v python tools/greedy_probe.py | cut -c1-330
done; done
