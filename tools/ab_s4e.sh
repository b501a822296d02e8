#!/bin/bash
# host-chunk calls: packed upload (default) against the kernel pull / copy-engine pipeline (DSP_E2E_PACK=0), in one call
python -m pytest tests/test_gpu_host_pull.py tests/test_gpu_semantics.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in 1 0 1 0; do echo "== DSP_E2E_PACK=$v"; DSP_E2E_PACK=$v python tools/e2e_trace.py 2>&1 | grep -v "^\[trace\] sub-batch"; done
