#!/bin/bash
# fifth session: pieces of the packed upload cut by content against three pieces fixed up front (same process, alternating calls)
python -m pytest tests/test_gpu_packed_upload.py tests/test_gpu_host_pull.py tests/test_gpu_slot_resolve.py -m gpu -x -q 2>&1 | tail -2
python tools/e2e_pieces_ab.py 2>&1 | tail -40
DSP_TRACE_ONE=2 python tools/e2e_resolve_ab.py 2>&1 | grep -E "trace" | head -64
