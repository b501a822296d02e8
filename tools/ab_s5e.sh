#!/bin/bash
# fifth session: ONE mesh launch beside the upload of its batch (DSP_E2E_PACK_PERSIST=1) against a mesh launch per piece
DSP_E2E_PACK_PERSIST=1 timeout 150 python -m pytest tests/test_gpu_packed_upload.py tests/test_gpu_host_pull.py tests/test_gpu_slot_resolve.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2 3; do
for v in "0 6" "1 6" "1 4" "1 8" "1 12"; do set -- $v
  echo -n "persist=$1 pieces=$2  "; DSP_E2E_PACK_PERSIST=$1 DSP_E2E_PACK_PERSIST_PIECES=$2 timeout 60 python tools/e2e_resolve_ab.py 2>&1 | tail -1 | cut -c1-260
done; done
echo "== trace, persist, 6 pieces"
DSP_E2E_PACK_PERSIST=1 DSP_TRACE_ONE=2 timeout 60 python tools/e2e_resolve_ab.py 2>&1 | grep -E "trace" | head -60
