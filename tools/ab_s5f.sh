#!/bin/bash
# fifth session: ONE mesh launch beside the upload of its batch, per-piece release words set by the unpack kernels, two side streams
DSP_E2E_PACK_PERSIST=1 timeout 150 python -m pytest tests/test_gpu_packed_upload.py tests/test_gpu_host_pull.py tests/test_gpu_slot_resolve.py tests/test_gpu_semantics.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
for v in "0 6" "1 3" "1 4" "1 6" "1 8" "1 12"; do set -- $v
  echo -n "persist=$1 pieces=$2  "; DSP_E2E_PACK_PERSIST=$1 DSP_E2E_PACK_PERSIST_PIECES=$2 timeout 60 python tools/e2e_resolve_ab.py 2>&1 | tail -1 | cut -c40-230
done; done
