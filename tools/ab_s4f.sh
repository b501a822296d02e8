#!/bin/bash
for v in "2" "3" "4"; do set -- $v; echo "== DSP_E2E_PACK_PIECES=$1"; DSP_E2E_PACK_PIECES=$1 python tools/e2e_trace.py 2>&1 | tail -32 | grep -E "untraced step us, kernel|all enq|finished" | tail -3; done
