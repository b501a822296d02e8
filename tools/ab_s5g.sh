#!/bin/bash
# fifth session, final build: the persistent launch (default) against a mesh launch per piece, same box, alternating processes
for v in 1 0 1 0 1 0; do echo -n "persist=$v  "; DSP_E2E_PACK_PERSIST=$v timeout 40 python tools/e2e_resolve_ab.py 2>&1 | tail -1 | cut -c40-200; done
