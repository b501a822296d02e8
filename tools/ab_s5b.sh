#!/bin/bash
# fifth session: in-memory trace of the packed upload (three traced calls per context; box slots, then map slots)
DSP_TRACE_ONE=3 python tools/e2e_resolve_ab.py 2>&1 | grep -E "trace|lib"
