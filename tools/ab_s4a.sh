#!/bin/bash
# merged-quad kernel: call time against the number of CTAs per SM (latency / throughput curve)
for c in 1 2 3 4 5; do echo "DSP_QUAD_CTAS=$c"; DSP_QUAD_CTAS=$c python tools/greedy_probe.py; done
