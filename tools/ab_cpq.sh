#!/bin/bash
# A/B the closest-point variants under gpu-rt_b200/variants/ (interleaved x3): bench-style near-surface queries
for rep in 1 2 3; do for f in gpu-rt_b200/variants/*.so; do v=$(basename $f .so); GPURT_LIB=$PWD/$f python tools/perf_trace.py 2>&1 | grep -E "cpq" | awk -v v=$v -v r=$rep '{printf "%s rep%s %s %s Mq/s\n", v, r, $1, $6}'; done; done
