#!/bin/bash
# A/B the traversal variants under gpu-rt_b200/variants/, interleaved and repeated to average out clock drift
for rep in 1 2 3; do for f in gpu-rt_b200/variants/*.so; do v=$(basename $f .so); GPURT_LIB=$PWD/$f python tools/perf_trace.py 2>&1 | grep -E "primary|bounce" | awk -v v=$v -v r=$rep '{printf "%s rep%s %s %s Mrays/s\n", v, r, $1, $6}'; done; done
