#!/bin/bash
# A/B of k_trace_closest's CTA size / minimum CTAs per SM (trace.cu GPURT_TRACE_BLOCK / GPURT_TRACE_MINB): the variants under
# gpu-rt_b200/variants/ (tools/build_variant.sh <name> trace.cu -D...), interleaved, REPS repetitions
for rep in $(seq 1 ${REPS:-2}); do
  for f in gpu-rt_b200/variants/*.so; do
    v=$(basename $f .so)
    GPURT_LIB=$PWD/$f python tools/perf_trace.py 2>&1 | grep -E "primary|bounce|mixed" | awk -v v=$v -v r=$rep '{printf "%s rep%s %s %s Mrays/s\n", v, r, $1, $6}'
  done
done
