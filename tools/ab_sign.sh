#!/bin/bash
# A/B of the node test (bvh8.cuh GPURT_NODE_TEST_SIGN): variants/libgpurt_nosign.so vs libgpurt_sign.so, interleaved
for rep in 1 2; do
  for v in nosign sign; do
    GPURT_LIB=$PWD/gpu-rt_b200/variants/libgpurt_$v.so python tools/perf_trace.py 2>&1 | grep -E "primary|bounce|mixed|closest" | sed "s/^/$v rep$rep /" | cut -c1-150
  done
done
for v in nosign sign nosign sign; do
  for i in 0 2 3; do
    GPURT_LIB=$PWD/gpu-rt_b200/variants/libgpurt_$v.so python tools/mis_frame.py $i 10 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v mis_test integ$i', round(d['median'],4))"
  done
done
