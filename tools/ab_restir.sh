#!/bin/bash
# A/B of launch-bounds variants of the shade / tail kernels (gpu-rt_b200/variants/libgpurt_<name>.so: d<k> = direct integrator
# with at least k CTAs per SM, a<k> = material integrator; see render.cu GPURT_*_MINB), interleaved and repeated
for rep in 1 2; do
  for v in base d7 d8; do
    GPURT_LIB=$PWD/gpu-rt_b200/variants/libgpurt_$v.so python tools/mis_frame.py 0 10 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v mis_test integ0', round(d['median'],4))"
    GPURT_LIB=$PWD/gpu-rt_b200/variants/libgpurt_$v.so python tools/default_workload.py 12 0 | tail -1 | cut -c1-200
  done
  for v in base a9 a10; do
    GPURT_LIB=$PWD/gpu-rt_b200/variants/libgpurt_$v.so python tools/mis_frame.py 1 10 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v mis_test integ1', round(d['median'],4))"
    GPURT_LIB=$PWD/gpu-rt_b200/variants/libgpurt_$v.so python tools/default_workload.py 12 1 | tail -1 | cut -c1-200
  done
done
