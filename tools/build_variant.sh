#!/bin/bash
# Build gpu-rt_b200/variants/libgpurt_<name>.so = the current library with ONE source recompiled under extra flags, for the
# interleaved A/B scripts (tools/ab*.sh select a variant with GPURT_LIB).  The variants directory is not tracked.
#   tools/build_variant.sh r7 render.cu -DGPURT_RESTIR_MINB=7
#   tools/build_variant.sh base render.cu
#   tools/build_variant.sh nosign all -DGPURT_NODE_TEST_SIGN=0      (every CUDA source recompiled)
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/../gpu-rt_b200"
make -j8 > /dev/null
mkdir -p variants /tmp/gpurt_variant_$name
if [ "$src" = all ]; then
  objs=""
  for c in build sah_build trace cpq order gather render api; do
    nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-Wall,-ffp-contract=off \
         --expt-relaxed-constexpr -Xptxas -v -c csrc/$c.cu -o /tmp/gpurt_variant_$name/$c.o 2> /tmp/gpurt_variant_$name/$c.ptxas.log &
    objs="$objs /tmp/gpurt_variant_$name/$c.o"
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libgpurt_$name.so $objs build/host/scene.o build/host/jpeg.o \
       build/host/host_api.o build/host/sponza_standin.o -lz
  exit 0
fi
obj=/tmp/gpurt_variant_$name/${src%.cu}.o
nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-Wall,-ffp-contract=off \
     --expt-relaxed-constexpr -Xptxas -v -c csrc/$src -o $obj 2> /tmp/gpurt_variant_$name/ptxas.log
objs=""
for o in build.o sah_build.o trace.o cpq.o order.o gather.o render.o api.o; do
  if [ "$o" = "${src%.cu}.o" ]; then objs="$objs $obj"; else objs="$objs build/csrc/$o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libgpurt_$name.so $objs build/host/scene.o build/host/jpeg.o \
     build/host/host_api.o build/host/sponza_standin.o -lz
grep -E "registers|spill" /tmp/gpurt_variant_$name/ptxas.log | sort | uniq -c | sort -rn | head -5
