#!/bin/bash
# A/B of build variants under gpu-rt_b200/variants/ (interleaved x 3): default build of the stand-in and of a 10 M-triangle soup
for rep in 1 2 3; do for f in gpu-rt_b200/variants/*.so; do v=$(basename $f .so); GPURT_LIB=$PWD/$f python tools/profile_build_scene.py 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v rep$rep standin', round(min(d['build_ms']),4))"; done; done
for f in gpu-rt_b200/variants/*.so; do v=$(basename $f .so); GPURT_LIB=$PWD/$f python tools/profile_build.py --tris 10000000 2>&1 | tail -1 | cut -c1-200 | sed "s/^/$v 10M /"; done
