#!/bin/bash
# Host half under AddressSanitizer + UndefinedBehaviorSanitizer, mutation-fuzzed through the C ABI (tools/fuzz_loader.cpp).
#   tools/fuzz_loader.sh [iterations per seed, default 2000] [rng seed]
# Seeds: every glTF / GLB scene and every PNG / JPEG texture under tests/data.  Exit code 0 = sanitizers silent on all of them.
set -e
cd "$(dirname "$0")/.."
N=${1:-2000}; R=${2:-1}
B=/tmp/gpurt_fuzz_build; mkdir -p $B
SAN="-fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer"
for f in scene jpeg host_api sponza_standin; do
  g++ -O1 -g -std=c++17 -ffp-contract=off $SAN -c gpu-rt_b200/host/$f.cpp -o $B/$f.o &
done
g++ -O1 -g -std=c++17 $SAN -c tools/fuzz_loader.cpp -o $B/fuzz_loader.o &
wait
g++ $SAN -o $B/fuzz_loader $B/*.o -lz
export ASAN_OPTIONS=detect_leaks=1:allocator_may_return_null=1:max_allocation_size_mb=4096 UBSAN_OPTIONS=print_stacktrace=1
rc=0
for s in tests/data/media/cube.gltf tests/data/media/cbox/cbox.gltf tests/data/media/mis_test/mis_test.gltf \
         tests/data/synth/features.gltf tests/data/synth/textures.gltf tests/data/synth/embedded.glb \
         tests/data/synth/*.png tests/data/synth/*.jpg; do
  $B/fuzz_loader "$s" "$N" "$R" 2> $B/last.err || { rc=1; echo "FAILED on $s (rng seed $R):"; head -n 40 $B/last.err; }
done
exit $rc
