#!/bin/bash
# A/B of library variants over three scene types (dense config 3, example6 with sky, sparse city).
# usage: tools/ab_scenes.sh variant [variant ...]   ("default" = the in-tree library)
for v in "$@"; do
  if [ "$v" = default ]; then lib=""; else lib="$PWD/voxel_rt2_b200/variants/libvoxelrt_$v.so"; fi
  echo "=== $v"
  VRT_LIB=$lib python tools/perf_probe.py --sky 1 --iters 8 2>&1 | grep -E "spp/launch=8" | sed 's/^/dense    /'
  VRT_LIB=$lib python tools/perf_probe.py --scene example6 --R 128 --sky 1 --iters 8 2>&1 | grep -E "spp/launch=8" | sed 's/^/example6 /'
  VRT_LIB=$lib python tools/perf_probe.py --scene city --R 128 --sky 0 --iters 8 2>&1 | grep -E "spp/launch=8" | sed 's/^/city     /'
done
