#!/bin/bash
# Quick A/B of library variants: dense config 3 and example6, one process each, sky tables cached on disk between processes.
# usage: tools/ab_quick.sh <out-log> variant ...   ("default" = the in-tree library)
out=$1; shift
: > $out
export VRT_SKY_CACHE=/tmp/vrt_sky_cache
mkdir -p $VRT_SKY_CACHE
for v in "$@"; do
  if [ "$v" = default ]; then lib=""; else lib="$PWD/voxel_rt2_b200/variants/libvoxelrt_$v.so"; fi
  echo "=== $v" >> $out
  VRT_LIB=$lib timeout 120 python tools/perf_probe.py --sky 1 --iters 10 2>&1 | grep -E "spp/launch=8|Error|error" | sed 's/^/dense    /' >> $out
  VRT_LIB=$lib timeout 120 python tools/perf_probe.py --scene example6 --R 128 --sky 1 --iters 10 2>&1 | grep -E "spp/launch=8|Error|error" | sed 's/^/example6 /' >> $out
done
cat $out
