#!/bin/bash
# A/B device-time probe of library variants on the config-3 workload (and optionally others).
# usage: tools/ab_probe.sh <out-log> [variant-name ...]   ("default" = voxel_rt2_b200/libvoxelrt.so)
out=$1; shift
: > $out
for v in "$@"; do
  if [ "$v" = default ]; then lib=""; else lib="$PWD/voxel_rt2_b200/variants/libvoxelrt_$v.so"; fi
  for rep in 1 2; do
    echo "=== $v (run $rep)" >> $out
    VRT_LIB=$lib python tools/perf_probe.py --sky 1 --iters 8 2>&1 | grep -E "spp/launch|stats" >> $out
  done
done
cat $out
