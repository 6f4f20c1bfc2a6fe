#!/bin/bash
# A/B of library variants / environment settings over three scene types (dense config 3, example6 with sky, sparse city).
# usage: tools/ab_env.sh <out-log> "label|variant|ENV=VAL ENV2=VAL2" ...   (variant "default" = the in-tree library)
out=$1; shift
: > $out
for spec in "$@"; do
  IFS='|' read -r label v envs <<< "$spec"
  if [ "$v" = default ]; then lib=""; else lib="$PWD/voxel_rt2_b200/variants/libvoxelrt_$v.so"; fi
  echo "=== $label" >> $out
  for rep in 1 2; do
  env $envs VRT_LIB=$lib timeout 300 python tools/perf_probe.py --sky 1 --iters 8 2>&1 | grep -E "spp/launch=8|Error|error" | sed 's/^/dense    /' >> $out
  done
  env $envs VRT_LIB=$lib timeout 300 python tools/perf_probe.py --scene example6 --R 128 --sky 1 --iters 8 2>&1 | grep -E "spp/launch=8|Error|error" | sed 's/^/example6 /' >> $out
  env $envs VRT_LIB=$lib timeout 300 python tools/perf_probe.py --scene city --R 128 --sky 0 --iters 8 2>&1 | grep -E "spp/launch=8|Error|error" | sed 's/^/city     /' >> $out
done
cat $out
