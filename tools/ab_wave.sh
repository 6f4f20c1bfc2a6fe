#!/bin/bash
# A/B of the two path-kernel formulations (and library variants) on the config-3 workload, example6 and the sparse city scene.
# usage: tools/ab_wave.sh <out-log> kernel[:variant] ...    e.g.  path wave wave:wave_s2
out=$1; shift
: > $out
for kv in "$@"; do
  k=${kv%%:*}; v=${kv#*:}; [ "$v" = "$kv" ] && v=default
  if [ "$v" = default ]; then lib=""; else lib="$PWD/voxel_rt2_b200/variants/libvoxelrt_$v.so"; fi
  echo "=== kernel $k variant $v" >> $out
  VRT_KERNEL=$k VRT_LIB=$lib timeout 300 python tools/perf_probe.py --sky 1 --iters 8 2>&1 | grep -E "spp/launch=8|Error|error" | sed 's/^/dense    /' >> $out
  VRT_KERNEL=$k VRT_LIB=$lib timeout 300 python tools/perf_probe.py --scene example6 --R 128 --sky 1 --iters 8 2>&1 | grep -E "spp/launch=8|Error|error" | sed 's/^/example6 /' >> $out
  VRT_KERNEL=$k VRT_LIB=$lib timeout 300 python tools/perf_probe.py --scene city --R 128 --sky 0 --iters 8 2>&1 | grep -E "spp/launch=8|Error|error" | sed 's/^/city     /' >> $out
done
cat $out
