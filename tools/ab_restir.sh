#!/bin/bash
# A/B of library variants on the ReSTIR frame (config 4): example6 with sky and example3, temporal reuse on.
# usage: tools/ab_restir.sh <out-log> variant ...   ("default" = the in-tree library)
out=$1; shift
: > $out
for v in "$@"; do
  if [ "$v" = default ]; then lib=""; else lib="$PWD/voxel_rt2_b200/variants/libvoxelrt_$v.so"; fi
  echo "=== $v" >> $out
  VRT_LIB=$lib timeout 300 python tools/perf_probe.py --scene example6 --R 128 --sky 1 --iters 4 --restir 4 --temporal 1 2>&1 | grep -E "restir|rror" | sed 's/^/example6 /' >> $out
  VRT_LIB=$lib timeout 300 python tools/perf_probe.py --scene example3 --R 128 --sky 0 --iters 4 --restir 4 --temporal 1 2>&1 | grep -E "restir|rror" | sed 's/^/example3 /' >> $out
done
cat $out
