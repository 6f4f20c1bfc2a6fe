#!/bin/bash
# Registers / spills of the static-camera path kernel for a set of -D defines (no GPU needed).
# usage: tools/ptxas_probe.sh "VRT_PATH_THREADS=1024 VRT_STASH_MASK=7" ...
for spec in "$@"; do
  defs=""; for d in $spec; do defs="$defs -D$d"; done
  echo "=== $spec"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -ccbin /usr/bin/g++ $defs \
    -c voxel_rt2_b200/csrc/vrt_render.cu -o /tmp/ptxas_probe.o 2>&1 | grep -A3 "k_pathILb0ELi0ELb1" | grep -E "spill|Used"
done
