#!/bin/bash
# Experiment behind DESIGN.md's grouping table: how k_symbols_w's twelve warps are grouped (symbols_warp.cuh:
# SW_GROUPS_PLAIN for the plain kernel, SW_GROUPS_FUSE for the one with the FIR inside; 1 = all in step, 2 / 4 = groups
# on disjoint sub-partitions, -1 = three groups across the sub-partitions).
#   bash tools/sw_groups_exp.sh build     builds the variants HERE into odr-dabmod_b200/exp/ (no GPU needed)
#   bash tools/sw_groups_exp.sh run       times them on the GPU box (C2 step, two-kernel and fused)
# Every run is wrapped in `timeout`: a grouping whose barrier counts do not match its groups hangs the kernel.
cd "$(dirname "$0")/.."
VARIANTS="one:-DSW_GROUPS_PLAIN=1,-DSW_GROUPS_FUSE=1 disjoint2:-DSW_GROUPS_PLAIN=2,-DSW_GROUPS_FUSE=2 disjoint4:-DSW_GROUPS_PLAIN=4,-DSW_GROUPS_FUSE=4 across3:-DSW_GROUPS_PLAIN=-1,-DSW_GROUPS_FUSE=-1"
if [ "$1" = "build" ]; then
  mkdir -p odr-dabmod_b200/exp
  for v in $VARIANTS; do
    name=${v%%:*}; flags=$(echo ${v#*:} | tr ',' ' ')
    ( nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags \
      -c -o /tmp/dabmod_$name.o odr-dabmod_b200/csrc/dabmod_b200.cu 2>/dev/null && \
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o odr-dabmod_b200/exp/libdabmod_b200_$name.so \
      /tmp/dabmod_$name.o odr-dabmod_b200/csrc/coder_b200.o ) &
  done
  wait
  ls -la odr-dabmod_b200/exp
else
  for v in shipped $VARIANTS; do
    name=${v%%:*}
    if [ $name = shipped ]; then unset DABMOD_B200_LIB; else export DABMOD_B200_LIB=$PWD/odr-dabmod_b200/exp/libdabmod_b200_$name.so; fi
    echo "== $name"
    timeout 120 python tools/fuse_time.py 1024 2>&1 | tail -2
  done
fi
