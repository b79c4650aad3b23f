#!/bin/bash
# Experiment: grouping of k_symbols_w's warps.  build here, run on the GPU box.
cd "$(dirname "$0")/.."
VARIANTS="a3:-DSW_GROUPS_N=3,-DSW_GROUPS_ACROSS,-DSW_STAGGER_NS=6000u a2:-DSW_GROUPS_N=2,-DSW_GROUPS_ACROSS,-DSW_STAGGER_NS=9000u"
if [ "$1" = "build" ]; then
  mkdir -p odr-dabmod_b200/exp
  for v in $VARIANTS; do
    name=${v%%:*}; flags=$(echo ${v#*:} | tr ',' ' ')
    ( nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags \
      -c -o /tmp/dabmod_$name.o odr-dabmod_b200/csrc/dabmod_b200.cu && \
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o odr-dabmod_b200/exp/libdabmod_b200_$name.so \
      /tmp/dabmod_$name.o odr-dabmod_b200/csrc/coder_b200.o ) &
  done
  wait
  ls -la odr-dabmod_b200/exp
else
  for v in base $VARIANTS; do
    name=${v%%:*}
    if [ $name = base ]; then unset DABMOD_B200_LIB; else export DABMOD_B200_LIB=$PWD/odr-dabmod_b200/exp/libdabmod_b200_$name.so; fi
    echo "== $name"
    python tools/fuse_time.py 2>&1 | tail -2
  done
fi
