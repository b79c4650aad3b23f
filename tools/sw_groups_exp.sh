#!/bin/bash
# Experiment: k_symbols_w[_fir] with its warps in 1 / 2 / 4 staggered groups aligned to the SM sub-partitions,
# and different staggers.  Builds the variants HERE (no GPU needed): bash tools/sw_groups_exp.sh build
# Times them on the GPU box:                                        bash tools/sw_groups_exp.sh run
cd "$(dirname "$0")/.."
VARIANTS="g4:-DSW_GROUPS_N=4 g2s2:-DSW_GROUPS_N=2,-DSW_STAGGER_NS=2500u g2s10:-DSW_GROUPS_N=2,-DSW_STAGGER_NS=10000u g4s5:-DSW_GROUPS_N=4,-DSW_STAGGER_NS=5000u"
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
