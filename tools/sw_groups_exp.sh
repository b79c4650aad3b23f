#!/bin/bash
# Experiment: k_symbols_w with its warps in 1 / 2 / 4 staggered groups aligned to the SM sub-partitions.
# Builds the three variants HERE (no GPU needed): bash tools/sw_groups_exp.sh build
# Times them on the GPU box:                      bash tools/sw_groups_exp.sh run
cd "$(dirname "$0")/.."
if [ "$1" = "build" ]; then
  mkdir -p odr-dabmod_b200/exp
  for g in 1 4; do
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DSW_GROUPS_N=$g \
      -c -o /tmp/dabmod_g$g.o odr-dabmod_b200/csrc/dabmod_b200.cu || exit 1
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o odr-dabmod_b200/exp/libdabmod_b200_g$g.so \
      /tmp/dabmod_g$g.o odr-dabmod_b200/csrc/coder_b200.o || exit 1
  done
  ls -la odr-dabmod_b200/exp
else
  for v in base g1 g4; do
    if [ $v = base ]; then unset DABMOD_B200_LIB; else export DABMOD_B200_LIB=$PWD/odr-dabmod_b200/exp/libdabmod_b200_$v.so; fi
    echo "== $v"
    python bench.py --steps 20 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['roofline']['all_kernels'].items()})"
  done
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "native or config2 or full" 2>&1 | tail -2
fi
