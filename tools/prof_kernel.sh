#!/bin/bash
# usage: tools/prof_kernel.sh <kernel regex> <other_configs prefix> <tag>
# one ncu --set full capture of one kernel of one configuration
K=$1; C=$2; TAG=$3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --only "$C" --steps 2 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
