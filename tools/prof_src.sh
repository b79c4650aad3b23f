#!/bin/bash
# usage: tools/prof_src.sh <kernel regex> <other_configs prefix> <tag>
# one ncu --set full capture; keeps the text summary and the per-instruction stall samples (CSV), not the .ncu-rep
K=$1; C=$2; TAG=$3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --only "$C" --steps 2 > gpurun_out/ncu_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep > gpurun_out/sum_$TAG.txt
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep --all | grep -E "average_warps_issue_stalled" >> gpurun_out/sum_$TAG.txt
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/src_$TAG.csv 2>/dev/null
rm -f gpurun_out/prof_$TAG.ncu-rep
