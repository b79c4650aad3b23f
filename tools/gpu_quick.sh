#!/bin/bash
# quick visit: microbenchmarks + bench line
set -x
mkdir -p gpurun_out
./tools/ubench/fp32_pipes > gpurun_out/fp32_pipes.txt 2>&1; cat gpurun_out/fp32_pipes.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
