#!/bin/bash
# round 2, first visit: gpu tests + bench (with the sharded-stream leg at N=1)
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc; free -g | head -2
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | tail -4
tail -5 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
print(json.dumps(d.get("sharded_stream"), indent=1)[:3000])
print(d.get("cpu_baseline"))
PY
