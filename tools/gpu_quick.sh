#!/bin/bash
# quick: gpu tests + bench line (no profiler)
TAG=${1:-cur}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
for k,v in d["roofline"]["all_kernels"].items(): print(k, v)
for o in d.get("other_configs") or []: print(o)
PY
tail -5 gpurun_out/bench.err
