#!/bin/bash
# round 2: N-GPU visit: multi-process NCCL parity test + bench under torchrun
N=${1:-2}; TAG=${2:-r02b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader
nproc; nvidia-smi topo -m | head -12
( time python -m pytest tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -5 ) 2>&1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err ) 2>&1 | tail -4
grep -v "^$" gpurun_out/bench_${TAG}_n$N.err | tail -8
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${TAG}_n$N.json") if l.startswith("{")][-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"])
print(json.dumps(d.get("sharded_stream"), indent=1)[:4000])
PY
