#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full captures of the kernels.
# usage: bash tools/gpu_round.sh [tag]   (NCU=0 skips the profiler passes)
# gpurun copies back at most 64 MiB: every capture is summarised on the box (tools/ncu_summary.py) and only
# the two captures of the headline step are kept as .ncu-rep.
set -x
TAG=${1:-cur}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref_$TAG.json
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_symbols -s 2 -c 1 -f -o gpurun_out/prof_symbols_$TAG python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_sym.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fir -s 2 -c 1 -f -o gpurun_out/prof_fir_$TAG python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_fir.log 2>&1
bash tools/prof_kernel.sh k_resample_q c5 resq_$TAG
bash tools/prof_kernel.sh k_resample_up "c3 " resup_$TAG
bash tools/prof_kernel.sh k_symbols_fix n4 fix_$TAG
bash tools/prof_kernel.sh "^k_symbols$" "c4 TM IV" sym4_$TAG
SUM=gpurun_out/ncu_summary_$TAG.txt
: > $SUM
for r in symbols fir resq resup fix sym4; do
  python tools/ncu_summary.py gpurun_out/prof_${r}_$TAG.ncu-rep >> $SUM
  python tools/ncu_summary.py gpurun_out/prof_${r}_$TAG.ncu-rep --all | grep -E "average_warps_issue_stalled" >> $SUM
done
rm -f gpurun_out/prof_resq_$TAG.ncu-rep gpurun_out/prof_resup_$TAG.ncu-rep gpurun_out/prof_fix_$TAG.ncu-rep gpurun_out/prof_sym4_$TAG.ncu-rep
ls -la gpurun_out | tail -12
fi
