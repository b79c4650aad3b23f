#!/bin/bash
# Round-2 closing visit of one GPU box: parity tests, smoke, bench (both arms), ncu launch list of the bench command,
# full captures of the kernel families (summarised on the box by tools/ncu_summary.py; only the headline capture is
# kept as .ncu-rep), SASS extracts.
set -x
TAG=${1:-r02final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref_$TAG.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_symbols_w -s 2 -c 1 -f -o gpurun_out/prof_symbols_$TAG python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_sym.log 2>&1
timeout 300 bash tools/prof_kernel.sh k_resample_q c5 resq_$TAG
timeout 300 bash tools/prof_kernel.sh k_resample_up "c3 " resup_$TAG
timeout 300 bash tools/prof_kernel.sh k_symbols_fix n4 fix_$TAG
timeout 300 bash tools/prof_kernel.sh "k_symbols_w" "c1" sym1_$TAG
timeout 300 bash tools/prof_kernel.sh "k_symbols_wg" "c4 TM IV" sym4_$TAG
SUM=gpurun_out/ncu_summary_$TAG.txt
: > $SUM
for r in symbols resq resup fix sym1 sym4; do
  python tools/ncu_summary.py gpurun_out/prof_${r}_$TAG.ncu-rep >> $SUM
  python tools/ncu_summary.py gpurun_out/prof_${r}_$TAG.ncu-rep --all | grep -E "average_warps_issue_stalled|gcc__cache_requests_type_instruction.sum.pct|sm__icc_request_hit_rate" >> $SUM
done
rm -f gpurun_out/prof_resq_$TAG.ncu-rep gpurun_out/prof_resup_$TAG.ncu-rep gpurun_out/prof_fix_$TAG.ncu-rep gpurun_out/prof_sym1_$TAG.ncu-rep gpurun_out/prof_sym4_$TAG.ncu-rep
ls -la gpurun_out | tail -12
