#!/bin/bash
# One GPU visit: parity tests, bench, in-pipeline timeline, then ncu (launch list + full capture of the cost kernels).
# usage: tools/gpu_round.sh <tag> [ncu kernel regex]
tag=${1:-x}
kre=${2:-"rollout_states_kernel|control_rows_kernel"}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_c3_$tag.json
timeout 300 python tools/timeline.py c3 40 > gpurun_out/timeline_c3_$tag.txt 2>&1; cat gpurun_out/timeline_c3_$tag.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3_$tag.csv python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s 12 -c 12 -f -o gpurun_out/prof_$tag python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --l2 keep > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_src.csv 2>/dev/null
ls -la gpurun_out | tail -8
