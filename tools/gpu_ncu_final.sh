#!/bin/bash
# launch list + ncu --set full of the main kernels for the final state of a round
tag=${1:-x}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3_$tag.csv python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:states_specialised|control_rows_tile|sample_rollouts_dmma|weights_update_kernel|noiseless" -s 12 -c 10 -f -o gpurun_out/prof_$tag python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --l2 keep > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_src.csv 2>/dev/null
ls -la gpurun_out/prof_${tag}*
