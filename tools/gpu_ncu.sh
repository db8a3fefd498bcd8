#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> <kernel regex> [extra ncu args]
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on $3 -k "regex:$2" -s 6 -c 2 -f -o gpurun_out/prof_$1 python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --l2 keep > gpurun_out/ncu_full_$1.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv > gpurun_out/prof_$1_src.csv 2>/dev/null
ls -la gpurun_out/prof_$1*
