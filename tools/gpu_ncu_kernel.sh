#!/bin/bash
# ncu --set full of selected kernels of the C3 bench loop + digests.  usage: tools/gpu_ncu_kernel.sh <tag> <kernel regex> [launch-skip] [count]
tag=${1:-x}; kre=${2:-banded}; skip=${3:-6}; cnt=${4:-3}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s $skip -c $cnt -f -o gpurun_out/prof_$tag python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --skip-c4 --l2 keep > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_src.csv 2>/dev/null
python tools/ncu_keys.py gpurun_out/prof_${tag}_raw.csv > gpurun_out/ncu_keys_$tag.txt
for k in $(echo $kre | tr '|' ' '); do python tools/ncu_hot.py gpurun_out/prof_${tag}_src.csv $k 40 >> gpurun_out/ncu_hot_$tag.txt; done
cat gpurun_out/ncu_keys_$tag.txt; cat gpurun_out/ncu_hot_$tag.txt
rm -f gpurun_out/prof_$tag.ncu-rep
