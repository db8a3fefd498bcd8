#!/bin/bash
# Final state of the round: timelines, one bench line per other workload, ncu launch list + full capture of the three loop kernels.
cd "$(dirname "$0")/.."
tag=r5k
O=gpurun_out
mkdir -p $O
timeout 300 python tools/timeline.py c3 40 > $O/timeline_c3_$tag.txt 2>&1; tail -9 $O/timeline_c3_$tag.txt
timeout 300 python tools/timeline.py c3 40 flush > $O/timeline_c3_${tag}_flushed.txt 2>&1; tail -9 $O/timeline_c3_${tag}_flushed.txt
timeout 300 python tools/timeline.py c5 20 > $O/timeline_c5_$tag.txt 2>&1; tail -9 $O/timeline_c5_$tag.txt
for w in c2 c4 c5; do
  timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --workload $w > $O/bench_${w}_$tag.json 2> $O/bench_${w}_$tag.err; echo "$w rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3_$tag.csv python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > $O/ncu_launch_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:states_specialised|sample_rollouts_banded|weights_update_kernel" -s 12 -c 9 -f -o $O/prof_$tag python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --l2 keep > $O/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/prof_$tag.ncu-rep --page raw --csv > $O/prof_${tag}_raw.csv 2>/dev/null
ncu -i $O/prof_$tag.ncu-rep --page source --csv > $O/prof_${tag}_src.csv 2>/dev/null
ls -la $O/prof_${tag}*
