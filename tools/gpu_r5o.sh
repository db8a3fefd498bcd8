#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --skip-cpu-baseline --skip-c4 > $O/bench_c3_r5o.json 2> $O/bench_c3_r5o.err; echo "bench rc=$?"; cut -c1-300 $O/bench_c3_r5o.json
STOMP_B200_TIMER_END=host timeout 600 python bench.py --skip-cpu-baseline --skip-c4 > $O/bench_c3_r5o_host_end.json 2> $O/bench_c3_r5o_host_end.err; echo "bench (host-side events) rc=$?"; cut -c1-300 $O/bench_c3_r5o_host_end.json
timeout 300 python tools/ab_early_sampler.py c3 40 5 2>&1 | tail -8
