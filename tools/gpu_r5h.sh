#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r5h.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_r5h.log
timeout 600 python bench.py > $O/bench_c3_r5h.json 2> $O/bench_c3_r5h.err; echo "bench rc=$?"; cat $O/bench_c3_r5h.json
STOMP_B200_SAMPLER_EARLY=0 timeout 600 python bench.py --skip-cpu-baseline > $O/bench_c3_r5h_late.json 2> $O/bench_c3_r5h_late.err; echo "bench (late) rc=$?"; cat $O/bench_c3_r5h_late.json
timeout 300 python tools/e2e_breakdown.py c3 > $O/r5h_e2e_breakdown.txt 2>&1; tail -25 $O/r5h_e2e_breakdown.txt
