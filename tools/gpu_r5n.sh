#!/bin/bash
# end event of a timed region behind the last kernel of stomp_b200_run (default) against after its closing wait (STOMP_B200_TIMER_END=host)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r5n.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_r5n.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r5n.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_r5n.log
timeout 600 python bench.py > $O/bench_c3_r5n.json 2> $O/bench_c3_r5n.err; echo "bench rc=$?"; cat $O/bench_c3_r5n.json
STOMP_B200_TIMER_END=host timeout 600 python bench.py --skip-cpu-baseline --skip-c4 > $O/bench_c3_r5n_host_end.json 2> $O/bench_c3_r5n_host_end.err; echo "bench (end event after the wait) rc=$?"; cut -c1-400 $O/bench_c3_r5n_host_end.json
