#!/bin/bash
# Re-entry check of the round's code on a fresh box: GPU tests (with durations), smoke, bench (both arms).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=12 ) > $O/pytest_gpu_r5a.log 2>&1; echo "pytest rc=$?"; tail -22 $O/pytest_gpu_r5a.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_r5a.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_r5a.log
timeout 600 python bench.py > $O/bench_c3_r5a.json 2> $O/bench_c3_r5a.err; echo "bench rc=$?"; cat $O/bench_c3_r5a.json
