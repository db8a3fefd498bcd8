#!/bin/bash
# last visit of the round on one GPU: the GPU tests, smoke, the default bench line
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r5t.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_r5t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r5t.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_r5t.log
timeout 600 python bench.py > $O/bench_c3_r5t.json 2> $O/bench_c3_r5t.err; echo "bench rc=$?"; cat $O/bench_c3_r5t.json | cut -c1-5000
