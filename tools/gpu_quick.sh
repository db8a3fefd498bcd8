#!/bin/bash
# Short GPU visit: parity tests + in-pipeline timeline (+ optional bench).  usage: tools/gpu_quick.sh <tag> [bench]
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 300 python tools/timeline.py c3 40 > gpurun_out/timeline_c3_$tag.txt 2>&1; cat gpurun_out/timeline_c3_$tag.txt
if [ -n "$2" ]; then
  timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; echo "bench rc=$?"
  cat gpurun_out/bench_c3_$tag.json
fi
