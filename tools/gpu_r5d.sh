#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r5d.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_r5d.log
timeout 300 python tools/ab_early_sampler.py c3 40 7 > $O/r5d_early_sampler_ab.txt 2>&1; cat $O/r5d_early_sampler_ab.txt
for v in "STOMP_B200_GRAPH=0" "STOMP_B200_GRAPH=0 STOMP_B200_SAMPLER_EARLY=0"; do
  echo "== $v"; env $v timeout 300 python tools/timeline.py c3 40 2>&1 | tail -9
done > $O/r5d_early_sampler_timeline.txt 2>&1; cat $O/r5d_early_sampler_timeline.txt
