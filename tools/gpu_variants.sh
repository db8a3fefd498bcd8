#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/variants_$1.txt; : > $out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$1.log 2>&1; echo "pytest rc=$?" | tee -a $out; tail -2 gpurun_out/pytest_gpu_$1.log | tee -a $out
IFS='|' read -ra VARS <<< "$2"
for v in "${VARS[@]}"; do
  echo "=== $v" | tee -a $out
  env $v timeout 300 python tools/timeline.py c3 40 2>&1 | grep -E "per iteration|sample |cost |reuse|weights|update|apply|period" | tee -a $out
done
