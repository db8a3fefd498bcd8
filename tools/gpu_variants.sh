#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/variants_$1.txt; : > $out
IFS='|' read -ra VARS <<< "$2"
for v in "${VARS[@]}"; do
  echo "=== $v" | tee -a $out
  env $v timeout 300 python tools/timeline.py c3 40 2>&1 | grep -E "per iteration|sample |cost |rows|period" | tee -a $out
done
