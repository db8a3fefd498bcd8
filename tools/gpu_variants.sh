#!/bin/bash
# timeline of the C3 loop under environment knobs, steady state and isolated / L2-flushed.  usage: tools/gpu_variants.sh <tag> "VAR=val VAR2=val" "..." ...
tag=${1:-x}; shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "=== $v" | tee -a gpurun_out/variants_$tag.txt
  env $v timeout 300 python tools/timeline.py ${WORKLOAD:-c3} 40 2>&1 | grep -E "per iteration|cost  |period" | tee -a gpurun_out/variants_$tag.txt
  env $v timeout 300 python tools/timeline.py ${WORKLOAD:-c3} 20 flush 2>&1 | grep -E "per iteration|cost  " | tee -a gpurun_out/variants_$tag.txt
done
