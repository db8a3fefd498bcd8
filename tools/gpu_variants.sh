#!/bin/bash
# timeline of the C3 loop under environment knobs.  usage: tools/gpu_variants.sh <tag> "VAR=val VAR2=val" "..." ...
tag=${1:-x}; shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "=== $v"
  env $v timeout 300 python tools/timeline.py c3 40 2>&1 | grep -E "per iteration|median  |gaps|period" | tee -a gpurun_out/variants_$tag.txt
done
