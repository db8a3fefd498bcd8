#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in "STOMP_B200_GRAPH=0 STOMP_B200_PDL=13" "STOMP_B200_GRAPH=0 STOMP_B200_PDL=9" "STOMP_B200_GRAPH=0 STOMP_B200_PDL=15" "STOMP_B200_GRAPH=0 STOMP_B200_PDL=11"; do
  echo "== $v"; env $v timeout 300 python tools/timeline.py c3 40 2>&1 | tail -9
done > $O/r5f_early_sampler_pdl_masks.txt 2>&1; cat $O/r5f_early_sampler_pdl_masks.txt
