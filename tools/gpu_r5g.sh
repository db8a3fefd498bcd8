#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in "STOMP_B200_GRAPH=0 STOMP_B200_DEBUG_SKIP=1" "STOMP_B200_GRAPH=1 STOMP_B200_DEBUG_SKIP=1"; do
  echo "== $v"; env $v timeout 300 python tools/timeline.py c3 40 2>&1 | tail -9
done > $O/r5g_host_bound_check.txt 2>&1; cat $O/r5g_host_bound_check.txt
