#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
{
for m in 5 13 5 13; do
  echo "=== STOMP_B200_PDL=$m"
  STOMP_B200_PDL=$m python tools/timeline.py c3 30 flush | grep -E "workload|period"
  STOMP_B200_PDL=$m python tools/e2e_breakdown.py c3 60 20 8 | grep -E "end to end|solve  |total"
done
} > $O/pdl_tail_r4j.txt 2>&1
cat $O/pdl_tail_r4j.txt
