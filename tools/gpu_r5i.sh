#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in "A=1" "STOMP_B200_UPDATE_CARVEOUT=100" "STOMP_B200_UPDATE_CARVEOUT=50"; do
  echo "== $v"; env $v timeout 300 python tools/timeline.py c3 40 2>&1 | tail -9
done > $O/r5i_update_carveout.txt 2>&1; cat $O/r5i_update_carveout.txt
