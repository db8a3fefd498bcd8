#!/bin/bash
# compare lag of the hoisted state kernel (links between issuing a link's gathers and comparing them; default 2)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in "STOMP_B200_STATES_LAG=0" "STOMP_B200_STATES_LAG=1" "STOMP_B200_STATES_LAG=2" "STOMP_B200_STATES_LAG=1 STOMP_B200_STATES_MIN_BLOCKS=8" "STOMP_B200_STATES_LAG=1 STOMP_B200_STATES_MIN_BLOCKS=10"; do
  echo "== $v"; env $v timeout 300 python tools/timeline.py c3 40 2>&1 | grep -E "cost  |period"
  env $v timeout 300 python tools/timeline.py c3 40 flush 2>&1 | grep -E "cost  "
done > $O/r5s_state_kernel_compare_lag.txt 2>&1; cat $O/r5s_state_kernel_compare_lag.txt
for v in "STOMP_B200_STATES_LAG=1" "STOMP_B200_STATES_LAG=2" "STOMP_B200_STATES_LAG=0"; do
  echo "== c5 $v"; env $v timeout 300 python tools/timeline.py c5 20 2>&1 | grep -E "cost  |period"
done > $O/r5s_state_kernel_compare_lag_c5.txt 2>&1; cat $O/r5s_state_kernel_compare_lag_c5.txt
