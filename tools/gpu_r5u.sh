#!/bin/bash
# other knobs of the hoisted, lag-1 state kernel (sincos batching, CTA size, prefetch)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in "A=1" "STOMP_B200_STATES_BATCH=2" "STOMP_B200_STATES_BATCH=4" "STOMP_B200_STATES_BATCH=7" "STOMP_B200_STATES_BLOCK=64" "STOMP_B200_STATES_BLOCK=96" "STOMP_B200_STATES_BLOCK=160" "STOMP_B200_STATES_BLOCK=256" "STOMP_B200_STATES_PREFETCH=4" "STOMP_B200_STATES_PER_THREAD=2"; do
  echo "== $v"; env $v timeout 120 python tools/timeline.py c3 30 2>&1 | grep -E "cost  |period"
done > $O/r5u_state_kernel_other_knobs.txt 2>&1; cat $O/r5u_state_kernel_other_knobs.txt
