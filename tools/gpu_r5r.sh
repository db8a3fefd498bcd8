#!/bin/bash
# the hoisted state kernel under different register bounds (default: 9 CTAs of 128 threads per SM, 56 registers)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in "A=1" "STOMP_B200_STATES_MIN_BLOCKS=8" "STOMP_B200_STATES_MIN_BLOCKS=10" "STOMP_B200_STATES_MIN_BLOCKS=12" "STOMP_B200_STATES_LAG=1" "STOMP_B200_STATES_LAG=3"; do
  echo "== $v"; env $v timeout 300 python tools/timeline.py c3 40 2>&1 | grep -E "cost  |period|per iteration"
done > $O/r5r_state_kernel_register_bounds.txt 2>&1; cat $O/r5r_state_kernel_register_bounds.txt
