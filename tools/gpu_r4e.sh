#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
{
for cfg in "X=1" "STOMP_B200_STATES_PER_THREAD=2" "STOMP_B200_STATES_PER_THREAD=2 STOMP_B200_STATES_MIN_BLOCKS=5" "STOMP_B200_STATES_PER_THREAD=2 STOMP_B200_STATES_MIN_BLOCKS=4" "STOMP_B200_STATES_PER_THREAD=2 STOMP_B200_STATES_BLOCK=64" "STOMP_B200_STATES_PER_THREAD=2 STOMP_B200_STATES_BLOCK=96" "STOMP_B200_STATES_PER_THREAD=2 STOMP_B200_STATES_BLOCK=256" "STOMP_B200_STATES_PER_THREAD=2 STOMP_B200_STATES_LAG=0" "STOMP_B200_STATES_PER_THREAD=2 STOMP_B200_STATES_LAG=7"; do
  echo "=== $cfg"
  env $cfg python tools/timeline.py c3 40 | grep -E "cost|period"
  env $cfg python tools/timeline.py c3 20 flush | grep -E "cost  |period"
done
for cfg in "X=1" "STOMP_B200_STATES_PER_THREAD=2"; do
  echo "=== c5 $cfg"; env $cfg python tools/timeline.py c5 30 | grep -E "cost  |period"
  echo "=== c4 $cfg"; env $cfg python tools/timeline.py c4 10 | grep -E "cost  |period"
done
} > $O/state_spt_variants_r4e.txt 2>&1
cat $O/state_spt_variants_r4e.txt
