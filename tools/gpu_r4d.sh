#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_self_collision.py -m gpu -x -q > $O/pytest_gpu_r4d.log 2>&1; tail -5 $O/pytest_gpu_r4d.log
timeout 600 python tools/self_collision_cost.py > $O/self_collision_cost_r4d.json 2>&1; grep -E '"pair_rule|"world|"list|"cost"|us_per' $O/self_collision_cost_r4d.json
python tools/self_collision_loop.py 6 && ncu --set full --clock-control none --import-source on -k regex:stomp_b200_states_specialised -s 6 -c 1 -o $O/prof_r4d_self python tools/self_collision_loop.py 6 > $O/ncu_r4d.log 2>&1
ncu -i $O/prof_r4d_self.ncu-rep --page raw --csv > $O/prof_r4d_self_raw.csv 2>/dev/null
