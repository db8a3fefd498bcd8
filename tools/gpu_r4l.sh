#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_self_collision.py -m gpu -x -q > $O/pytest_gpu_r4l.log 2>&1; tail -3 $O/pytest_gpu_r4l.log
timeout 600 python tools/self_collision_cost.py > $O/self_collision_cost_r4l.json 2>&1; grep -E '"pair_rule|"world|"list|"cost"|us_per' $O/self_collision_cost_r4l.json
