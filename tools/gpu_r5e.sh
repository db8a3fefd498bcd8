#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/host_enqueue_cost.py c3 150 > $O/r5e_host_enqueue_cost.txt 2>&1; cat $O/r5e_host_enqueue_cost.txt
