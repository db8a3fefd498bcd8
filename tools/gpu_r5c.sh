#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r5c.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_r5c.log
