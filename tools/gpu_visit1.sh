#!/bin/bash
# round-2 visit 1: new full-size parity + SDF builder tests, then the whole GPU suite, then a bench line
tag=${1:-v1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$tag.txt 2>&1
nproc >> gpurun_out/smi_$tag.txt; free -g >> gpurun_out/smi_$tag.txt
timeout 1200 python -m pytest tests/test_sdf_builder.py tests/test_full_size_parity.py -m gpu -x -q --durations=12 > gpurun_out/pytest_new_$tag.log 2>&1; echo "pytest new rc=$?"
tail -25 gpurun_out/pytest_new_$tag.log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 --deselect tests/test_full_size_parity.py --deselect tests/test_sdf_builder.py > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/pytest_gpu_$tag.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_c3_$tag.json; tail -3 gpurun_out/bench_c3_$tag.err
