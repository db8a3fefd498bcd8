#!/bin/bash
# tests + C3 timeline + one bench line per workload (single GPU)
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python tools/timeline.py c3 40 2>&1 | tee gpurun_out/timeline_c3_$tag.txt | grep -E "per iteration|sample |cost |rows|weights|update|period"
for w in c3 c2 c4 c5; do
  timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --workload $w > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err; echo "$w rc=$?"
  python - gpurun_out/bench_${w}_$tag.json <<'PY'
import json,sys
try:
    d=json.loads([x for x in open(sys.argv[1]).read().splitlines() if x.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'steady', d['steady_state']['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['with_control_rows']['frac'], d['kernel_ms_per_step'], d['clocks'])
except Exception as e: print('ERR', e)
PY
done
