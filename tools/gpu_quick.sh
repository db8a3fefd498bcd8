#!/bin/bash
# short visit: parity tests, one bench line (no CPU arm), the steady-state timeline.  usage: tools/gpu_quick.sh <tag> [pytest -k expression]
tag=${1:-x}; kexpr=${2:-}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/pytest_gpu_$tag.log 2>&1; else timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; fi
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_c3_$tag.err
python - gpurun_out/bench_c3_$tag.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','graph_replays')}, 'steady', d['steady_state'], 'e2e', d['e2e']['value'])
print('roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'], 'kernels', d['kernel_ms_per_step'])
print('c4', d['query_sharded_c4'])
PY
STOMP_B200_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-c4 > gpurun_out/bench_c3_${tag}_nopdl.json 2>/dev/null
python - gpurun_out/bench_c3_${tag}_nopdl.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('NO PDL:', {k:d[k] for k in ('value','ms_per_step','gpu_launches','graph_replays')}, 'steady', d['steady_state'], 'e2e', d['e2e']['value'])
PY
STOMP_B200_GRAPH=0 timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-c4 > gpurun_out/bench_c3_${tag}_nograph.json 2>/dev/null
python - gpurun_out/bench_c3_${tag}_nograph.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('NO GRAPH:', {k:d[k] for k in ('value','ms_per_step','gpu_launches','graph_replays')}, 'steady', d['steady_state'], 'e2e', d['e2e']['value'])
PY
timeout 300 python tools/timeline.py c3 40 > gpurun_out/timeline_c3_$tag.txt 2>&1; cat gpurun_out/timeline_c3_$tag.txt
