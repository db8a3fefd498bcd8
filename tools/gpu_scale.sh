#!/bin/bash
# the driver's scaling line at N ranks: the C3 bench under torchrun (rollout sharding + the query-sharded C4 figure), then rank 0's steady timeline
n=${1:-2}; tag=${2:-x}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29544 bench.py --gpus $n --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/scale_c3_n${n}_$tag.json 2> gpurun_out/scale_c3_n${n}_$tag.err; echo "n$n rc=$?"
python - gpurun_out/scale_c3_n${n}_$tag.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','graph_replays','parity_ok')}, 'steady', d['steady_state']['ms_per_step'], 'e2e', d['e2e']['value'], 'c4', d['query_sharded_c4'].get('value'), d['config']['exchange'][:20])
PY
run 29541 tools/timeline.py c3 40 2>&1 | grep -v "^\*\*\*\|^$\|NCCL version\|Setting OMP\|Warning" | tee gpurun_out/timeline_c3_n${n}_$tag.txt
