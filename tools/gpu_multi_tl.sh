#!/bin/bash
# N-GPU visit: steady-state and isolated (L2-flushed) kernel timelines of rank 0 under rollout sharding, then the C3 bench line
n=${1:-2}; tag=${2:-x}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29541 tools/timeline.py c3 40 2>&1 | grep -v "^\*\*\*\|^$\|NCCL version\|Setting OMP" | tee gpurun_out/timeline_c3_n${n}_$tag.txt
run 29542 tools/timeline.py c3 20 flush 2>&1 | grep -v "^\*\*\*\|^$\|NCCL version\|Setting OMP" | tee gpurun_out/timeline_c3_n${n}_flush_$tag.txt
run 29544 bench.py --gpus $n --steps 20 --warmup 5 --skip-cpu-baseline --skip-c4 > gpurun_out/scale_c3_n${n}_$tag.json 2> gpurun_out/scale_c3_n${n}_$tag.err; echo "n$n rc=$?"
python - gpurun_out/scale_c3_n${n}_$tag.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','graph_replays','parity_ok')}, 'steady', d['steady_state']['ms_per_step'], 'e2e', d['e2e']['value'])
PY
