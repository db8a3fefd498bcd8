#!/bin/bash
# N GPUs on the round's final code: the C3 bench line under torchrun (sharded-vs-single parity_ok inside, query-sharded C4 beside it)
cd "$(dirname "$0")/.."
n=${1:-8}; tag=${2:-r5p}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/scale_c3_n${n}_$tag.json 2> gpurun_out/scale_c3_n${n}_$tag.err; echo "n$n rc=$?"
python - gpurun_out/scale_c3_n${n}_$tag.json <<'PY'
import json,sys
try:
    l=[x for x in open(sys.argv[1]).read().splitlines() if x.startswith('{')][-1]; d=json.loads(l)
    print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','parity_ok')}, d['steady_state']['value'], d['e2e']['value'], d['kernel_ms_per_step'], d['graph_replays'], (d.get('query_sharded_c4') or {}).get('value'))
except Exception as e: print('ERR',e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
