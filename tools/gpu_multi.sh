#!/bin/bash
# N-GPU visit: sharded-vs-single parity check under torchrun with N ranks, then the C3 bench at N (rollout sharding)
n=${1:-2}; tag=${2:-x}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/scale_c3_n${n}_$tag.json 2> gpurun_out/scale_c3_n${n}_$tag.err; echo "n$n rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $n --steps 10 --warmup 3 --skip-cpu-baseline --workload c4 > gpurun_out/scale_c4_n${n}_$tag.json 2> gpurun_out/scale_c4_n${n}_$tag.err; echo "c4 n$n rc=$?"
for f in gpurun_out/scale_*_n${n}_$tag.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    l=[x for x in open(sys.argv[1]).read().splitlines() if x.startswith('{')][-1]; d=json.loads(l)
    print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['steady_state']['value'], d['e2e']['value'], d['kernel_ms_per_step'])
except Exception as e: print('ERR',e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
