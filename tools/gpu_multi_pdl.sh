#!/bin/bash
# N-GPU: steady-state timelines of rank 0 under PDL masks (rollout sharding)
n=${1:-2}; tag=${2:-x}; shift 2
mkdir -p gpurun_out
for m in "$@"; do
  echo "=== STOMP_B200_PDL=$m" | tee -a gpurun_out/multi_pdl_n${n}_$tag.txt
  STOMP_B200_PDL=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 tools/timeline.py c3 40 2>&1 | grep -E "per iteration|median  |gaps|gap update|period" | tee -a gpurun_out/multi_pdl_n${n}_$tag.txt
done
