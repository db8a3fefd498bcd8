#!/bin/bash
# programmatic dependent launch, edge by edge: steady-state timeline (no graph: the timeline switches it off) and the bench's steady / isolated figures
tag=${1:-x}; mkdir -p gpurun_out
for m in 0 1 2 4 5 7; do
  echo "=== STOMP_B200_PDL=$m" | tee -a gpurun_out/pdl_$tag.txt
  STOMP_B200_PDL=$m timeout 300 python tools/timeline.py c3 40 2>&1 | grep -E "per iteration|median  |gaps|gap update" | tee -a gpurun_out/pdl_$tag.txt
  STOMP_B200_PDL=$m timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-c4 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('  bench: value ms', round(d['ms_per_step']*1e3,1), 'steady ms', round(d['steady_state']['ms_per_step']*1e3,1), 'e2e G', round(d['e2e']['value']/1e9,2))" | tee -a gpurun_out/pdl_$tag.txt
done
