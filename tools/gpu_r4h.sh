#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r4h.log 2>&1; tail -3 $O/pytest_gpu_r4h.log
{
for cfg in "X=1" "STOMP_B200_STATES_PREFETCH=0" "STOMP_B200_STATES_PREFETCH=4" "STOMP_B200_STATES_MIN_BLOCKS=7"; do
  echo "=== $cfg"
  env $cfg python tools/timeline.py c3 40 | grep -E "cost  |period"
  env $cfg python tools/timeline.py c3 20 flush | grep -E "cost  |period"
done
} > $O/state_prefetch_variants_r4h.txt 2>&1
cat $O/state_prefetch_variants_r4h.txt
python tools/e2e_breakdown.py c3 60 20 8 > $O/e2e_breakdown_r4h.txt 2>&1; cat $O/e2e_breakdown_r4h.txt
python tools/e2e_breakdown.py c4 12 20 8 > $O/e2e_breakdown_c4_r4h.txt 2>&1; cat $O/e2e_breakdown_c4_r4h.txt
for w in c2 c4 c5; do timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --skip-c4 > $O/bench_${w}_r4h.json 2> $O/bench_${w}_r4h.err; echo "$w rc=$?"; done
python - <<'PY'
import json
for n in ('c2','c4','c5'):
    d=json.loads([l for l in open(f'gpurun_out/bench_{n}_r4h.json') if l.startswith('{')][-1])
    print(n, 'value %.2fG'%(d['value']/1e9), 'steady %.2fG'%(d['steady_state']['value']/1e9), 'e2e %.2fG'%(d['e2e']['value']/1e9), 'roof', round(d['roofline']['frac'],3), 'span', d['roofline'].get('kernel_span'))
PY
