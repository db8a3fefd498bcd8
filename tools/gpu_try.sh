#!/bin/bash
# retry wrapper around gpurun for "no slot right now" answers (exit 3 / transient): usage tools/gpu_try.sh <timeout> <logfile> <command...>
to=$1; log=$2; shift 2
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$log" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  exit $rc
done
exit 3
