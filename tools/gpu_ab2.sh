#!/bin/bash
# A/B of environment switches on the 200-step headline loop (packed FPS, 8 streams): bash tools/gpu_ab2.sh VAR=val ...
for v in "" "$@"; do
  echo "== ${v:-default}"
  env $v timeout 300 python bench.py --steps 200 --warmup 8 --no-cpu-baseline --no-e2e --no-extra --no-batch1 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
print('value', round(l['value'],1), 'ms/step', round(l['ms_per_step'],4), 'streams', l['config']['streams_per_gpu'])
print({k:v for k,v in list(l['kernel_totals_ms_per_step'].items())[:7]})"
done
