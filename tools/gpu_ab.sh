#!/bin/bash
# A/B of an environment switch on the headline loop: bash tools/gpu_ab.sh VAR=val [VAR=val ...]
for v in "" "$@"; do
  echo "== ${v:-default}"
  env $v timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
print('value', round(l['value'],1), 'ms/step', round(l['ms_per_step'],4), 'batch1 ms', round(l['batch1']['ms_per_frame'],3), 'graph', l['batch1'].get('cuda_graph_ms_per_frame'))
print({k:v for k,v in list(l['kernel_totals_ms_per_step'].items())[:8]})"
done
