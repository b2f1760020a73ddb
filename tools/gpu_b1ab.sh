#!/bin/bash
# batch-1 latency A/B of environment switches: bash tools/gpu_b1ab.sh VAR=val ...
for v in "" "$@"; do
  echo "== ${v:-default}"
  env $v timeout 300 python bench.py --frames 1 --steps 20 --warmup 5 --streams 1 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
b=l['batch1']
print('batch1 ms', round(b['ms_per_frame'],3), 'graph', round(b['cuda_graph_ms_per_frame'],3), 'pipelined', round(b['pipelined_ms_per_frame'],3))
print({k:v for k,v in list(l['kernel_totals_ms_per_step'].items())[:7]})"
done
