#!/bin/bash
# per-kernel breakdown of ONE frame per step (batch-1 latency): bash tools/gpu_b1.sh
timeout 300 python bench.py --frames 1 --steps 20 --warmup 5 --streams 1 --no-cpu-baseline --no-e2e --no-extra --kernels 80 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
print('ms/step', round(l['ms_per_step'],4), 'batch1', l['batch1'])
print('profiled_step_ms', l.get('profiled_step_ms'), 'launches/step', l['launches_per_step'])
print(l['kernel_totals_ms_per_step'])
for k in l['kernels_ms_per_step'][:45]: print('  ', k)"
