#!/bin/bash
# A/B of environment settings on the driver's 20-step loop and the 200-step loop, with the kernel totals
for v in "" "$@"; do
  for a in "--steps 20 --warmup 5" "--steps 200 --warmup 8"; do
    env $v timeout 300 python bench.py $a --no-cpu-baseline --no-e2e --no-extra --no-batch1 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
t=l['kernel_totals_ms_per_step']
print('${v:-default}', '| $a |', round(l['value'],1), {k:t[k] for k in ('fps','knn','group') if k in t})"
  done
done
