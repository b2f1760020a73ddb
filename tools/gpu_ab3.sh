#!/bin/bash
# repeated A/B (3x, interleaved) of ONE environment setting on the 200-step packed loop and on the driver's 20-step loop
V="$1"
for rep in 1 2 3; do
  for v in "" "$V"; do
    for a in "--steps 200 --warmup 8" "--steps 20 --warmup 5"; do
      env $v timeout 300 python bench.py $a --no-cpu-baseline --no-e2e --no-extra --no-batch1 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
print('${v:-default}', '| $a |', round(l['value'],1))"
    done
  done
done
