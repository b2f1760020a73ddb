#!/bin/bash
# stream-count sweep of the headline loop, optionally under environment switches:
#   bash tools/gpu_streams.sh "4 5 6 8" "" "DPM_FPS_T=512 DPM_FPS_PAIR=1"
NS=${1:-"4 5 6 8"}
shift
for v in "$@"; do
  for n in $NS; do
    echo "== ${v:-default} streams=$n"
    env $v timeout 300 python bench.py --steps ${STEPS:-120} --warmup 6 --streams $n --no-cpu-baseline --no-e2e --no-extra --no-batch1 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
print('value', round(l['value'],1), 'ms/step', round(l['ms_per_step'],4))"
  done
done
