#!/bin/bash
# stream-count sweep at the driver's 20 steps: automatic FPS mapping, then the packed one
for n in 4 5 6 7 8; do python bench.py --steps 20 --warmup 5 --streams $n --no-cpu-baseline --no-extra --no-e2e --no-batch1 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
print('mode0 streams', l['config']['streams_per_gpu'], round(l['value'],1))"; done
for n in 8 10; do python bench.py --steps 20 --warmup 5 --streams $n --pack-min-steps 1 --no-cpu-baseline --no-extra --no-e2e --no-batch1 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
print('packed streams', l['config']['streams_per_gpu'], round(l['value'],1))"; done
