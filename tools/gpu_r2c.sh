#!/bin/bash
# packed-FPS check: FPS parity on all mappings + the driver's and the default bench invocations, with a mode-0 control
timeout 300 python -m pytest tests/test_gpu_index_ops.py -q -x -k fps 2>&1 | tail -3
for a in "--steps 20 --warmup 5" "--steps 200" "--steps 20 --warmup 5 --fps-mode 0" "--steps 200 --fps-mode 0"; do
  echo "== $a"
  timeout 300 python bench.py $a --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
print('value', round(l['value'],1), 'ms/step', round(l['ms_per_step'],4), 'e2e', round(l['e2e']['value'],1), 'streams', l['config']['streams_per_gpu'], 'batch1', round(l['batch1']['ms_per_frame'],3), l['batch1'].get('pipelined_ms_per_frame'))
print(l['roofline']['fps_mapping'], l['roofline']['launch_ms'], l['roofline']['sms_used'], l['roofline']['us_per_pick'])
print({k:v for k,v in list(l['kernel_totals_ms_per_step'].items())[:6]})"
done
