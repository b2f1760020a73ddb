#!/bin/bash
mkdir -p gpurun_out
for a in "--steps 20 --warmup 5 --pipeline-frames 0" "--steps 200 --no-extra"; do
  echo "== $a"
  timeout 600 python bench.py $a --no-cpu-baseline 2>gpurun_out/r2d_err.log | tee gpurun_out/r2d_last.json | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1])
print('value', round(l['value'],1), 'ms/step', round(l['ms_per_step'],4), 'e2e', round(l['e2e']['value'],1), 'streams', l['config']['streams_per_gpu'], l['config']['fps_mode'][:40])
s=l.get('sustained')
print('sustained', s and (round(s['value'],1), s['steps'], s['streams_per_gpu'], s['fps_mode'][:12]))
print(l['roofline']['fps_mapping'][:40], l['roofline']['launch_ms'], l['roofline']['sms_used'])
print('strong', l.get('strong') and l['strong']['ms_per_step'], 'batch1', l['batch1']['ms_per_frame'])"
  tail -3 gpurun_out/r2d_err.log
done
