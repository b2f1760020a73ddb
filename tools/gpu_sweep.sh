#!/bin/bash
# streams x frames sweep of the device-resident bench.  gpurun --timeout 900 -- bash tools/gpu_sweep.sh tag
TAG=${1:-sw}
mkdir -p gpurun_out
for cfg in "4 32" "6 32" "8 32" "2 64" "3 64" "4 64" "2 128"; do
  set -- $cfg
  ( timeout 300 python bench.py --steps 12 --warmup 4 --streams $1 --frames $2 --no-cpu-baseline --no-batch1 2>&1 | tail -1 ) > gpurun_out/${TAG}_s$1_f$2.log
  python - <<PY
import json
l = json.loads(open("gpurun_out/${TAG}_s$1_f$2.log").read().strip().splitlines()[-1])
print("streams $1 frames $2 value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 3), "e2e", l["e2e"] and round(l["e2e"]["value"], 1))
PY
done
