#!/bin/bash
# bench under each value of an environment variable.  gpurun -- bash tools/gpu_envsweep.sh tag VAR "v1 v2 ..." [bench args]
TAG=$1; VAR=$2; VALS=$3; shift; shift; shift
mkdir -p gpurun_out
for v in $VALS; do
  ( env $VAR=$v timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline "$@" 2>&1 | tail -1 ) > gpurun_out/${TAG}_$v.log
  python - <<PY
import json
l = json.loads(open("gpurun_out/${TAG}_$v.log").read().strip().splitlines()[-1])
print("$VAR=$v value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 3), "e2e", l["e2e"] and round(l["e2e"]["value"], 1), "batch1", l.get("batch1") and round(l["batch1"]["ms_per_frame"], 3), {k: v for k, v in list(l["kernel_totals_ms_per_step"].items())[:7]})
PY
done
