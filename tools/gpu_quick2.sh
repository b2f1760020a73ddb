#!/bin/bash
# all GPU tests + default bench.  gpurun --timeout 1200 -- bash tools/gpu_quick2.sh tag
TAG=${1:-q}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log | cut -c1-400
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernels 80 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench.log
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.log",):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 3), "e2e", l["e2e"] and round(l["e2e"]["value"], 1))
        print("   ", l.get("kernel_totals_ms_per_step"))
        for k in l["kernels_ms_per_step"][:80]:
            print("   ", k)
    except Exception as e:
        print(f, "unreadable", e, open(f).read()[-1500:])
PY
