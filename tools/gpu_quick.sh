#!/bin/bash
# quick iteration: parity tests + a short bench.  gpurun --timeout 900 -- bash tools/gpu_quick.sh tag [pytest-args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -25 ) > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/${TAG}_bench.log
( timeout 300 python bench.py --steps 10 --warmup 3 --streams 1 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_s1.log
( timeout 300 python bench.py --steps 12 --warmup 3 --streams 4 --frames 8 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_s4f8.log
( timeout 300 python bench.py --steps 10 --warmup 3 --frames 1 --streams 1 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_f1.log
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.log", "gpurun_out/${TAG}_bench_s1.log", "gpurun_out/${TAG}_bench_s4f8.log", "gpurun_out/${TAG}_bench_f1.log"):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 3), "e2e", l["e2e"] and round(l["e2e"]["value"], 1))
        print("   ", l.get("kernel_totals_ms_per_step"))
        for k in l["kernels_ms_per_step"][:16]:
            print("   ", k)
    except Exception as e:
        print(f, "unreadable", e, open(f).read()[-1500:])
PY
