#!/bin/bash
# refresh of the bench lines + ncu launch list on the final code (the full ncu captures of tools/gpu_check3.sh stay valid:
# those kernels did not change):   gpurun --timeout 1500 -- bash tools/gpu_final.sh r02i
TAG=${1:-r02i}
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/${TAG}_smoke.log
( timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --kernels 80 2>gpurun_out/${TAG}_bench.err | tail -1 ) > gpurun_out/${TAG}_bench.json
( timeout 900 python bench.py --gpus 1 --kernels 80 --no-extra 2>/dev/null | tail -1 ) > gpurun_out/${TAG}_bench_default200.json
( timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 2>/dev/null | tail -1 ) > gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --streams 1 --no-cpu-baseline --no-e2e --no-batch1 --no-extra > gpurun_out/${TAG}_ncu_bench.log 2>&1
cat gpurun_out/${TAG}_smoke.log
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_bench_default200.json"):
    l = json.loads(open(f).read())
    print(f, round(l["value"], 1), "e2e", round(l["e2e"]["value"], 1), "streams", l["config"]["streams_per_gpu"], "batch1", l["batch1"])
    print("  sustained", l.get("sustained") and round(l["sustained"]["value"], 1), "strong", l.get("strong") and l["strong"]["ms_per_step"])
print(open("gpurun_out/${TAG}_bench_reference.json").read()[:300])
PY
