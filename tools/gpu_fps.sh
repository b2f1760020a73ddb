#!/bin/bash
# cluster-FPS validation: parity in both mappings, per-pick latency, phase profile, batch-1 bench.
#   gpurun --timeout 1200 -- bash tools/gpu_fps.sh tag
TAG=${1:-f}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_index_ops.py -m gpu -x -q -k "fps or deterministic" 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest_fps.log
cat gpurun_out/${TAG}_pytest_fps.log
( timeout 400 python tools/fps_bench.py gpurun_out/${TAG}_fps_bench.json 2>&1 | grep -v "'mode': 1" | tail -45 ) > gpurun_out/${TAG}_fps_bench.log
cat gpurun_out/${TAG}_fps_bench.log
if [ -f deeppointmap_b200/libdpm_prof.so ]; then
  ( DPM_LIB=$PWD/deeppointmap_b200/libdpm_prof.so timeout 120 python tools/fps_profile.py 2>&1 | tail -40 ) > gpurun_out/${TAG}_fps_profile.log
  cat gpurun_out/${TAG}_fps_profile.log
fi
( timeout 300 python bench.py --steps 10 --warmup 3 --frames 1 --streams 1 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_f1.log
python - <<PY
import json
l = json.loads(open("gpurun_out/${TAG}_bench_f1.log").read().strip().splitlines()[-1])
print("frames=1 value", l["value"], "ms/step", l["ms_per_step"], "batch1", l.get("batch1"))
print(l.get("kernel_totals_ms_per_step"))
PY
