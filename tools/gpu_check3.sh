#!/bin/bash
# Final round-2 measurement campaign (packed FPS mapping included) in one gpurun call: parity tests, smoke, bench (driver flags) + reference arm, ncu launch
# list, full captures of the top kernels at 32 frames (throughput mapping) and 1 frame (latency mapping).
#   gpurun --timeout 2400 -- bash tools/gpu_check2.sh r02h
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/${TAG}_smi.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/${TAG}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/${TAG}_smoke.log
( timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --kernels 80 2>gpurun_out/${TAG}_bench.err | tail -1 ) > gpurun_out/${TAG}_bench.json
( timeout 900 python bench.py --gpus 1 --kernels 80 --no-extra 2>/dev/null | tail -1 ) > gpurun_out/${TAG}_bench_default200.json
( timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 2>/dev/null | tail -1 ) > gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --streams 1 --no-cpu-baseline --no-e2e --no-batch1 --no-extra > gpurun_out/${TAG}_ncu_bench.log 2>&1
bash tools/gpu_ncu.sh ${TAG} 32 'fps_grid_kernel<.int.2, .int.1024' 'knn_grid_kernel' 'linear_tc_kernel<.int.256, .int.2, .bool.1, .bool.0' 'linear_tc_kernel<.int.128, .int.1, .bool.1, .bool.1' \
    'linear_tc_kernel<.int.256, .int.2, .bool.1, .bool.1' 'attention_tc5_kernel' 'group_lane_kernel<.int.32, .bool.0' 'pair_select_kernel'
bash tools/gpu_ncu.sh ${TAG}b1 1 'fps_grid_cluster_kernel<.int.2'
BENCH_EXTRA='--pack-min-steps 1' bash tools/gpu_ncu.sh ${TAG}pk 32 'fps_grid_kernel<.int.2, .int.512'
cat gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_smoke.log
python - <<PY
import json
l = json.loads(open("gpurun_out/${TAG}_bench.json").read())
for k in ("value", "ms_per_step", "e2e", "batch1", "sustained", "strong", "vs_reference_gpu", "pipeline_infer", "kernel_totals_ms_per_step"):
    print(k, "=", json.dumps(l.get(k))[:900])
print(open("gpurun_out/${TAG}_bench_reference.json").read()[:600])
PY
ls -la gpurun_out | grep ${TAG} | head -40
