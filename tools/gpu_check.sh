#!/bin/bash
# One gpurun call: parity tests, smoke, bench (+ reference arm), ncu launch list and full captures of the top
# kernels.  Usage: gpurun --timeout 1700 -- bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/${TAG}_smi.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/${TAG}_smoke.log
( timeout 600 python bench.py --steps 10 --warmup 3 --kernels 80 2>&1 | tail -3 ) > gpurun_out/${TAG}_bench.log
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --streams 1 --no-cpu-baseline --no-e2e --no-batch1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
bash tools/gpu_ncu.sh ${TAG} 32 'fps_grid_kernel<.int.2' 'knn_grid_kernel' 'linear_tc_kernel<.int.256, .int.2, .bool.1' \
    'attention_tc_kernel' 'group_lane_kernel<.int.32, .bool.0'
cat gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_smoke.log gpurun_out/${TAG}_bench.log
