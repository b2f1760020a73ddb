#!/bin/bash
# One gpurun call: parity tests, smoke, bench, ncu launch list and full captures of the two
# index kernels.  Usage: gpurun --timeout 1500 -- bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/${TAG}_smi.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/${TAG}_smoke.log
( timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 ) > gpurun_out/${TAG}_bench.log
( timeout 300 python bench.py --steps 10 --warmup 3 --frames 1 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_f1.log
( timeout 300 python bench.py --steps 10 --warmup 3 --frames 8 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_f8.log
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 ) > gpurun_out/${TAG}_bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_grid_kernel -c 1 -f -o gpurun_out/${TAG}_fps \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_fps.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_grid_kernel -c 1 -f -o gpurun_out/${TAG}_knn \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_knn.log 2>&1
cat gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_smoke.log gpurun_out/${TAG}_bench.log
