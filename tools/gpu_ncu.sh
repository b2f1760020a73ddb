#!/bin/bash
# full ncu capture of kernels whose DEMANGLED name matches the given regexes (template arguments included).
#   gpurun -- bash tools/gpu_ncu.sh tag frames 'fps_grid_kernel' 'linear_tc_kernel<256' ...
TAG=$1; FR=$2; shift; shift
mkdir -p gpurun_out
for RX in "$@"; do
  NAME=$(echo "$RX" | tr -cd 'a-zA-Z0-9_')
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$RX" -c 1 -f \
    -o gpurun_out/${TAG}_${NAME} python bench.py --steps 1 --warmup 1 --frames $FR --streams 1 --no-cpu-baseline --no-e2e --no-extra $BENCH_EXTRA \
    > gpurun_out/${TAG}_ncu_${NAME}.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_${NAME}.log | cut -c1-200
done
