#!/bin/bash
# compute-sanitizer memcheck + racecheck over the kernels of the final session: packed FPS (two teams with named barriers),
# FPS with 16 points per lane, multi-pass kNN (K > 32), narrow-tile GEMMs + ln_fused_order_kernel, bias-free encoder.
#   gpurun --timeout 1500 -- bash tools/gpu_sanitize3.sh
mkdir -p gpurun_out
for tool in ${1:-memcheck racecheck}; do
  ( timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_index_ops.py tests/test_gpu_dense.py tests/test_gpu_encoder.py -m gpu -x -q \
      -k "(fps_bit_exact and packed-4096) or (fps_bit_exact and packed-65536) or (fps_lengths and packed) or (fps_duplicates and packed) or (beyond and one-sm-300000) or (more_than_32 and not 70000) or few_rows or bias_false" 2>&1 | grep -v "Host Frame" | tail -40 ) > gpurun_out/sanitize3_$tool.log
  echo "== $tool"; tail -12 gpurun_out/sanitize3_$tool.log | cut -c1-300
done
