#!/bin/bash
# compute-sanitizer memcheck + racecheck over a small selection of the GPU tests.
#   gpurun --timeout 1500 -- bash tools/gpu_sanitize.sh
mkdir -p gpurun_out
SEL='tests/test_gpu_dense.py::test_linear_layernorm_fused tests/test_gpu_decoder.py::test_attention_core tests/test_gpu_frontend.py::test_frontend_golden_other_parameters_and_limits tests/test_gpu_encoder.py::test_small_config_random_weights_batch tests/test_gpu_decoder.py::test_registration_random_weights_batched_matches_single'
for tool in ${1:-memcheck racecheck}; do
  ( timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest $SEL tests/test_gpu_index_ops.py -m gpu -x -q \
      -k "fused and 1000 or attention_core or frontend_golden or small_config_random or batched_matches_single or hybrid_grid_ties or information_matrix_golden or (fps_bit_exact and 4096)" 2>&1 | grep -v "Host Frame" | tail -60 ) > gpurun_out/sanitize_$tool.log
  echo "== $tool"; tail -12 gpurun_out/sanitize_$tool.log | cut -c1-300
done
