#!/bin/bash
# compute-sanitizer memcheck + racecheck over the kernels added in round 2: cluster FPS (st.async / mbarrier / DSMEM),
# pairing (last-block pattern), LowPassFilter, key-padding masks, tcgen05 attention, bulk-copy weight tiles.
#   gpurun --timeout 1800 -- bash tools/gpu_sanitize2.sh
mkdir -p gpurun_out
for tool in ${1:-memcheck racecheck}; do
  ( timeout 800 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_frontend.py tests/test_gpu_index_ops.py tests/test_gpu_dense.py -m gpu -x -q \
      -k "key_padding or (attention_pairs and 130) or (attention_pairs and 33) or (low_pass and 1500) or real_weights_synthetic_golden or (fps_bit_exact and cluster-4096) or (fps_bit_exact and cluster-14500) or (fps_duplicates and cluster) or (fps_grid_ties and cluster-5000) or (fused and 1000)" 2>&1 | grep -v "Host Frame" | tail -60 ) > gpurun_out/sanitize2_$tool.log
  echo "== $tool"; tail -14 gpurun_out/sanitize2_$tool.log | cut -c1-300
done
