#!/bin/bash
# round-2 state check: full GPU parity suite, FPS mappings, bench line.
TAG=${1:-r02a}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
( timeout 400 python tools/fps_bench.py gpurun_out/${TAG}_fps_bench.json 2>&1 | tail -50 ) > gpurun_out/${TAG}_fps_bench.log
cat gpurun_out/${TAG}_fps_bench.log
( timeout 600 python bench.py --steps 20 --warmup 3 --kernels 80 2>&1 | tail -3 ) > gpurun_out/${TAG}_bench.log
cat gpurun_out/${TAG}_bench.log
