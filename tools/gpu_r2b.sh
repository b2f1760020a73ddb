#!/bin/bash
# parity suite + bench with the driver's flags + reference arm
TAG=${1:-r02b}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --kernels 30 ) > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.err
python - <<PY
import json
l = json.loads([x for x in open("gpurun_out/${TAG}_bench.log").read().strip().splitlines() if x.startswith("{")][-1])
for k in ("value", "ms_per_step", "e2e", "batch1", "sustained", "strong", "caller_sizes", "index_ops_by_cloud", "reference_gpu", "vs_reference_gpu", "roofline", "kernel_totals_ms_per_step", "kernel_totals_concurrent_ms_per_step", "cpu_baseline"):
    print(k, "=", json.dumps(l.get(k))[:1500])
PY
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/${TAG}_bench_ref.log 2>&1
tail -c 1200 gpurun_out/${TAG}_bench_ref.log
