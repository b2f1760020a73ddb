#!/bin/bash
# multi-GPU check: gpurun --gpus N -- bash tools/gpu_multi.sh N tag
N=${1:-2}; TAG=${2:-r02m}
mkdir -p gpurun_out
nvidia-smi -L | head -8
( timeout 600 python -m pytest tests/test_gpu_frames.py -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/${TAG}_n${N}_pytest.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/${TAG}_n${N}.err | tail -1 ) > gpurun_out/${TAG}_n${N}_bench.json
python - <<PY
import json
l = json.loads(open("gpurun_out/${TAG}_n${N}_bench.json").read())
for k in ("n_gpus", "value", "ms_per_step", "e2e", "sustained", "strong"):
    print(k, "=", json.dumps(l.get(k))[:600])
PY
tail -3 gpurun_out/${TAG}_n${N}.err
