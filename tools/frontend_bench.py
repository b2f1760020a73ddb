#!/usr/bin/env python
"""Timing of ops.preprocess_frame (raw KITTI-shaped rows -> encoder input) next to the CPU oracle (numpy, what the
reference does per frame in its dataloader)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from deeppointmap_b200 import _C, ops  # noqa: E402
from oracle import frontend_ref  # noqa: E402
from test_gpu_frontend import _raw_frame  # noqa: E402

for n in (122000, 65536):
    raw = _raw_frame(1, n)
    d = torch.from_numpy(raw).cuda()
    for _ in range(3):
        out = ops.preprocess_frame(d)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        out = ops.preprocess_frame(d)
    torch.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) / 20 * 1e3
    _C.prof_begin()
    ops.preprocess_frame(d)
    prof = [(t, round(ms * 1e3, 1)) for t, a, b, ms in _C.prof_end() if t != "host_gap"]
    t0 = time.perf_counter()
    want = frontend_ref.preprocess_bin(raw)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    ext = raw[~np.isnan(raw).any(1), :3]
    nvox = int(np.prod(((ext.max(0) - ext.min(0)) / np.float32(0.3)).astype(np.int64) + 1))
    print(f"N={n}: GPU {gpu_ms:.3f} ms per frame incl. the count read-back ({out.shape[1]} points kept, {nvox / 1e6:.1f} M voxel "
          f"table); kernels (us) {prof}; CPU oracle {cpu_ms:.1f} ms; equal: {torch.equal(out.cpu(), want)}", flush=True)
