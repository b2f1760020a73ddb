#!/usr/bin/env python
"""stand-alone FPS (no radius bound on the cell size) against the grid's target points per cell: DPM_GRID_PPC=.. python tools/fps_ppc.py"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeppointmap_b200 import data, ops
for kind in ("kitti", "cube"):
    gen = data.kitti_shape_cloud if kind == "kitti" else data.uniform_cube_cloud
    for B in (1, 32):
        pts = torch.stack([gen(s, 65536).T.contiguous() for s in range(B)]).cuda()
        for _ in range(2):
            ops.sample_farthest_points(pts, K=4096)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.sample_farthest_points(pts, K=4096)
        e1.record()
        torch.cuda.synchronize()
        print(f"ppc={os.environ.get('DPM_GRID_PPC', '32')} {kind} B={B}: {e0.elapsed_time(e1) / 5:.3f} ms per call", flush=True)
