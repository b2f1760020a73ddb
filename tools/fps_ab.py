#!/usr/bin/env python
"""A/B timing of the FPS kernel variants (DPM_FPS_MODE) on the bench's clouds.  One process per mode."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ROOT)
    from deeppointmap_b200 import data, ops
    for (n, k, B) in [(65536, 4096, 32), (65536, 4096, 1), (4096, 1024, 32), (16384, 4096, 8)]:
        pts = torch.stack([data.kitti_shape_cloud(s, n).T.contiguous() for s in range(B)]).cuda()
        ref = None
        ts = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, idx = ops.sample_farthest_points(pts, None, k)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"  mode {os.environ.get('DPM_FPS_MODE', '0')} N={n} K={k} B={B}: {min(ts):.3f} ms  (checksum {int(idx.sum())})", flush=True)
else:
    for mode in sys.argv[1:] or ["0", "1", "2"]:
        env = dict(os.environ, DPM_FPS_MODE=mode)
        subprocess.run([sys.executable, __file__, "child"], env=env)
