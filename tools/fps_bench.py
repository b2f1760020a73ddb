"""FPS latency per pick on the two mappings (one SM per cloud / 8-CTA cluster per cloud), KITTI-shape and the
adversarial uniform cube.  python tools/fps_bench.py [out.json]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from deeppointmap_b200 import data, ops  # noqa: E402

dev = "cuda:0"
from deeppointmap_b200 import _C  # noqa: E402
torch.zeros(1, device=dev)
print("cluster capacity (clouds):", _C.lib().dpm_fps_cluster_capacity(), flush=True)
rows = []
for kind in ("kitti", "cube"):
    for n, k, B in ((65536, 4096, 1), (65536, 4096, 4), (65536, 4096, 8), (65536, 4096, 12), (65536, 4096, 16), (16384, 4096, 1), (4096, 1024, 1),
                    (4096, 1024, 32), (1024, 256, 1), (256, 64, 1), (131072, 4096, 1)):
        mk = data.kitti_shape_cloud if kind == "kitti" else data.uniform_cube_cloud
        pts = torch.stack([mk(s + 1, n).T.contiguous() for s in range(B)]).to(dev)
        ref = None
        for mode in (1, 2):
            ops.set_fps_mode(mode)
            for _ in range(2):
                _, idx = ops.sample_farthest_points(pts, K=k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                _, idx = ops.sample_farthest_points(pts, K=k)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            same = True if ref is None else bool(torch.equal(ref, idx))
            ref = idx if ref is None else ref
            rows.append({"cloud": kind, "N": n, "K": k, "B": B, "mode": mode, "ms": round(ms, 4),
                         "us_per_pick": round(1e3 * ms / (k - 1), 4), "same_as_mode1": same})
            print(rows[-1], flush=True)
ops.set_fps_mode(0)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
