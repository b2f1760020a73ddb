import sys, time, torch
sys.path.insert(0, "/root/repo")
from deeppointmap_b200 import data, ops
for n in (16384, 46000):
    p = (data.kitti_shape_cloud(1, n) * 60).T.contiguous().cuda()[None]
    for K in (12, 18):
        for _ in range(2): r = ops.knn_points(p, p, K=K)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): r = ops.knn_points(p, p, K=K)
        torch.cuda.synchronize()
        print(f"self-kNN N={n} K={K}: {(time.perf_counter()-t0)/5*1e3:.2f} ms")
