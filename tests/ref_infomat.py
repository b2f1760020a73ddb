"""Helpers shared by the information-matrix pin test and its golden generator: import the reference's
system/modules/utils.py (build container only) with stub `open3d` / `matplotlib` modules and a CPU
`knn_points`, and the seeded test cases."""
import collections
import math
import os
import sys
import types

import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_knn_points(p1, p2, K=1, return_nn=False, return_sorted=True, **kw):
    """pytorch3d.ops.knn_points contract on CPU: direct-difference squared distances, ascending"""
    d = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2)
    d = (d[..., 0] + d[..., 1]) + d[..., 2]
    dists, idx = torch.topk(d, K, dim=2, largest=False, sorted=True)
    return collections.namedtuple("KNN", "dists idx knn")(dists, idx, None)


def reference_module():
    saved = {k: sys.modules.get(k) for k in ("open3d", "matplotlib", "matplotlib.pyplot", "pytorch3d")}
    for name in ("open3d", "matplotlib", "matplotlib.pyplot"):
        if saved[name] is None:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pytorch3d"] = None  # the module's own import fails -> has_torch3d False; the branch is re-enabled below
    sys.path.insert(0, REF)
    try:
        import importlib
        RU = importlib.import_module("system.modules.utils")
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    RU.has_torch3d = True
    RU.knn_points = _cpu_knn_points
    return RU


def reference_information_matrix(p1, p2, T, radius=1.0):
    RU = reference_module()
    assert radius == 1.0, "the reference hard-codes radius = 1.0 (utils.py:72)"
    return RU.calculate_information_matrix_from_pcd(p1, p2, T, device="cpu")


def _pose(yaw_deg, t):
    a = math.radians(yaw_deg)
    T = torch.eye(4)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = math.cos(a), -math.sin(a), math.sin(a), math.cos(a)
    T[:3, 3] = torch.tensor(t)
    return T


def cases():
    """name -> (p1 (3,N1) metres, p2 (3,N2) metres, SE3, radius)"""
    sys.path.insert(0, ROOT)
    from deeppointmap_b200 import data
    out = {}
    c0 = data.kitti_shape_cloud(11, 6000) * 60.0
    c1, _, _ = data.rigid_move(c0 / 60.0, yaw_deg=1.5, t_m=(0.8, 0.1, 0.0), jitter_m=0.02, seed=12)
    out["kitti6k"] = (c0, (c1 * 60.0)[:, :5500].contiguous(), _pose(1.5, (0.8, 0.1, 0.0)), 1.0)
    g = torch.Generator().manual_seed(5)
    a = torch.rand(3, 1500, generator=g) * 20 - 10
    b = a[:, :900] + 0.3 * torch.randn(3, 900, generator=g)
    out["uniform"] = (a, b.contiguous(), _pose(0.0, (0.0, 0.0, 0.0)), 1.0)
    out["far_apart"] = (a, (b + 100.0).contiguous(), _pose(10.0, (1.0, 2.0, 3.0)), 1.0)  # no correspondence at all
    return out
