#!/usr/bin/env python
"""Timing of ops.information_matrix (system/modules/utils.py:60-104) next to the reference's own formulation
on the same GPU (our knn_points + the reference's mask / outer-product / .cpu() sequence) and the CPU oracle."""
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deeppointmap_b200 import data, ops  # noqa: E402
from oracle import infomat_ref  # noqa: E402


def reference_formulation(pcd1, pcd2, SE3):
    """the reference's pytorch3d branch, line for line in behaviour, with knn_points = ours"""
    R, T = SE3[:3, :3].cuda(), SE3[:3, 3:].cuda()
    p1 = (R @ pcd1 + T).T.unsqueeze(0)
    p2 = pcd2.T.unsqueeze(0)
    res = ops.knn_points(p1, p2, K=1, return_nn=False, return_sorted=False)
    idx, dists = res.idx.squeeze(0).squeeze(-1), res.dists.squeeze(0).squeeze(-1)
    t = pcd2[:, idx[dists <= 1.0]].T
    x, y, z = t[:, 0], t[:, 1], t[:, 2]
    GTG = torch.zeros(6, 6, device="cuda")
    for cols in (((1, z), (2, -y), (3, None)), ((0, -z), (2, x), (4, None)), ((0, y), (1, -x), (5, None))):
        G = torch.zeros(t.shape[0], 6, 1, device="cuda")
        for c, v in cols:
            G[:, c, 0] = 1.0 if v is None else v
        GTG += (G @ G.transpose(1, 2)).sum(0)
    return GTG.cpu()


for n in (16384, 65536):
    c0 = data.kitti_shape_cloud(31, n) * 60.0
    c1, _, _ = data.rigid_move(c0 / 60.0, yaw_deg=2.0, t_m=(1.0, 0.1, 0.0), jitter_m=0.02, seed=32)
    c1 = c1 * 60.0
    a = math.radians(2.0)
    T = torch.eye(4)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = math.cos(a), -math.sin(a), math.sin(a), math.cos(a)
    T[:3, 3] = torch.tensor([1.0, 0.1, 0.0])
    d0, d1, Td = c0.cuda(), c1.cuda(), T.cuda()
    for name, fn in (("fused dpm_information_matrix_f32 (+ .cpu())", lambda: ops.information_matrix(d0, d1, Td).cpu()),
                     ("reference formulation on our knn_points", lambda: reference_formulation(d0, d1, T))):
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            out = fn()
        torch.cuda.synchronize()
        print(f"N={n}: {name}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per edge", flush=True)
    t0 = time.perf_counter()
    want, cnt = infomat_ref.information_matrix(c0, c1, T)
    print(f"N={n}: CPU oracle {1e3 * (time.perf_counter() - t0):.1f} ms; {cnt} correspondences; fused vs oracle rel err "
          f"{float((ops.information_matrix(d0, d1, Td).cpu() - want).abs().max() / want.abs().max()):.1e}", flush=True)
