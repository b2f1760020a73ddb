"""oracle/outlier_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of OutlierFilter's CUDA branch (/root/reference/dataloader/transforms.py:236-246) with the
kNN of oracle/dpm_oracle.c (pytorch3d knn_points contract).  Pinned by
tests/test_oracle_pin.py::test_outlier_filter_matches_reference (the reference class itself, its CUDA branch forced
onto CPU tensors with a CPU knn_points)."""
import torch

from . import index_ops as IO


def outlier_filter(xyz: torch.Tensor, nb_neighbors: int = 10, std_ratio: float = 3.0):
    """xyz (N,3) -> (kept rows (n,3), mask (N,) bool, statistic (N,), threshold)"""
    p = xyz.float().contiguous().unsqueeze(0)                                  # transforms.py:237
    d2, _ = IO.knn(p, p, None, nb_neighbors + 1)                               # :238
    dists = torch.sqrt(d2.squeeze(0)[:, 1:])                                   # :239
    points_dist = dists.mean(1)                                                # :240
    mean, std = points_dist.mean(), points_dist.std()                          # :241-242
    outlier_dist = mean + std_ratio * std                                      # :243
    mask = points_dist <= outlier_dist                                         # :244
    return xyz[mask], mask, points_dist, float(outlier_dist)
