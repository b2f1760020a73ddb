"""oracle/lowpass_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of LowPassFilter (/root/reference/dataloader/transforms.py:256-297).  The reference takes its normals
from open3d (`estimate_normals(KDTreeSearchParamRadius)`), which this image does not have: `radius_normals` restates
open3d's published algorithm (geometry/EstimateNormals.cpp: covariance of the points within the radius, the query
included; eigenvector of the smallest eigenvalue; fewer than 3 neighbours -> (0, 0, 1)) in fp64 with scipy's kd-tree
-- the same stand-in deeppointmap_b200/compat_shims/open3d provides.  kNN = oracle/dpm_oracle.c (pytorch3d contract).
Pinned by tests/test_oracle_pin.py::test_low_pass_filter_matches_reference: the reference CLASS itself, run on CPU
with that normal stand-in and a CPU knn_points / knn_gather, returns exactly the rows this function keeps."""
import numpy as np
import torch

from . import index_ops as IO


def radius_normals(xyz: torch.Tensor, radius: float) -> torch.Tensor:
    from scipy.spatial import cKDTree
    pts = xyz.double().numpy()
    tree = cKDTree(pts)
    out = np.tile(np.array([0.0, 0.0, 1.0]), (len(pts), 1))
    for i, nb in enumerate(tree.query_ball_point(pts, float(radius))):
        if len(nb) < 3:
            continue
        q = pts[np.asarray(nb)]
        w, v = np.linalg.eigh(np.cov(q.T, bias=True))
        out[i] = v[:, 0]
    return torch.from_numpy(out).float()


def low_pass_filter(xyz: torch.Tensor, normals_radius: float = 0.5, normals_num: int = 16, filter_std: float = 2.0,
                    flux: int = 4, normals: torch.Tensor = None):
    """xyz (N,3) metres -> (kept rows, mask (N,) bool, sim (N,), threshold)"""
    xyz = xyz.float().contiguous()
    n = radius_normals(xyz, normals_radius) if normals is None else normals.float()     # transforms.py:269-272
    _, idx = IO.knn(xyz[None], xyz[None], None, normals_num + 1)                         # :276
    grouped = n[idx[0, :, 1:]]                                                           # :277-278  (N, K, 3)
    similarity = (grouped @ n.unsqueeze(-1)).squeeze(-1).abs()                           # :280
    sim, _ = torch.topk(similarity, k=flux, dim=-1)                                      # :281
    sim = sim.sum(1)                                                                     # :282
    thr = sim.mean() - filter_std * sim.std()
    mask = sim > thr                                                                     # :283
    return xyz[mask], mask, sim, float(thr)
