"""oracle/infomat_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of calculate_information_matrix_from_pcd, the pytorch3d branch
(/root/reference/system/modules/utils.py:60-104): transform the source cloud, 1-NN into the
target (pytorch3d knn_points contract: direct-difference fp32 distances, ties to the lower index --
oracle/dpm_oracle.c), keep d2 <= radius**2, and sum G G^T over the matched TARGET points for the
three Jacobian rows the reference writes out (utils.py:87-101).

Pinned by tests/test_oracle_pin.py::test_information_matrix_matches_reference (the reference
function itself, run in the build container with a CPU knn_points) and tests/golden/infomat.npz.
"""
import torch

from . import index_ops as IO


def information_matrix(pointcloud_1: torch.Tensor, pointcloud_2: torch.Tensor, SE3: torch.Tensor, radius: float = 1.0):
    """(3,N1), (3,N2), (4,4) -> (6,6) float32, number of correspondences"""
    R, T = SE3[:3, :3].float(), SE3[:3, 3:].float()                       # PoseTool.Rt, utils.py:42-50
    p1 = (R @ pointcloud_1.float() + T).T.contiguous().unsqueeze(0)       # utils.py:77
    p2 = pointcloud_2.float().T.contiguous().unsqueeze(0)                 # utils.py:78
    dists, idx = IO.knn(p1, p2, None, 1)                                  # utils.py:80
    idx, dists = idx[0, :, 0], dists[0, :, 0]
    corres = idx[dists <= (radius ** 2)]                                  # utils.py:83-85
    t = pointcloud_2.float()[:, corres].T                                 # utils.py:87
    x, y, z = t[:, 0], t[:, 1], t[:, 2]
    one, zero = torch.ones_like(x), torch.zeros_like(x)
    GTG = torch.zeros(6, 6)
    for row in ((zero, z, -y, one, zero, zero), (-z, zero, x, zero, one, zero), (y, -x, zero, zero, zero, one)):
        G = torch.stack(row, dim=1).unsqueeze(-1)                         # (n,6,1), utils.py:90-101
        GTG += (G @ G.transpose(1, 2)).sum(0)
    return GTG, int(corres.numel())
