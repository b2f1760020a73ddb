"""oracle/frontend_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement (numpy, as the reference) of the neighbourhood-free part of the inference YAML's transform
chain, paths relative to /root/reference:
  dataloader/heads/bin.py:16-17          float32 (N,4) -> xyz, NaN rows dropped
  dataloader/transforms.py:331-356       VoxelSample(voxel_size, 'first')
  dataloader/transforms.py:387-397       DistanceSample(min, max)
  dataloader/transforms.py:400-407       CoordinatesNormalization(ratio)
(the open3d / pytorch3d OutlierFilter and LowPassFilter of the shipped YAML are not part of it).

Pinned by tests/test_oracle_pin.py::test_frontend_matches_reference_transforms (the reference's own transform
classes run on a sample frame in the build container) and tests/golden/frontend.npz.
"""
import numpy as np
import torch


def preprocess_bin(raw: np.ndarray, voxel_size: float = 0.3, min_dis: float = 1.0, max_dis: float = 60.0,
                   ratio: float = 60.0) -> torch.Tensor:
    """raw float32 (N,4) KITTI frame -> (3, n) fp32 normalised cloud (reduced transform chain:
    VoxelSample('first') -> DistanceSample -> CoordinatesNormalization; the open3d/pytorch3d
    OutlierFilter and LowPassFilter of the shipped YAML are skipped)."""
    xyz = np.asarray(raw, dtype=np.float32).reshape(-1, 4)[:, :3]
    xyz = xyz[np.isnan(xyz).sum(1) == 0]
    lo, hi = xyz.min(axis=0), xyz.max(axis=0)
    X, Y, _ = ((hi - lo) / voxel_size).astype(np.int32) + 1
    v = ((xyz - lo) / voxel_size).astype(np.int32)
    vid = (v[:, 0] + v[:, 1] * X + v[:, 2] * X * Y).astype(np.int32)
    _, first = np.unique(vid, return_index=True)
    pts = torch.from_numpy(xyz[first])
    d = torch.norm(pts, p=2, dim=1)
    pts = pts[(min_dis <= d) & (d <= max_dis)]
    pts = pts / ratio
    return pts.T.contiguous()
