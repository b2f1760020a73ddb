"""Deterministic synthetic inputs for the hot path and the reduced
numpy-only preprocessing of a raw KITTI .bin frame (config 1).

Reference behaviour mirrored by `preprocess_bin` (paths relative to /root/reference):
  dataloader/heads/bin.py:16-17          float32 (N,4) -> xyz, NaN rows dropped
  dataloader/transforms.py:331-356       VoxelSample(voxel_size, 'first')
  dataloader/transforms.py:393-397       DistanceSample(min, max)
  dataloader/transforms.py:400-407       CoordinatesNormalization(ratio)
"""
import math

import numpy as np
import torch


def kitti_shape_cloud(seed: int, n: int = 65536, scale: float = 60.0) -> torch.Tensor:
    """'KITTI-shape' cloud: 70 % ground disc, 30 % vertical facades, 1 m <= |p| <= 60 m,
    shuffled, divided by `scale`.  Returns (3, n) fp32 (channel-first, like ToTensor)."""
    g = torch.Generator().manual_seed(int(seed))
    m = int(n * 1.6) + 1024

    def u(*shape):
        return torch.rand(*shape, generator=g, dtype=torch.float64)

    ng = int(m * 0.7)
    r = torch.sqrt(u(ng) * (60.0 ** 2 - 1.0) + 1.0)
    th = 2 * math.pi * u(ng)
    ground = torch.stack([r * torch.cos(th), r * torch.sin(th),
                          -1.73 + 0.02 * torch.randn(ng, generator=g, dtype=torch.float64)], dim=1)
    nf = m - ng
    planes = 64
    centre = u(planes, 2) * 100.0 - 50.0
    heading = u(planes) * math.pi
    length = 5.0 + 20.0 * u(planes)
    pid = torch.randint(0, planes, (nf,), generator=g)
    along = (u(nf) - 0.5) * length[pid]
    fx = centre[pid, 0] + along * torch.cos(heading[pid])
    fy = centre[pid, 1] + along * torch.sin(heading[pid])
    fz = -1.73 + 6.0 * u(nf)
    facade = torch.stack([fx, fy, fz], dim=1)
    pts = torch.cat([ground, facade], dim=0)
    d = pts.norm(dim=1)
    pts = pts[(d >= 1.0) & (d <= 60.0)]
    perm = torch.randperm(pts.shape[0], generator=g)
    pts = pts[perm]
    if pts.shape[0] < n:  # top up by resampling with jitter (never hit for the default margins)
        extra = pts[torch.randint(0, pts.shape[0], (n - pts.shape[0],), generator=g)]
        pts = torch.cat([pts, extra + 0.01 * torch.randn(extra.shape, generator=g, dtype=torch.float64)], dim=0)
    pts = pts[:n]
    return (pts / scale).to(torch.float32).T.contiguous()


def uniform_cube_cloud(seed: int, n: int) -> torch.Tensor:
    """No-structure adversarial case: uniform in [-1,1]^3.  (3, n) fp32."""
    g = torch.Generator().manual_seed(int(seed))
    return (torch.rand(3, n, generator=g, dtype=torch.float32) * 2.0 - 1.0).contiguous()


def rigid_move(cloud: torch.Tensor, yaw_deg: float, t_m, scale: float = 60.0, jitter_m: float = 0.0, seed: int = 0):
    """Apply x' = R x + t (metres) to a normalised (3,n) cloud; returns (cloud', R, t)."""
    a = math.radians(yaw_deg)
    R = torch.tensor([[math.cos(a), -math.sin(a), 0.0], [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]],
                     dtype=torch.float64)
    t = torch.tensor(t_m, dtype=torch.float64).view(3, 1)
    x = cloud.to(torch.float64) * scale
    y = R @ x + t
    if jitter_m > 0:
        g = torch.Generator().manual_seed(int(seed))
        y = y + jitter_m * torch.randn(y.shape, generator=g, dtype=torch.float64)
    return (y / scale).to(torch.float32).contiguous(), R.to(torch.float32), t.to(torch.float32)


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
