"""Deterministic synthetic inputs for the hot path.  The preprocessing of a raw
KITTI .bin frame is `deeppointmap_b200.ops.preprocess_frame` (CUDA); its CPU restatement lives with the other
checkers in oracle/frontend_ref.py."""
import math

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
