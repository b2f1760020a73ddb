"""Frame-to-frame odometry over a synthetic sequence (BASELINE.json config 5).

The reference's `pipeline/infer.py` cannot start on this image and its tree does not
exist on the GPU box, so this is the part of its odometry loop that IS the hot path -- what
`ExtractionThread.process` + `OdometryThread` do per scan (system/modules/odometry.py:36-54, 96-121):
encode scans in batches of `EXTRACTOR_BATCHSIZE`, register every scan against its predecessor, chain the
relative poses -- with the descriptors kept on the device instead of `.cpu()` / `.to(device)` round trips.

`corridor_world` / `corridor_frames` build a deterministic KITTI-shaped world (ground strip + facades along a
road) and what a 60 m sensor moving along a smooth trajectory sees of it, so the estimated trajectory can be
checked against ground truth.
"""
import math
from typing import List, Tuple

import torch


def corridor_world(seed: int, length_m: float, half_width_m: float = 70.0, ground_density: float = 7.0,
                   facades_per_100m: int = 24, device="cpu") -> torch.Tensor:
    """(3, P) fp32 world points in metres: a ground strip along +x and vertical facades on both sides."""
    g = torch.Generator().manual_seed(int(seed))
    L = float(length_m) + 140.0
    ng = int(L * 2 * half_width_m * ground_density)
    ground = torch.stack([torch.rand(ng, generator=g) * L - 70.0, (torch.rand(ng, generator=g) * 2 - 1) * half_width_m,
                          -1.73 + 0.02 * torch.randn(ng, generator=g)])
    nf = max(1, int(L / 100.0 * facades_per_100m))
    cx = torch.rand(nf, generator=g) * L - 70.0
    cy = (torch.rand(nf, generator=g) * 2 - 1) * (half_width_m - 8.0)
    cy = torch.where(cy.abs() < 6.0, cy.sign() * 6.0 + cy, cy)  # keep the road itself free
    heading = torch.rand(nf, generator=g) * math.pi
    length = 5.0 + 20.0 * torch.rand(nf, generator=g)
    per = 2500
    pid = torch.arange(nf).repeat_interleave(per)
    along = (torch.rand(nf * per, generator=g) - 0.5) * length[pid]
    facade = torch.stack([cx[pid] + along * torch.cos(heading[pid]), cy[pid] + along * torch.sin(heading[pid]),
                          -1.73 + 6.0 * torch.rand(nf * per, generator=g)])
    return torch.cat([ground, facade], dim=1).to(device=device, dtype=torch.float32).contiguous()


def street_world(seed: int, length_m: float, half_width_m: float = 60.0, device="cpu") -> torch.Tensor:
    """(3, P) fp32 world points in metres with the kind of structure a street scan has, for runs where the trained
    network has to register consecutive scans well (the reference's whole pipeline): a gently undulating ground sampled
    on a jittered lattice, kerbs along the road, building facades with window recesses, poles, parked boxes and tree
    crowns.  Deterministic in `seed`."""
    g = torch.Generator().manual_seed(int(seed))
    L = float(length_m) + 140.0
    u = lambda *shape: torch.rand(*shape, generator=g)
    parts = []
    # ground: 0.35 m lattice + jitter, low-frequency height field
    gx = torch.arange(-70.0, L - 70.0, 0.35)
    gy = torch.arange(-half_width_m, half_width_m, 0.35)
    X, Y = torch.meshgrid(gx, gy, indexing="ij")
    X = X.flatten() + 0.1 * (u(X.numel()) - 0.5)
    Y = Y.flatten() + 0.1 * (u(Y.numel()) - 0.5)
    Z = -1.73 + 0.15 * torch.sin(X / 9.0) * torch.cos(Y / 7.0) + 0.05 * torch.sin(X / 2.3 + Y / 3.1)
    parts.append(torch.stack([X, Y, Z]))
    # kerbs: two 12 cm steps along the road
    for side in (-1.0, 1.0):
        kx = torch.arange(-70.0, L - 70.0, 0.05)
        for dz in (0.0, 0.06, 0.12):
            parts.append(torch.stack([kx, torch.full_like(kx, side * 4.0), torch.full_like(kx, -1.73 + dz)]))
    # facades: rectangles 6-25 m long, 4-12 m high, with window recesses every 3 m
    nb = max(4, int(L / 100.0 * 14))
    for i in range(nb):
        cx = float(u(1)) * L - 70.0
        side = 1.0 if i % 2 == 0 else -1.0
        cy = side * (9.0 + 18.0 * float(u(1)))
        ln, ht = 6.0 + 19.0 * float(u(1)), 4.0 + 8.0 * float(u(1))
        yaw = (float(u(1)) - 0.5) * 0.5
        a = torch.arange(-ln / 2, ln / 2, 0.12)
        h = torch.arange(0.0, ht, 0.12)
        A, H = torch.meshgrid(a, h, indexing="ij")
        A, H = A.flatten(), H.flatten()
        recess = (((A + ln) % 3.0) < 1.2) & (((H % 3.0) > 1.0) & ((H % 3.0) < 2.4))
        depth = torch.where(recess, torch.full_like(A, 0.25 * side), torch.zeros_like(A))
        parts.append(torch.stack([cx + A * math.cos(yaw) - depth * math.sin(yaw), cy + A * math.sin(yaw) + depth * math.cos(yaw), -1.73 + H]))
    # poles and tree trunks + crowns
    npole = max(8, int(L / 100.0 * 30))
    for i in range(npole):
        px, py = float(u(1)) * L - 70.0, (1.0 if i % 2 else -1.0) * (4.5 + 10.0 * float(u(1)))
        hgt = 3.0 + 5.0 * float(u(1))
        z = torch.arange(0.0, hgt, 0.04)
        th = u(z.numel()) * 2 * math.pi
        parts.append(torch.stack([px + 0.12 * torch.cos(th), py + 0.12 * torch.sin(th), -1.73 + z]))
        if i % 3 == 0:
            n = 1500
            d = torch.randn(3, n, generator=g)
            d = d / d.norm(dim=0, keepdim=True) * (1.2 + 0.8 * u(n))
            parts.append(torch.stack([px + d[0], py + d[1], -1.73 + hgt + d[2].abs()]))
    # parked boxes (cars): 4.2 x 1.8 x 1.5 m
    ncar = max(6, int(L / 100.0 * 16))
    for i in range(ncar):
        bx, by = float(u(1)) * L - 70.0, (1.0 if i % 2 else -1.0) * (2.9 + 0.4 * float(u(1)))
        n = 2500
        f = torch.randint(0, 5, (n,), generator=g)
        a_, b_, c_ = u(n) - 0.5, u(n) - 0.5, u(n)
        x = torch.where(f == 0, torch.full_like(a_, -0.5), torch.where(f == 1, torch.full_like(a_, 0.5), a_)) * 4.2
        y = torch.where(f == 2, torch.full_like(b_, -0.5), torch.where(f == 3, torch.full_like(b_, 0.5), b_)) * 1.8
        z = torch.where(f == 4, torch.ones_like(c_), c_) * 1.5
        parts.append(torch.stack([bx + x, by + y, -1.73 + z]))
    return torch.cat(parts, dim=1).to(device=device, dtype=torch.float32).contiguous()


def trajectory(n_frames: int, step_m: float = 1.0, max_yaw_deg: float = 2.0) -> torch.Tensor:
    """(n,4,4) fp64 sensor poses in the world: `step_m` per frame along a gently weaving heading."""
    poses, x, y, yaw = [], 0.0, 0.0, 0.0
    for i in range(n_frames):
        T = torch.eye(4, dtype=torch.float64)
        c, s = math.cos(yaw), math.sin(yaw)
        T[0, 0], T[0, 1], T[1, 0], T[1, 1] = c, -s, s, c
        T[0, 3], T[1, 3] = x, y
        poses.append(T)
        yaw_rate = math.radians(max_yaw_deg) * math.sin(2 * math.pi * i / 97.0) * math.cos(2 * math.pi * i / 41.0)
        yaw = max(-0.35, min(0.35, yaw + yaw_rate))
        x += step_m * math.cos(yaw)
        y += step_m * math.sin(yaw)
    return torch.stack(poses)


@torch.no_grad()
def corridor_frames(world: torch.Tensor, poses: torch.Tensor, n_points: int, seed: int = 0, max_range_m: float = 60.0,
                    jitter_m: float = 0.01, scale: float = 60.0, stable: bool = False) -> torch.Tensor:
    """What the sensor sees: (F, 3, n_points) fp32 normalised clouds in the sensor frames (1 m <= |p| <= range,
    a fresh random subset and jitter per frame) on the world tensor's device.  stable=True: every world point has a
    fixed priority and a frame keeps the n_points in range with the highest one, so consecutive frames sample the
    SAME surface points (a static world re-observed, as in bench.py's pairs) instead of fresh ones."""
    dev = world.device
    g = torch.Generator(device=dev).manual_seed(int(seed))
    out = torch.empty((poses.shape[0], 3, n_points), dtype=torch.float32, device=dev)
    prio = None
    if stable:
        gp = torch.Generator(device=dev).manual_seed(977)
        prio = torch.rand(world.shape[1], generator=gp, device=dev)
    for i in range(poses.shape[0]):
        Tinv = torch.linalg.inv(poses[i]).to(device=dev, dtype=torch.float32)
        near = ((world[0] - float(poses[i, 0, 3])).abs() <= max_range_m)
        p = Tinv[:3, :3] @ world[:, near] + Tinv[:3, 3:]
        d = p.norm(dim=0)
        inr = (d >= 1.0) & (d <= max_range_m)
        p = p[:, inr]
        if p.shape[1] < n_points:
            raise ValueError(f"frame {i}: only {p.shape[1]} world points in range, need {n_points} (raise ground_density)")
        if stable:
            sel = torch.topk(prio[near][inr], n_points).indices
            sel = sel[torch.randperm(n_points, generator=g, device=dev)]
        else:
            sel = torch.randperm(p.shape[1], generator=g, device=dev)[:n_points]
        out[i] = (p[:, sel] + jitter_m * torch.randn((3, n_points), generator=g, device=dev)) / scale
    return out


@torch.no_grad()
def run_odometry(encoder, decoder, frames: torch.Tensor, batch: int = 32, coor_scale: float = 60.0,
                 num_sample: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
    """frames (F,3,N) on the device -> (relative (F-1,16) pose records [R(9) T(3) rmse ...] of scan i-1 -> scan i,
    absolute (F,4,4) fp64 poses chained from identity).  One encoder call and one registration call per batch,
    descriptors stay on the device, one host read at the end."""
    F = frames.shape[0]
    S, Cd = encoder._out_points, encoder.final_channel + 3
    desc = torch.zeros((batch + 1, Cd, S), dtype=torch.float32, device=frames.device)
    records: List[torch.Tensor] = []
    have_prev = False
    for b0 in range(0, F, batch):
        nb = min(batch, F - b0)
        encoder.descriptors(frames[b0:b0 + nb], None, coor_scale=coor_scale, out=desc[1:nb + 1])
        lo = 0 if have_prev else 1  # the first scan of the sequence has no predecessor
        if nb + 1 - lo >= 2:
            res, _ = decoder.registration_forward_batch(desc[lo:nb], desc[lo + 1:nb + 1], num_sample)
            records.append(res)
        desc[0].copy_(desc[nb])
        have_prev = True
    rel = torch.cat(records) if records else torch.zeros((0, 16), device=frames.device)
    host = rel.double().cpu()
    poses = [torch.eye(4, dtype=torch.float64)]
    for i in range(host.shape[0]):
        T = torch.eye(4, dtype=torch.float64)
        T[:3, :3] = host[i, 0:9].view(3, 3)
        T[:3, 3] = host[i, 9:12]
        poses.append(poses[-1] @ torch.linalg.inv(T))  # p_i = T p_{i-1}  =>  pose_i = pose_{i-1} T^-1
    return rel, torch.stack(poses)


def relative_errors(rel: torch.Tensor, gt_poses: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """per-pair translation error (m) and rotation error (deg) of the estimated scan i-1 -> i transforms"""
    host = rel.double().cpu()
    te, re = [], []
    for i in range(host.shape[0]):
        gt = torch.linalg.inv(gt_poses[i + 1]) @ gt_poses[i]
        R, t = host[i, 0:9].view(3, 3), host[i, 9:12]
        te.append(float((t - gt[:3, 3]).norm()))
        c = float(((R.T @ gt[:3, :3]).trace() - 1) / 2)
        re.append(math.degrees(math.acos(max(-1.0, min(1.0, c)))))
    return torch.tensor(te), torch.tensor(re)
