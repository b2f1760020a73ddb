"""GPU parity: FPS / kNN / kNN+radius / ball query through the C-ABI vs the C oracle.
Bar: BIT-EXACT indices (and distances)."""
import os

import pytest
import torch

from oracle import index_ops as IO
from deeppointmap_b200 import data, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(params=[1, 2, 3], ids=["one-sm", "cluster", "packed"])
def fps_mode(request):
    """every FPS test runs on all three mappings: one CTA per cloud (grid.cu / fps.cu), an 8-CTA cluster per cloud
    (fps_cluster.cu) and two clouds per CTA (two teams with their own named barriers, grid.cu)"""
    ops.set_fps_mode(request.param)
    yield request.param
    ops.set_fps_mode(0)


def _cloud(kind, seed, n):
    c = data.kitti_shape_cloud(seed, n) if kind == "kitti" else data.uniform_cube_cloud(seed, n)
    return c.T.contiguous()  # (n, 3)


@pytest.mark.parametrize("n,k", [(16, 16), (64, 16), (256, 64), (1000, 100), (1024, 256), (4096, 1024),
                                 (14500, 4096), (20000, 512), (65536, 4096), (100000, 1024), (131072, 64)])
def test_fps_bit_exact(n, k, fps_mode):
    pts = torch.stack([_cloud("kitti", n, n), _cloud("cube", n + 1, n)])
    want = IO.fps(pts, None, k)
    _, got = ops.sample_farthest_points(pts.to(DEV), K=k)
    assert torch.equal(got.cpu(), want)


def test_fps_lengths_padding_and_gather(fps_mode):
    n, k = 3000, 700
    pts = torch.stack([_cloud("kitti", 1, n), _cloud("cube", 2, n), _cloud("kitti", 3, n)])
    pts = torch.cat([pts, torch.arange(n, dtype=torch.float32).view(1, n, 1).expand(3, n, 1)], dim=2)  # D = 4
    lengths = torch.tensor([3000, 500, 1])
    want = IO.fps(pts, lengths, k)
    out, got = ops.sample_farthest_points(pts.to(DEV), lengths.to(DEV), K=k)
    assert torch.equal(got.cpu(), want)
    assert (got[1, 500:] == -1).all() and (got[2, 1:] == -1).all()
    ref = torch.gather(pts, 1, want.clamp(min=0)[..., None].expand(-1, -1, 4)).clone()
    ref[want < 0] = 0
    assert torch.equal(out.cpu(), ref)  # masked_gather: -1 rows are zero


def test_fps_duplicates_first_maximum(fps_mode):
    pts = torch.zeros(1, 600, 3)
    pts[0, 300:] = 1.0
    want = IO.fps(pts, None, 5)
    _, got = ops.sample_farthest_points(pts.to(DEV), K=5)
    assert got.cpu().tolist() == want.tolist() == [[0, 300, 0, 0, 0]]


def test_fps_batch_of_full_frames(fps_mode):
    pts = torch.stack([_cloud("kitti", s, 65536) for s in range(3)])  # odd batch: the packed mapping's last team idles
    want = IO.fps(pts, None, 4096)
    _, got = ops.sample_farthest_points(pts.to(DEV), K=4096)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("n,k", [(300000, 300), (600000, 200)])
def test_fps_beyond_the_cluster_limit(n, k, fps_mode):
    """262 145 .. 1 048 576 points (pytorch3d has no limit; VERDICT r1 missing #7): the one-CTA grid kernel with 16 / 32
    points per lane, whatever mapping is asked for; ragged lengths included."""
    pts = torch.stack([_cloud("kitti", 7, n), _cloud("cube", 8, n)])
    lengths = torch.tensor([n, n - 12345])
    want = IO.fps(pts, lengths, k)
    _, got = ops.sample_farthest_points(pts.to(DEV), lengths.to(DEV), K=k)
    assert torch.equal(got.cpu(), want)


def test_fps_rejects_oversize():
    with pytest.raises(NotImplementedError):
        ops.sample_farthest_points(torch.zeros(1, (1 << 20) + 1, 3, device=DEV), K=4)


@pytest.mark.parametrize("s,n,k", [(64, 500, 33), (100, 3000, 100), (33, 70000, 64), (16, 40, 40), (8, 300, 257)])
def test_knn_more_than_32_neighbours(s, n, k):
    """K > 32 (pytorch3d has no limit; VERDICT r1 missing #7): multi-pass kernel, indices and distances bit-exact,
    ragged clouds and K > length zero-padded like the K <= 32 paths."""
    p2 = torch.stack([_cloud("kitti", n, n), _cloud("cube", n + 5, n)])
    p1 = torch.stack([p2[0, :s] + 0.001, _cloud("cube", 99, s)])
    l2 = torch.tensor([n, max(1, min(n, k - 3))])
    l1 = torch.tensor([s, s - 2])
    wd, wi = IO.knn(p1, p2, l2, k)
    got = ops.knn_points(p1.to(DEV), p2.to(DEV), lengths1=l1.to(DEV), lengths2=l2.to(DEV), K=k)
    gi, gd = got.idx.cpu(), got.dists.cpu()
    assert torch.equal(gi[0], wi[0]) and torch.equal(gd[0], wd[0])
    assert torch.equal(gi[1, :s - 2], wi[1, :s - 2]) and torch.equal(gd[1, :s - 2], wd[1, :s - 2])
    assert (gi[1, s - 2:] == 0).all() and (gd[1, s - 2:] == 0).all()
    assert (gi[1, :, int(l2[1]):] == 0).all() and (gd[1, :, int(l2[1]):] == 0).all()


@pytest.mark.parametrize("n,k,r", [(3000, 48, 0.08), (400, 40, 0.5)])
def test_hybrid_more_than_32_neighbours(n, k, r):
    p2 = torch.stack([_cloud("kitti", n + 3, n), _cloud("cube", n + 4, n)])
    p1 = torch.cat([p2[:, :200], p2[:, :8] + 50.0], dim=1).contiguous()  # the last 8 queries see nothing in the radius
    pad = torch.zeros(2, n, dtype=torch.bool)
    pad[1, n - 100:] = True
    want = IO.hybrid(p1, p2, (~pad).sum(1), k, r)
    got = ops.hybrid_query(r, k, p2.to(DEV), p1.to(DEV), pad.to(DEV))
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("s,n,k", [(16, 16, 16), (64, 256, 32), (256, 1024, 32), (1024, 4096, 32), (300, 5000, 11),
                                   (4096, 14500, 32), (4096, 65536, 32), (777, 2049, 1), (50, 2048, 17)])
def test_knn_bit_exact(s, n, k):
    p2 = torch.stack([_cloud("kitti", n, n), _cloud("cube", n + 5, n)])
    p1 = torch.stack([p2[0, :s] + 0.001, _cloud("cube", 99, s)])
    wd, wi = IO.knn(p1, p2, None, k)
    got = ops.knn_points(p1.to(DEV), p2.to(DEV), K=k)
    assert torch.equal(got.idx.cpu(), wi)
    assert torch.equal(got.dists.cpu(), wd)  # distances bit-exact too


def test_knn_lengths_and_zero_padding():
    p2 = torch.stack([_cloud("kitti", 1, 500), _cloud("cube", 2, 500)])
    p1 = p2[:, :40].clone()
    l2 = torch.tensor([500, 7])
    wd, wi = IO.knn(p1, p2, l2, 16)
    got = ops.knn_points(p1.to(DEV), p2.to(DEV), lengths2=l2.to(DEV), K=16)
    assert torch.equal(got.idx.cpu(), wi) and torch.equal(got.dists.cpu(), wd)
    assert (got.idx[1, :, 7:] == 0).all() and (got.dists[1, :, 7:] == 0).all()


@pytest.mark.parametrize("n,k", [(46000, 12), (46000, 18), (2048, 32), (30000, 1)])
def test_knn_grid_shells_self_and_foreign_queries(n, k):
    """plain kNN on a large cloud takes the cell-grid + shell-expansion path: self queries (OutlierFilter /
    LowPassFilter, dataloader/transforms.py:236-289), queries far outside the cloud, ragged lengths, a cloud with
    fewer than K valid points, lattice ties -- all bit-exact (indices and distances) against the C oracle."""
    cloud = _cloud("kitti", n, n) * 60.0  # metres, like the filters see it
    far = torch.tensor([[500.0, -300.0, 40.0], [-1e4, 0.0, 0.0], [0.0, 0.0, 1e3], [59.9, 59.9, 5.0]])
    lattice = (torch.stack(torch.meshgrid(torch.arange(16.0), torch.arange(16.0), torch.arange(8.0), indexing="ij"), -1)
               .reshape(-1, 3) * 0.5)
    p2 = torch.stack([cloud, torch.cat([lattice, cloud[: n - lattice.shape[0]] + 100.0])])
    q = torch.stack([torch.cat([cloud[:1500], far, cloud[-500:] + 0.3]), torch.cat([lattice[:1500] + 0.25, far, lattice[:500]])])
    l2 = torch.tensor([n, min(n, lattice.shape[0] + 7)])
    l1 = torch.tensor([q.shape[1], 1700])
    wd, wi = IO.knn(q, p2, l2, k)
    got = ops.knn_points(q.to(DEV), p2.to(DEV), lengths1=l1.to(DEV), lengths2=l2.to(DEV), K=k)
    gi, gd = got.idx.cpu(), got.dists.cpu()
    assert torch.equal(gi[0], wi[0]) and torch.equal(gd[0], wd[0])
    assert torch.equal(gi[1, :1700], wi[1, :1700]) and torch.equal(gd[1, :1700], wd[1, :1700])
    assert (gi[1, 1700:] == 0).all() and (gd[1, 1700:] == 0).all()          # queries past lengths1: zeros
    tiny = ops.knn_points(q[:1, :64].to(DEV), p2[:1].to(DEV), lengths2=torch.tensor([5], device=DEV), K=k)
    wd5, wi5 = IO.knn(q[:1, :64], p2[:1], torch.tensor([5]), k)
    assert torch.equal(tiny.idx.cpu(), wi5) and torch.equal(tiny.dists.cpu(), wd5)  # fewer than K points: zero padded


def test_knn_ties_lower_index_first():
    p2 = torch.zeros(1, 100, 3)
    p2[0, 50:] = 2.0
    got = ops.knn_points(torch.zeros(1, 3, 3, device=DEV), p2.to(DEV), K=32)
    assert got.idx[0, 0].cpu().tolist() == list(range(32))


@pytest.mark.parametrize("s,n,k,r", [(4096, 65536, 32, 0.05), (4096, 4096, 32, 0.1), (1024, 4096, 32, 0.1),
                                     (256, 256, 32, 0.4), (16, 16, 16, 1.6), (64, 64, 32, 0.8), (500, 3000, 16, 0.02)])
def test_hybrid_bit_exact(s, n, k, r):
    p2 = torch.stack([_cloud("kitti", n + 3, n), _cloud("cube", n + 4, n)])
    p1 = p2[:, torch.randperm(n, generator=torch.Generator().manual_seed(0))[:s]].contiguous()
    pad = torch.zeros(2, n, dtype=torch.bool)
    want = IO.hybrid(p1, p2, (~pad).sum(1), k, r)
    got = ops.hybrid_query(r, k, p2.to(DEV), p1.to(DEV), pad.to(DEV))
    assert torch.equal(got.cpu(), want)


def test_hybrid_no_point_in_radius_and_padding():
    p2 = _cloud("cube", 1, 400)[None]
    p1 = torch.tensor([[[5.0, 5.0, 5.0], [0.0, 0.0, 0.0]]])  # first centre far from everything
    pad = torch.zeros(1, 400, dtype=torch.bool)
    pad[:, 300:] = True
    want = IO.hybrid(p1, p2, (~pad).sum(1), 8, 0.05)
    got = ops.hybrid_query(0.05, 8, p2.to(DEV), p1.to(DEV), pad.to(DEV))
    assert torch.equal(got.cpu(), want)
    assert len(set(got[0, 0].cpu().tolist())) == 1  # every slot = the nearest point
    assert got.max() < 300


@pytest.mark.parametrize("k,r", [(8, 0.1), (32, 0.3), (64, 0.2)])
def test_ball_query_bit_exact(k, r):
    p2 = torch.stack([_cloud("kitti", 1, 3000), _cloud("cube", 2, 3000)])
    p1 = p2[:, :200].clone()
    wd, wi = IO.ball_query(p1, p2, torch.tensor([3000, 1000]), k, r)
    got = ops.ball_query(p1.to(DEV), p2.to(DEV), lengths2=torch.tensor([3000, 1000], device=DEV), K=k, radius=r,
                         return_nn=True)
    assert torch.equal(got.idx.cpu(), wi) and torch.equal(got.dists.cpu(), wd)
    assert got.knn.shape == (2, 200, k, 3)


def test_sortedness_and_idempotence_at_full_size():
    """Size-independent properties at the BASELINE size (65 536 points)."""
    p2 = _cloud("kitti", 0, 65536)[None].to(DEV)
    _, idx = ops.sample_farthest_points(p2, K=4096)
    assert idx.unique().numel() == 4096 and idx[0, 0] == 0  # FPS never repeats a point on distinct inputs
    ctr = torch.gather(p2, 1, idx[..., None].expand(-1, -1, 3))
    res = ops.knn_points(ctr, p2, K=32)
    assert (res.dists[..., 1:] >= res.dists[..., :-1]).all()  # ascending
    assert torch.equal(res.idx[..., 0], idx)  # each centre's nearest point is itself (d2 = 0)
    assert (res.dists[..., 0] == 0).all()
    again = ops.knn_points(ctr, p2, K=32)
    assert torch.equal(again.idx, res.idx)  # deterministic


# ---- the cell-grid kernels (N >= 2048): ties, padding, degenerate clouds, determinism ----------
def _lattice(n, seed):
    """integer-lattice cloud: masses of exactly equal distances (tie-break stress)"""
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(0, 12, (n, 3), generator=g).float() * 0.05).contiguous()


@pytest.mark.parametrize("n,k", [(2048, 300), (5000, 1000), (40000, 600)])
def test_fps_grid_ties_on_lattice(n, k, fps_mode):
    pts = torch.stack([_lattice(n, 1), _lattice(n, 2)])
    want = IO.fps(pts, None, k)
    _, got = ops.sample_farthest_points(pts.to(DEV), K=k)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("n,k", [(6000, 900), (12000, 700)])  # register-resident kernel / pruned (grid) kernel
def test_fps_grid_lengths_and_degenerate_clouds(n, k, fps_mode):
    pts = torch.stack([_cloud("kitti", 1, n), _cloud("cube", 2, n), torch.zeros(n, 3), _cloud("kitti", 3, n),
                       _cloud("cube", 4, n) * torch.tensor([1.0, 0.0, 0.0])])  # identical points; a line
    lengths = torch.tensor([n, 2500, n, 1, n])
    want = IO.fps(pts, lengths, k)
    _, got = ops.sample_farthest_points(pts.to(DEV), lengths.to(DEV), K=k)
    assert torch.equal(got.cpu(), want)
    assert (got[3, 1:] == -1).all()


@pytest.mark.parametrize("n,s,k,r", [(2048, 500, 32, 0.11), (8000, 2000, 16, 0.05), (30000, 1000, 32, 0.1)])
def test_hybrid_grid_ties_on_lattice(n, s, k, r):
    p2 = torch.stack([_lattice(n, 3), _lattice(n, 4)])
    p1 = torch.stack([p2[0, :s], _lattice(s, 5) + 0.013])
    pad = torch.zeros(2, n, dtype=torch.bool)
    pad[1, n // 2:] = True
    want = IO.hybrid(p1, p2, (~pad).sum(1), k, r)
    got = ops.hybrid_query(r, k, p2.to(DEV), p1.to(DEV), pad.to(DEV))
    assert torch.equal(got.cpu(), want)


def test_hybrid_grid_queries_outside_the_cloud_and_huge_radius():
    n = 4096
    p2 = _cloud("cube", 7, n)[None]
    p1 = torch.tensor([[[5.0, 5.0, 5.0], [1.02, 0.0, 0.0], [-1.04, -1.04, -1.04], [0.0, 0.0, 0.0], [1e6, 0.0, 0.0]]])
    pad = torch.zeros(1, n, dtype=torch.bool)
    for r in (0.05, 0.3, 50.0):
        want = IO.hybrid(p1, p2, (~pad).sum(1), 32, r)
        got = ops.hybrid_query(r, 32, p2.to(DEV), p1.to(DEV), pad.to(DEV))
        assert torch.equal(got.cpu(), want), r


def test_grid_results_are_deterministic(fps_mode):
    """the cell-sorted layout depends on atomic order; the results must not"""
    pts = torch.stack([_cloud("kitti", 5, 65536), _lattice(65536, 6)]).to(DEV)
    ref_f = ops.sample_farthest_points(pts, K=1024)[1]
    ctr = torch.gather(pts, 1, ref_f[..., None].expand(-1, -1, 3))
    pad = torch.zeros(2, 65536, dtype=torch.bool, device=DEV)
    ref_q = ops.hybrid_query(0.05, 32, pts, ctr, pad)
    for _ in range(3):
        assert torch.equal(ops.sample_farthest_points(pts, K=1024)[1], ref_f)
        assert torch.equal(ops.hybrid_query(0.05, 32, pts, ctr, pad), ref_q)


# ---- information matrix: 1-NN within a radius + G^T G reduction ---------------
@pytest.mark.parametrize("n1,n2", [(16384, 16384), (65536, 60000), (1500, 900), (5000, 100), (1, 1)])
def test_information_matrix_vs_oracle(n1, n2):
    """tolerance 1e-4 relative to the largest entry (fp32 sums of ~1e4 products in the reference)"""
    import math
    from oracle import infomat_ref
    from deeppointmap_b200 import ops
    c0 = data.kitti_shape_cloud(21, n1) * 60.0
    c1, _, _ = data.rigid_move(c0 / 60.0, yaw_deg=2.0, t_m=(1.0, 0.1, 0.0), jitter_m=0.02, seed=22)
    c1 = (c1 * 60.0)[:, torch.randperm(n1, generator=torch.Generator().manual_seed(3))[:n2] % n1].contiguous()
    a = math.radians(2.0)
    T = torch.eye(4)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = math.cos(a), -math.sin(a), math.sin(a), math.cos(a)
    T[:3, 3] = torch.tensor([1.0, 0.1, 0.0])
    want, n = infomat_ref.information_matrix(c0, c1, T, 1.0)
    got, cnt = ops.information_matrix(c0.to(DEV), c1.to(DEV), T, 1.0, return_count=True)
    assert abs(int(cnt) - n) <= max(2, n // 2000)  # a transformed coordinate may differ by 1 ulp from torch's matmul
    assert float((got.cpu() - want).abs().max()) <= 1e-4 * max(float(want.abs().max()), 1.0)


def test_information_matrix_golden_and_empty():
    import numpy as np
    from conftest import GOLDEN
    from deeppointmap_b200 import ops
    g = np.load(os.path.join(GOLDEN, "infomat.npz"))
    for name in ("kitti6k", "uniform", "far_apart"):
        got = ops.information_matrix(torch.from_numpy(g[name + "_p1"]).to(DEV), torch.from_numpy(g[name + "_p2"]).to(DEV),
                                     torch.from_numpy(g[name + "_T"]), float(g[name + "_radius"])).cpu()
        want = torch.from_numpy(g[name + "_info"])
        assert float((got - want).abs().max()) <= 1e-4 * max(float(want.abs().max()), 1e-30), name
    assert float(ops.information_matrix(torch.zeros(3, 0, device=DEV), torch.zeros(3, 5, device=DEV), torch.eye(4)).abs().max()) == 0.0
