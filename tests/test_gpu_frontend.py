"""GPU parity of the preprocessing front-end (dpm_frontend_f32) vs oracle/frontend_ref.py: the SAME points in
the SAME order, bit for bit (integer / index work; the only floating-point outputs are exact divisions)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from deeppointmap_b200 import data, ops
from oracle import frontend_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _raw_frame(seed: int, n: int) -> np.ndarray:
    """KITTI-like raw rows in metres: structured cloud out to 75 m (so DistanceSample has work), near-duplicates
    inside voxels (so 'first' has work), intensity column, a few NaN rows"""
    g = torch.Generator().manual_seed(seed)
    base = data.kitti_shape_cloud(seed, max(1, n // 2)) * 75.0
    dup = base[:, torch.randint(0, base.shape[1], (n - n // 2,), generator=g)] + 0.05 * torch.randn(3, n - n // 2, generator=g)
    xyz = torch.cat([base, dup], dim=1)[:, torch.randperm(n, generator=g)]
    raw = torch.cat([xyz, torch.rand(1, n, generator=g)], dim=0).T.contiguous().numpy().astype(np.float32)
    if n >= 74:
        raw[:: n // 37, 1] = np.nan
    return raw


@pytest.mark.parametrize("n", [122000, 65536, 5000, 64, 1])
def test_frontend_bit_exact(n):
    raw = _raw_frame(n, n)
    want = frontend_ref.preprocess_bin(raw)
    got = ops.preprocess_frame(torch.from_numpy(raw).to(DEV)).cpu()
    assert got.shape == want.shape
    assert torch.equal(got, want)


def test_frontend_golden_other_parameters_and_limits():
    g = np.load(os.path.join(GOLDEN, "frontend.npz"))
    got = ops.preprocess_frame(torch.from_numpy(g["raw"]).to(DEV)).cpu()
    assert torch.equal(got, torch.from_numpy(g["out"]))
    raw = _raw_frame(3, 30000)
    for voxel, lo, hi, ratio in ((0.5, 2.0, 40.0, 40.0), (0.1, 0.0, 1000.0, 1.0)):
        want = frontend_ref.preprocess_bin(raw, voxel, lo, hi, ratio)
        got = ops.preprocess_frame(torch.from_numpy(raw).to(DEV), voxel, lo, hi, ratio, max_voxels=1 << 28).cpu()
        assert torch.equal(got, want)
    xyz3 = torch.from_numpy(np.ascontiguousarray(raw[:, :3])).to(DEV)                 # (N,3) rows work too
    assert torch.equal(ops.preprocess_frame(xyz3).cpu(), frontend_ref.preprocess_bin(raw))
    allnan = torch.full((100, 4), float("nan"), device=DEV)
    assert ops.preprocess_frame(allnan).shape == (3, 0)
    with pytest.raises(ValueError):                                                     # one far outlier: grid too large
        far = torch.from_numpy(raw).to(DEV).clone()
        far[0, 0] = 1.0e6
        ops.preprocess_frame(far, max_voxels=1 << 20)


def test_frontend_feeds_the_encoder(cfg):
    """raw rows -> preprocess_frame -> Encoder: the front-end's output is a valid encoder input"""
    from deeppointmap_b200 import Encoder
    enc = Encoder(cfg).eval().to(DEV)
    pts = ops.preprocess_frame(torch.from_numpy(_raw_frame(9, 100000)).to(DEV))
    assert pts.shape[0] == 3 and pts.shape[1] > 4096 and float(pts.norm(dim=0).max()) <= 1.0 + 1e-6
    coor, fea, pad = enc(pts[None], torch.zeros(1, pts.shape[1], dtype=torch.bool, device=DEV))
    assert coor.shape == (1, 3, 256) and fea.shape == (1, 128, 256) and not bool(pad.any())


@pytest.mark.parametrize("n", [46000, 6000, 300])
def test_outlier_filter_vs_oracle(n):
    """The kept set equals the oracle's except, at most, for points whose statistic lies within 1e-5 (relative) of
    the threshold: the reference sums the global mean / std in fp32 in a build-dependent order, we in fp64."""
    from oracle import outlier_ref
    raw = _raw_frame(n + 1, n)
    xyz = torch.from_numpy(raw[~np.isnan(raw).any(1), :3]).contiguous()
    g = torch.Generator().manual_seed(n)
    xyz[torch.randint(0, xyz.shape[0], (20,), generator=g)] += 30.0 * torch.randn(20, 3, generator=g)   # real outliers
    kept_w, mask_w, stat, thr = outlier_ref.outlier_filter(xyz, 10, 3.0)
    kept, mask = ops.outlier_filter(xyz.to(DEV), 10, 3.0, return_mask=True)
    mask = mask.cpu()
    diff = mask != mask_w
    assert int((~mask_w).sum()) >= 3                                   # the filter has work to do
    assert bool(((stat[diff] - thr).abs() <= 1e-5 * abs(thr)).all()), (int(diff.sum()), thr)
    assert int(diff.sum()) <= 2
    assert torch.equal(kept.cpu(), xyz[mask])                          # survivors, original order
    if not bool(diff.any()):
        assert torch.equal(kept.cpu(), kept_w)


def test_frontend_with_outlier_filter_is_the_yaml_chain():
    """VoxelSample -> DistanceSample -> OutlierFilter(10, 3.0) -> CoordinatesNormalization, the shipped YAML's chain
    minus the open3d LowPassFilter (configs/infer/DeepPointMap_B_Main_SemanticKITTI.yaml:21-29)."""
    from oracle import outlier_ref
    raw = _raw_frame(77, 100000)
    metres = (frontend_ref.preprocess_bin(raw, ratio=1.0)).T.contiguous()
    kept_w, mask_w, stat, thr = outlier_ref.outlier_filter(metres, 10, 3.0)
    want = (kept_w / 60.0).T.contiguous()
    got = ops.preprocess_frame(torch.from_numpy(raw).to(DEV), outlier=(10, 3.0)).cpu()
    assert abs(got.shape[1] - want.shape[1]) <= 2
    if got.shape == want.shape:
        assert torch.equal(got, want)


@pytest.mark.parametrize("n", [30000, 1500])
def test_low_pass_filter_vs_oracle(n):
    """LowPassFilter on the device vs oracle/lowpass_ref.py (pinned on the reference class).  Normals: fp32 moments +
    fp64 Jacobi here, fp64 kd-tree + LAPACK there -- they agree to ~1e-6 except where the two smallest eigenvalues of a
    neighbourhood nearly coincide (the normal is then ill-defined in any implementation), so the bar is: the similarity
    statistic within 1e-3 for >= 99.5 % of the points, and the kept sets equal up to points whose statistic differs or
    sits within 1e-4 of the threshold."""
    from oracle import lowpass_ref
    raw = _raw_frame(n + 3, 4 * n)
    metres = (frontend_ref.preprocess_bin(raw, ratio=1.0)).T.contiguous()[:n]
    kept_w, mask_w, sim_w, thr = lowpass_ref.low_pass_filter(metres, 0.5, 16, 2.0, 4)
    kept, mask = ops.low_pass_filter(metres.to(DEV), 0.5, 16, 2.0, 4, return_mask=True)
    mask = mask.cpu()
    assert 0 < int((~mask_w).sum()) < n // 2                           # the filter has work to do
    # the statistic itself, through the C ABI's sim_out
    from deeppointmap_b200 import _C
    lib = _C.lib()
    r = metres.to(DEV)
    sim = torch.empty(r.shape[0], device=DEV)
    out = torch.empty(r.shape[0], 3, device=DEV)
    cnt = torch.empty(1, dtype=torch.int32, device=DEV)
    nb = lib.dpm_low_pass_filter_workspace_bytes(r.shape[0], 16)
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    _C.check(lib.dpm_low_pass_filter_f32(r.data_ptr(), r.shape[0], 3, 0.5, 16, 2.0, 4, 1.0, out.data_ptr(), None, sim.data_ptr(),
                                         cnt.data_ptr(), ws.data_ptr(), nb, _C.stream_ptr()))
    ds = (sim.cpu() - sim_w).abs()
    assert float((ds < 1e-3).float().mean()) >= 0.995, float((ds < 1e-3).float().mean())
    diff = mask != mask_w
    explained = (ds >= 1e-4) | ((sim_w - thr).abs() <= 1e-4)
    assert bool(explained[diff].all()), int((diff & ~explained).sum())
    assert int(diff.sum()) <= max(3, n // 200)
    assert int(cnt.item()) == int(mask.sum()) and torch.equal(kept.cpu(), metres[mask])   # survivors, original order


def test_full_yaml_chain_on_device():
    """VoxelSample -> DistanceSample -> OutlierFilter -> LowPassFilter -> CoordinatesNormalization: every
    data-dependent step of configs/infer/DeepPointMap_B_Main_SemanticKITTI.yaml:21-29 on the device."""
    from oracle import lowpass_ref, outlier_ref
    raw = _raw_frame(78, 60000)
    metres = (frontend_ref.preprocess_bin(raw, ratio=1.0)).T.contiguous()
    k1, _, _, _ = outlier_ref.outlier_filter(metres, 10, 3.0)
    k2, _, _, _ = lowpass_ref.low_pass_filter(k1, 0.5, 16, 2.0, 4)
    want = (k2 / 60.0).T.contiguous()
    got = ops.preprocess_frame(torch.from_numpy(raw).to(DEV), outlier=(10, 3.0), lowpass=(0.5, 16, 2.0, 4)).cpu()
    assert got.shape[0] == 3 and abs(got.shape[1] - want.shape[1]) <= max(4, want.shape[1] // 200)
    # max_remain: the reference re-ranks by similarity and keeps that order
    sub = ops.low_pass_filter(k1.to(DEV), 0.5, 16, 2.0, 4, max_remain=1000)
    assert sub.shape == (1000, 3)
