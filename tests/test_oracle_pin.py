"""Pin the oracle: (1) against the reference itself, imported from /root/reference (build
container only), (2) against the committed golden fixtures those runs produced, (3) the C index
ops against the reference's own runnable fallbacks."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REF, ROOT, rel_err
from oracle import index_ops as IO
from oracle import model_ref as M
from deeppointmap_b200 import data

HAS_REF = os.path.isdir(os.path.join(REF, "network"))
needs_ref = pytest.mark.skipif(not HAS_REF, reason="/root/reference not present on this box")


@pytest.fixture(scope="module")
def reference():
    from oracle import ref_loader
    saved = {k: v for k, v in sys.modules.items() if k == "pytorch3d" or k.startswith("pytorch3d.")}
    # the pin needs the reference's own pure-torch fallbacks (CPU): `import pytorch3d` fails while its modules load
    ref_loader.activate("fallback")
    import yaml
    from easydict import EasyDict
    from network.encoder.encoder import Encoder
    from network.decoder.decoder import Decoder
    from network.encoder import utils as RU
    cfg = EasyDict(yaml.safe_load(open(f"{REF}/configs/infer/DeepPointMap_B_Main_SemanticKITTI.yaml")))
    ck = torch.load(f"{REF}/DeepPointMapAAAI.pth", map_location="cpu")
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    enc.load_state_dict(ck["encoder"], strict=True)
    dec.load_state_dict(ck["decoder"], strict=False)
    del sys.modules["pytorch3d"]
    sys.modules.update(saved)
    return dict(enc=enc, dec=dec, RU=RU, ck=ck)


# ---- arithmetic the index parity rests on ---------------------------------------------------
def test_d2_is_unfused_fp32():
    rng = np.random.RandomState(0)
    for _ in range(200):
        a, b = rng.randn(3).astype(np.float32), rng.randn(3).astype(np.float32)
        d = (a - b).astype(np.float32)
        sq = (d * d).astype(np.float32)
        want = np.float32(np.float32(sq[0] + sq[1]) + sq[2])
        assert IO.d2(a, b) == float(want)


def test_radius_compare_is_fp32():
    """`dists > radius ** 2` (utils.py:119): torch casts the Python double to the tensor dtype."""
    for r in (0.05, 0.1, 0.2, 0.4, 0.8, 1.6):
        r2 = np.float32(r ** 2)
        d = torch.tensor([np.nextafter(r2, np.float32(0)), r2, np.nextafter(r2, np.float32(10))])
        assert (d > (r ** 2)).tolist() == [False, False, True]
        assert IO.radius2_f32(r) == float(r2)


# ---- C index ops vs the reference's runnable fallbacks ---------------------------------------
@needs_ref
@pytest.mark.reference
@pytest.mark.parametrize("n,k,valid", [(3000, 257, 3000), (2048, 64, 1500), (100, 128, 100)])
def test_fps_matches_reference_sampler(reference, n, k, valid):
    RU = reference["RU"]
    pts = data.kitti_shape_cloud(7, n).T[None].contiguous()
    pts = torch.cat([pts, data.uniform_cube_cloud(8, n).T[None]], 0)
    pad = torch.zeros(2, n, dtype=torch.bool)
    pad[:, valid:] = True
    ref_pts, ref_mask = RU.Sampler.fps(points=pts, points_padding=pad, K=k)
    idx = IO.fps(pts, (~pad).sum(1), k)
    assert torch.equal(idx < 0, ref_mask)
    got = torch.gather(pts, 1, idx.clamp(min=0)[..., None].expand(-1, -1, 3))
    got[idx < 0] = 0
    assert torch.equal(got, ref_pts)
    # the pure-torch restatement bench.py's reference_gpu leg runs on the device
    assert torch.equal(M.fps_torch(pts, (~pad).sum(1), k), idx)


@needs_ref
@pytest.mark.reference
def test_hybrid_matches_reference_fallback_as_sets(reference):
    RU = reference["RU"]
    pts = data.kitti_shape_cloud(9, 6000).T[None].contiguous()
    pad = torch.zeros(1, 6000, dtype=torch.bool)
    ctr = pts[:, :700]
    ref = RU.Querier.hybrid_query(radius=0.05, K=32, points=pts, centers=ctr, points_padding=pad)
    got = IO.hybrid(ctr, pts, (~pad).sum(1), 32, 0.05)
    same = (ref.sort(-1)[0] == got.sort(-1)[0]).all(-1)
    # the fallback computes |a|^2+|b|^2-2ab through a matmul: near-ties may swap at the boundary
    assert same.float().mean() >= 0.995


@needs_ref
@pytest.mark.reference
def test_hybrid_beyond_32_neighbours_matches_reference_fallback(reference):
    """K > 32 (the multi-pass GPU kernel is judged against this oracle): the C oracle without its old 64-neighbour
    cap against the reference's own pure-torch hybrid / knn queries at K = 48 and K = 100."""
    RU = reference["RU"]
    pts = data.kitti_shape_cloud(12, 3000).T[None].contiguous()
    pad = torch.zeros(1, 3000, dtype=torch.bool)
    ctr = pts[:, :300]
    for K, r in ((48, 0.08), (100, 0.2)):
        ref = RU.Querier.hybrid_query(radius=r, K=K, points=pts, centers=ctr, points_padding=pad)
        got = IO.hybrid(ctr, pts, (~pad).sum(1), K, r)
        same = (ref.sort(-1)[0] == got.sort(-1)[0]).all(-1)
        assert same.float().mean() >= 0.99  # the fallback's matmul-form distances may swap a boundary neighbour
    ref = RU.Querier.knn_query(K=100, points=pts, centers=ctr, points_padding=pad)
    d2, got = IO.knn(ctr, pts, None, 100)
    assert (ref.sort(-1)[0] == got.sort(-1)[0]).all(-1).float().mean() >= 0.99
    assert (d2[..., 1:] >= d2[..., :-1]).all()


def test_knn_oracle_any_k_against_a_dense_sort():
    """the C oracle's (d2, index) order for K up to the cloud size, against a stable sort of the full distance matrix"""
    g = torch.Generator().manual_seed(3)
    p2 = torch.randn(2, 500, 3, generator=g)
    p2[1, 250:] = p2[1, :250]              # exact duplicates: ties resolve to the lower index
    p1 = torch.randn(2, 20, 3, generator=g)
    for K in (65, 257, 500):
        d2, idx = IO.knn(p1, p2, None, K)
        for b in range(2):
            for s in range(20):
                dd = torch.tensor([IO.d2(p1[b, s].numpy(), p2[b, i].numpy()) for i in range(500)])
                order = sorted(range(500), key=lambda i: (float(dd[i]), i))[:K]
                assert idx[b, s].tolist() == order
                break  # one query per cloud keeps the pure-Python loop short
    d2, idx = IO.knn(p1, p2, torch.tensor([500, 70]), 100)   # K > lengths2: zero padded
    assert (idx[1, :, 70:] == 0).all() and (d2[1, :, 70:] == 0).all()


def test_knn_contract_small():
    p2 = torch.tensor([[[0., 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3], [5, 5, 5]]])
    p1 = torch.tensor([[[0.1, 0, 0], [0, 0, 2.9]]])
    d2, idx = IO.knn(p1, p2, None, 3)
    assert idx.tolist() == [[[0, 1, 2], [3, 0, 1]]]
    assert torch.allclose(d2[0, 0], torch.tensor([0.01, 0.81, 4.01]), atol=1e-6)
    d2, idx = IO.knn(p1, p2, torch.tensor([2]), 3)  # lengths2 < K: zero padded
    assert idx.tolist() == [[[0, 1, 0], [0, 1, 0]]] and d2[0, 0, 2] == 0
    h = IO.hybrid(p1, p2, None, 3, 1.0)
    assert h.tolist() == [[[0, 1, 0], [3, 3, 3]]]
    bd, bi = IO.ball_query(p1, p2, None, 3, 2.5)
    assert bi.tolist() == [[[0, 1, 2], [3, -1, -1]]]


def test_knn_ties_resolve_to_lower_index():
    p2 = torch.zeros(1, 40, 3)
    p2[0, 20:] = 1.0
    d2, idx = IO.knn(torch.zeros(1, 1, 3), p2, None, 8)
    assert idx[0, 0].tolist() == list(range(8))
    f = IO.fps(p2, None, 3)
    assert f[0].tolist() == [0, 20, 0]  # all min-distances 0 after two picks: first maximum = index 0


# ---- model restatement vs the reference --------------------------------------------------------
@needs_ref
@pytest.mark.reference
def test_encoder_restatement_matches_reference(reference, cfg):
    c = data.kitti_shape_cloud(11, 6000)
    pad = torch.zeros(1, 6000, dtype=torch.bool)
    with torch.no_grad():
        ref = reference["enc"](c[None], pad)
    got = M.encoder_forward(reference["ck"]["encoder"], cfg, c[None], pad, "fallback")
    assert torch.equal(ref[0], got[0]) and torch.equal(ref[2], got[2])
    assert rel_err(got[1], ref[1]) < 2e-6
    # direct-difference kNN (the pytorch3d contract) vs the reference's matmul-formula fallback: a few
    # rows differ by a boundary neighbour (SURVEY.md section 7 "kNN parity definition"), so descriptors
    # agree element-wise almost everywhere and exactly once the reference's indices are injected.
    tr = {}
    M.encoder_forward(reference["ck"]["encoder"], cfg, c[None], pad, "fallback", trace=tr)
    direct = M.encoder_forward(reference["ck"]["encoder"], cfg, c[None], pad, "direct")
    err = (direct[1] - ref[1]).abs() / ref[1].abs().max()
    assert (err < 1e-4).float().mean() > 0.9 and err.max() < 2e-2  # one swapped neighbour ripples downstream
    inj = M.encoder_forward(reference["ck"]["encoder"], cfg, c[None], pad, "direct",
                            inject={"knn_idx": tr["knn_idx"]})
    assert rel_err(inj[1], ref[1]) < 2e-6


@needs_ref
@pytest.mark.reference
def test_encoder_restatement_with_padding(reference, cfg):
    c = data.kitti_shape_cloud(12, 5000)
    pad = torch.zeros(1, 5000, dtype=torch.bool)
    pad[:, 4500:] = True
    with torch.no_grad():
        ref = reference["enc"](c[None], pad)
    got = M.encoder_forward(reference["ck"]["encoder"], cfg, c[None], pad, "fallback")
    assert torch.equal(ref[0], got[0]) and rel_err(got[1], ref[1]) < 2e-6


@needs_ref
@pytest.mark.reference
def test_decoder_restatement_matches_reference(reference, cfg, golden_sample):
    d0, d1 = torch.from_numpy(golden_sample["desc0"]), torch.from_numpy(golden_sample["desc1"])
    with torch.no_grad():
        R, T, conf, rmse = reference["dec"].registration_forward(d0, d1, num_sample=0.5)
        lp = reference["dec"].loop_detection_forward(torch.stack([d0, d1]), torch.stack([d1, d0]))
    R2, T2, conf2, rmse2 = M.registration_forward(reference["ck"]["decoder"], cfg, d0, d1, 0.5)
    assert conf.shape == conf2.shape
    assert (R - R2).abs().max() < 1e-6 and (T - T2).abs().max() < 1e-5
    assert (conf - conf2).abs().max() < 1e-5 and abs(rmse - rmse2) < 1e-6
    lp2 = M.loop_detection_forward(reference["ck"]["decoder"], cfg, torch.stack([d0, d1]), torch.stack([d1, d0]))
    assert (lp - lp2).abs().max() < 1e-5


# ---- oracle vs committed golden fixtures (runs on every box that has the checkpoint) -------------
def test_oracle_reproduces_golden_sample_pair(cfg, checkpoint, golden_sample):
    g = golden_sample
    c0 = torch.from_numpy(g["cloud0"])
    pad = torch.zeros(1, c0.shape[1], dtype=torch.bool)
    tr = {}
    desc = M.descriptors(checkpoint["encoder"], cfg, c0[None], pad, "direct", trace=tr)[0]
    for i in range(5):  # FPS indices: bit-exact
        assert np.array_equal(tr["fps_idx"][i][0].numpy().astype(np.int32), g[f"fps0_{i}"])
    for i in range(2, 11):  # small-level group indices: set-equal rows (the reference used the matmul formula)
        a = np.sort(tr["knn_idx"][i][0].numpy().astype(np.int32), -1)
        assert (a == np.sort(g[f"knn0_{i}"], -1)).all(-1).mean() >= 0.995
    for i in range(2):
        same = tr["knn_idx"][i][0].sum(-1).numpy() == g[f"knn0_{i}_rowsum"]
        assert same.mean() >= 0.995
    assert rel_err(desc, torch.from_numpy(g["desc0"])) < 1e-4
    R, T, conf, rmse = M.registration_forward(checkpoint["decoder"], cfg, torch.from_numpy(g["desc0"]),
                                              torch.from_numpy(g["desc1"]), 0.5)
    assert np.abs(R.numpy() - g["R"]).max() < 1e-6 and np.abs(T.numpy() - g["T"]).max() < 1e-5
    assert len(conf) == len(g["conf"]) == 251 and abs(rmse - float(g["rmse"])) < 1e-6
    # the survey's sanity anchor
    assert abs(float(T[0]) + 0.107) < 2e-3 and abs(float(conf[0]) - 0.976) < 2e-3 and abs(rmse - 0.0376) < 1e-3


def test_oracle_reproduces_golden_synthetic(cfg, checkpoint, golden_synth):
    g = golden_synth
    c = data.kitti_shape_cloud(seed=3, n=8192)
    assert abs(c.double().sum().item() - float(g["cloud_checksum"])) < 1e-9
    pad = torch.zeros(1, 8192, dtype=torch.bool)
    tr = {}
    desc = M.descriptors(checkpoint["encoder"], cfg, c[None], pad, "direct", trace=tr)[0]
    for i in range(5):
        assert np.array_equal(tr["fps_idx"][i][0].numpy().astype(np.int32), g[f"fps_{i}"])
    assert rel_err(desc, torch.from_numpy(g["desc"])) < 1e-4
    R, T, conf, rmse = M.registration_forward(checkpoint["decoder"], cfg, torch.from_numpy(g["desc"]),
                                              torch.from_numpy(g["desc_moved"]), 0.5)
    assert np.abs(R.numpy() - g["R"]).max() < 1e-6 and np.abs(T.numpy() - g["T"]).max() < 1e-5
    assert len(conf) == len(g["conf"])


# ---- information matrix -----------------------------------------------------
@needs_ref
@pytest.mark.reference
def test_information_matrix_matches_reference():
    """oracle/infomat_ref.py vs the reference function itself (CPU knn_points stand-in for pytorch3d)."""
    from oracle import infomat_ref
    from ref_infomat import cases, reference_information_matrix
    for name, (p1, p2, T, radius) in cases().items():
        want = reference_information_matrix(p1, p2, T, radius)
        got, n = infomat_ref.information_matrix(p1, p2, T, radius)
        assert got.shape == (6, 6)
        if name == "far_apart":
            assert n == 0 and float(want.abs().max()) == 0.0 and float(got.abs().max()) == 0.0
        else:
            assert n > 100
            assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max()), name


def test_information_matrix_oracle_matches_golden():
    from oracle import infomat_ref
    g = np.load(os.path.join(ROOT, "tests", "golden", "infomat.npz"))
    for name in ("kitti6k", "uniform", "far_apart"):
        got, _ = infomat_ref.information_matrix(torch.from_numpy(g[name + "_p1"]), torch.from_numpy(g[name + "_p2"]),
                                                torch.from_numpy(g[name + "_T"]), float(g[name + "_radius"]))
        want = torch.from_numpy(g[name + "_info"])
        assert float((got - want).abs().max()) <= 1e-5 * max(float(want.abs().max()), 1e-30), name


# ---- preprocessing front-end ---------------------------------------------------
def _reference_transforms():
    """dataloader/transforms.py of the reference with a stub `open3d` (only its filters need the real one)"""
    import importlib
    import types
    saved = {k: sys.modules.get(k) for k in ("open3d", "pytorch3d")}
    if saved["open3d"] is None:
        sys.modules["open3d"] = types.ModuleType("open3d")
    sys.modules["pytorch3d"] = None
    sys.path.insert(0, REF)
    try:
        RT = importlib.import_module("dataloader.transforms")
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return RT


@needs_ref
@pytest.mark.reference
def test_frontend_matches_reference_transforms():
    """oracle/frontend_ref.py vs the reference's own BinReader rule + VoxelSample('first') + DistanceSample +
    CoordinatesNormalization on a shipped sample frame: identical points in identical order."""
    from oracle import frontend_ref
    RT = _reference_transforms()
    raw = np.fromfile(f"{REF}/data/sample/seq06/velodyne/000003.bin", dtype=np.float32).reshape(-1, 4)
    xyz = raw[:, :3]
    xyz = xyz[np.isnan(xyz).sum(1) == 0]                      # heads/bin.py:16-17
    pcd = RT.PointCloud(xyz.copy())
    for t in (RT.VoxelSample(0.3, "first"), RT.DistanceSample(1, 60), RT.CoordinatesNormalization(60)):
        pcd = t(pcd)
    want = pcd.xyz.T.contiguous()
    got = frontend_ref.preprocess_bin(raw)
    assert got.shape == want.shape and got.shape[1] > 10000
    assert torch.equal(got, want)


def test_frontend_oracle_matches_golden():
    from oracle import frontend_ref
    g = np.load(os.path.join(ROOT, "tests", "golden", "frontend.npz"))
    got = frontend_ref.preprocess_bin(g["raw"])
    assert torch.equal(got, torch.from_numpy(g["out"]))


# ---- scan-to-map input stage -----------------------------------------------------
def _map_tile_cases():
    g = torch.Generator().manual_seed(17)
    import math
    kps, poses = [], []
    for i in range(5):
        kp = torch.randn(131, 256, generator=g)
        kp[-3:] = kp[-3:] * 20.0
        a = 0.1 * i
        T = torch.eye(4)
        T[0, 0], T[0, 1], T[1, 0], T[1, 1] = math.cos(a), -math.sin(a), math.sin(a), math.cos(a)
        T[:3, 3] = torch.tensor([2.0 * i, -0.5 * i, 0.1 * i])
        kps.append(kp)
        poses.append(T)
    c = torch.eye(4)
    c[0, 0], c[0, 1], c[1, 0], c[1, 1] = math.cos(0.3), -math.sin(0.3), math.sin(0.3), math.cos(0.3)
    c[:3, 3] = torch.tensor([4.0, 1.0, -0.2])
    return kps, poses, c


@needs_ref
@pytest.mark.reference
def test_map_tile_matches_reference_posegraph():
    """oracle/maptile_ref.py vs PoseGraph.__global_mapping + the centring lines of global_map_query_graph"""
    import importlib
    import types
    from oracle import maptile_ref
    from ref_infomat import reference_module
    reference_module()  # leaves the reference's system.modules.utils importable (stub open3d / matplotlib)
    rw = types.ModuleType("readerwriterlock")

    class _Lock:
        def acquire(self, blocking=True):
            return True

        def release(self):
            pass

    class _RW:
        def gen_wlock(self):
            return _Lock()

        def gen_rlock(self):
            return _Lock()

    rw.rwlock = types.SimpleNamespace(RWLockFair=_RW)
    saved = sys.modules.get("readerwriterlock")
    sys.modules["readerwriterlock"] = rw
    sys.path[:0] = [os.path.join(ROOT, "deeppointmap_b200", "compat"), REF]
    sys.path.append(os.path.join(ROOT, "deeppointmap_b200", "compat_shims"))
    try:
        PG = importlib.import_module("system.modules.pose_graph")
    finally:
        sys.path.remove(REF)
        if saved is None:
            sys.modules.pop("readerwriterlock", None)
    kps, poses, center = _map_tile_cases()
    pg = PG.PoseGraph(args=None, agent_id=0, device="cpu")
    scans = []
    for i, (kp, T) in enumerate(zip(kps, poses)):
        sp = PG.ScanPack(timestamp=float(i), agent_id=0, timestep=i, key_points=kp, SE3_pred=T, coor_sys=0)
        pg.add_vertex(sp)
        scans.append(sp)
    tile, tokens = pg._PoseGraph__global_mapping(scans, full_pcd=False)
    R, t = PG.PoseTool.Rt(center)                                # pose_graph.py:505-507
    tile[-3:, :] = R.T @ (tile[-3:, :] - t)
    got = maptile_ref.map_tile(kps, poses, center)
    assert tokens.tolist() == [i for i in range(5) for _ in range(256)]
    assert torch.equal(got, tile)


# ---- OutlierFilter (second half) ---------------------------------------------------
@needs_ref
@pytest.mark.reference
def test_outlier_filter_matches_reference():
    """oracle/outlier_ref.py vs the reference's OutlierFilter class, CUDA branch (transforms.py:236-246) forced
    onto CPU tensors: `pcd.device` reads 'cuda:fake', knn_points is a plain-torch stand-in."""
    from oracle import frontend_ref, outlier_ref
    from ref_infomat import _cpu_knn_points
    RT = _reference_transforms()
    RT.has_t3d, RT.knn_points = True, _cpu_knn_points
    raw = np.fromfile(f"{REF}/data/sample/seq06/velodyne/000005.bin", dtype=np.float32).reshape(-1, 4)
    xyz = (frontend_ref.preprocess_bin(raw) * 60.0).T.contiguous()[:6000]     # metres, after voxel + distance sampling
    pcd = RT.PointCloud(xyz.numpy().copy())
    pcd.device = "cuda:fake"
    out = RT.OutlierFilter(nb_neighbors=10, std_ratio=3.0)(pcd)
    kept, mask, stat, thr = outlier_ref.outlier_filter(xyz, 10, 3.0)
    assert 0 < int((~mask).sum()) < 600
    assert torch.equal(out.xyz, kept)


@needs_ref
@pytest.mark.reference
def test_decoder_key_padding_masks_match_reference(reference):
    """the masks reach the attention only (descriptor_attention.py:33-42); restated in model_ref._mha"""
    dec, sd = reference["dec"], reference["ck"]["decoder"]
    cfg = M.default_config()
    g = torch.Generator().manual_seed(5)
    src, dst = torch.randn(131, 96, generator=g), torch.randn(131, 80, generator=g)
    src[128:] *= 20; dst[128:] *= 20
    sp, dp = torch.rand(96, generator=g) < 0.25, torch.rand(80, generator=g) < 0.25
    with torch.no_grad():
        R, T, c, r = dec.registration_forward(src, dst, sp.view(1, -1), dp.view(1, -1), num_sample=0.5)
        lp = dec.loop_detection_forward(src[None], dst[None], sp.view(1, -1), dp.view(1, -1))
    Rw, Tw, cw, rw = M.registration_forward(sd, cfg, src, dst, 0.5, s_pad=sp, d_pad=dp)
    assert c.shape == cw.shape and (R - Rw).abs().max() < 1e-4 and (T - Tw).abs().max() < 1e-3 and (c - cw).abs().max() < 1e-4
    lw = M.loop_detection_forward(sd, cfg, src[None], dst[None], sp.view(1, -1), dp.view(1, -1))
    assert (lp - lw).abs().max() < 1e-5


@needs_ref
@pytest.mark.reference
def test_low_pass_filter_matches_reference():
    """oracle/lowpass_ref.py vs the reference's LowPassFilter class (transforms.py:256-297) run on CPU: open3d is the
    numpy / scipy stand-in of compat_shims (the same normal algorithm the oracle restates), knn_points / knn_gather are
    plain-torch stand-ins of the pytorch3d contract.  Same rows kept."""
    import importlib
    from oracle import frontend_ref, lowpass_ref
    from ref_infomat import _cpu_knn_points
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "dpm_open3d_shim", os.path.join(ROOT, "deeppointmap_b200", "compat_shims", "open3d", "__init__.py"))
    shim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shim)
    shim.open3d = shim                          # the reference spells o3d.open3d.utility (transforms.py:270)
    RT = _reference_transforms()
    RT.o3d = shim
    RT.has_t3d, RT.knn_points = True, _cpu_knn_points
    RT.knn_gather = lambda x, idx: x[0][idx[0]][None]           # (1,N,C), (1,N,K) -> (1,N,K,C)
    raw = np.fromfile(f"{REF}/data/sample/seq06/velodyne/000004.bin", dtype=np.float32).reshape(-1, 4)
    xyz = (frontend_ref.preprocess_bin(raw) * 60.0).T.contiguous()[:5000]     # metres, after voxel + distance sampling
    pcd = RT.PointCloud(xyz.numpy().copy())
    out = RT.LowPassFilter(normals_radius=0.5, normals_num=16, filter_std=2.0, flux=4, max_remain=-1)(pcd)
    kept, mask, sim, thr = lowpass_ref.low_pass_filter(xyz, 0.5, 16, 2.0, 4)
    assert 0 < int((~mask).sum()) < 1500
    assert torch.equal(out.xyz, kept)


@needs_ref
def test_bias_false_state_dict_matches_the_reference(reference):
    """`encoder.bias: False` (network/encoder/encoder.py:22, utils.py:358-389): the drop-in holds the same parameters
    as the reference built with that flag -- no conv biases except the stem's -- and still fills a full pointer table."""
    import copy
    import yaml
    from easydict import EasyDict
    from network.encoder.encoder import Encoder as RefEncoder
    from deeppointmap_b200 import Encoder
    cfg = EasyDict(yaml.safe_load(open(f"{REF}/configs/infer/DeepPointMap_B_Main_SemanticKITTI.yaml")))
    cfg.encoder.bias = False
    ref = RefEncoder(copy.deepcopy(cfg))
    ours = Encoder(copy.deepcopy(cfg))
    rs, os_ = ref.state_dict(), ours.state_dict()
    assert list(rs.keys()) == list(os_.keys())
    assert all(rs[k].shape == os_[k].shape for k in rs)
    assert "point_mlp0.bias" in os_ and "downsampler.0.sa.mlp.0.bias" not in os_
    ours.load_state_dict(rs, strict=True)
    table = ours._ordered_params()
    full = Encoder(EasyDict(yaml.safe_load(open(f"{REF}/configs/infer/DeepPointMap_B_Main_SemanticKITTI.yaml"))))
    assert len(table) == len(full._ordered_params())
    assert [tuple(t.shape) for t in table] == [tuple(t.shape) for t in full._ordered_params()]
    assert list(ours.state_dict().keys()) == list(rs.keys())  # the zero vectors stay out of the state_dict
