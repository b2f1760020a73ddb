"""GPU parity of the decoder (registration / loop detection through the C-ABI) vs the oracle.
Bar: R, T, conf, rmse within 1e-4 relative; same correspondences selected."""
import ctypes
import os
import math

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import model_ref as M
from deeppointmap_b200 import Decoder, _C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _descs(seed, L, P=None, spread=20.0):
    g = torch.Generator().manual_seed(seed)
    shape = (131, L) if P is None else (P, 131, L)
    d = torch.randn(*shape, generator=g)
    d[..., 128:, :] *= spread
    return d


def _moved(desc, yaw_deg=3.0, t=(0.8, -0.3, 0.1), noise=0.02, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = math.radians(yaw_deg)
    R = torch.tensor([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1.0]])
    out = desc.clone()
    out[128:] = R @ desc[128:] + torch.tensor(t).view(3, 1)
    out[:128] += noise * torch.randn(128, desc.shape[1], generator=g)
    perm = torch.randperm(desc.shape[1], generator=g)
    return out[:, perm].contiguous()


def _dec(cfg, sd):
    d = Decoder(cfg).eval()
    d.load_state_dict(sd, strict=True)
    return d.to(DEV)


def test_posenc(cfg):
    xyz = (torch.rand(500, 3, generator=torch.Generator().manual_seed(0)) - 0.5) * 120.0
    want = M.pos_embedding(xyz[None], 256)[0]
    d = Decoder(cfg)
    dim_t = d._dim_t(torch.device(DEV))
    x = xyz.to(DEV)
    out = torch.empty(500, 256, device=DEV)
    _C.check(_C.lib().dpm_posenc_f32(x.data_ptr(), 3, dim_t.data_ptr(), 84, out.data_ptr(), 500, 256, _C.stream_ptr()))
    assert (out.cpu() - want).abs().max() < 2e-5  # |arg| up to ~190 rad: 1 ulp of the argument
    assert (out[:, 252:] == 0).all()


@pytest.mark.parametrize("Lq,Lk", [(256, 256), (100, 37), (33, 700), (512, 300)])
def test_attention_core(Lq, Lk):
    g = torch.Generator().manual_seed(Lq + Lk)
    H = 8
    q, k, v = torch.randn(Lq, 256, generator=g), torch.randn(Lk, 256, generator=g), torch.randn(Lk, 256, generator=g)
    qq, kk, vv = (t.view(-1, H, 32).transpose(0, 1).double() for t in (q, k, v))
    ref = (torch.softmax(qq @ kk.transpose(1, 2) / math.sqrt(32), -1) @ vv).transpose(0, 1).reshape(Lq, 256)
    qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
    out = torch.empty(Lq, 256, device=DEV)
    prob = torch.tensor([0, Lq, 0, Lk], dtype=torch.int32, device=DEV)
    _C.check(_C.lib().dpm_attention_f32(qd.data_ptr(), 256, kd.data_ptr(), 256, vd.data_ptr(), 256, out.data_ptr(), 256,
                                        prob.data_ptr(), 1, Lq, H, _C.stream_ptr()))
    assert rel_err(out, ref) < 1e-5


def test_kabsch_vs_oracle():
    g = torch.Generator().manual_seed(3)
    P, K = 4, 300
    src = torch.randn(P, 3, K, generator=g) * 10
    a = 0.3
    R = torch.tensor([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1.0]])
    dst = R @ src + torch.tensor([1.0, 2.0, 3.0]).view(1, 3, 1) + 0.05 * torch.randn(P, 3, K, generator=g)
    dst[:, :, ::17] += 3.0  # outliers
    w = torch.rand(P, K, generator=g)
    cnt = torch.tensor([300, 250, 64, 31], dtype=torch.int32)
    res = torch.zeros(P, 16, device=DEV)
    inl = torch.zeros(P, K, dtype=torch.uint8, device=DEV)
    conf = torch.zeros(P, K, device=DEV)
    s, d_, ww, c = src.to(DEV), dst.to(DEV), w.to(DEV), cnt.to(DEV)
    _C.check(_C.lib().dpm_kabsch_f32(s.data_ptr(), d_.data_ptr(), ww.data_ptr(), c.data_ptr(), P, K, res.data_ptr(),
                                     inl.data_ptr(), conf.data_ptr(), _C.stream_ptr()))
    res, inl, conf = res.cpu(), inl.cpu().bool(), conf.cpu()
    for p in range(P):
        n = int(cnt[p])
        # ties in the top-64 are impossible here (continuous random weights)
        Rw, Tw, iw, rm = M.solve_svd(w[p, :n].clone(), src[p, :, :n], dst[p, :, :n])
        assert torch.equal(inl[p, :n], iw)
        assert (res[p, 0:9].view(3, 3) - Rw).abs().max() < 1e-5
        assert (res[p, 9:12].view(3, 1) - Tw).abs().max() < 1e-4
        assert abs(float(res[p, 12]) - rm) < 1e-4 * max(1.0, rm)
        assert int(res[p, 13]) == n and int(res[p, 14]) == int(iw.sum())
        assert torch.equal(conf[p, :int(iw.sum())], w[p, :n][iw])


def _check_registration(dec, sd, cfg, src, dst, num_sample=0.5, conf_vs_fp64=False):
    tr = {}
    Rw, Tw, cw, rw = M.registration_forward(sd, cfg, src, dst, num_sample, trace=tr)
    R, T, c, r = dec.registration_forward(src.to(DEV), dst.to(DEV), num_sample=num_sample)
    assert R.shape == (3, 3) and T.shape == (3, 1) and isinstance(r, float)
    if conf_vs_fp64 and c.shape != cw.shape:
        # Map sizes: > 1000 correspondences go through two hard thresholds (|offset|^2 <= eps^2, err <= mean + 3 sigma); one
        # that sits on a threshold can fall either way under ANY fp32 summation order (the fp32 reference against its own
        # fp64 evaluation included).  Up to 2 such flips are tolerated; the pose, which averages over all of them, and
        # every confidence that both sides kept must still agree.
        assert abs(c.shape[0] - cw.shape[0]) <= 2 and cw.shape[0] > 500, f"inlier count {c.shape} vs oracle {cw.shape}"
        assert (R.cpu() - Rw).abs().max() < TOL and (T.cpu() - Tw).abs().max() < TOL * max(1.0, float(Tw.abs().max()))
        assert abs(r - rw) < 2e-3 * max(1.0, rw)
        a, b = torch.sort(c.cpu())[0], torch.sort(cw)[0]
        short, long_ = (a, b) if a.numel() < b.numel() else (b, a)
        idx = torch.searchsorted(long_, short).clamp(1, long_.numel() - 1)
        near = torch.minimum((long_[idx] - short).abs(), (long_[idx - 1] - short).abs())
        assert float(near.max()) < 3 * TOL
        return R, T, c, r
    assert c.shape == cw.shape, f"inlier count {c.shape} vs oracle {cw.shape}"
    assert (R.cpu() - Rw).abs().max() < TOL
    assert (T.cpu() - Tw).abs().max() < TOL * max(1.0, float(Tw.abs().max()))
    if conf_vs_fp64:
        # A confidence is exp(2 s / tau - lse_row - lse_col) with tau = 0.1: its condition number w.r.t. the cosine s
        # is 20, and with thousands of map descriptors the fp32 reference itself sits 3e-5 .. 1.5e-4 away from an
        # fp64 evaluation of the same formula (tools/probe_conf.py).  The bar at map sizes is therefore "as close to
        # the fp64 truth as the fp32 reference is", not a number below the reference's own rounding noise.
        sd64 = {k: v.double() for k, v in sd.items()}
        s6, d6, _, _ = M.attention_forward(sd64, cfg, src[None].double(), dst[None].double())
        P6 = M.pairing(sd64, cfg, s6, d6, num_sample)[3]
        c64 = P6[tr["src_index"], tr["dst_index"]].repeat(2)[tr["keep"]][tr["inlier"]]
        err_ref = float((cw.double() - c64).abs().max())
        err_gpu = float((c.cpu().double() - c64).abs().max())
        assert err_gpu < max(TOL, 1.5 * err_ref), f"confidence vs fp64: gpu {err_gpu:.2e}, fp32 reference {err_ref:.2e}"
    else:
        assert (c.cpu() - cw).abs().max() < TOL
    assert abs(r - rw) < TOL * max(1.0, rw)
    return R, T, c, r


def test_registration_real_weights_golden_pair(cfg, checkpoint, golden_sample):
    """BASELINE config 3 on the reference-generated descriptors of sample frames 0 / 1."""
    dec = _dec(cfg, checkpoint["decoder"])
    d0, d1 = torch.from_numpy(golden_sample["desc0"]), torch.from_numpy(golden_sample["desc1"])
    R, T, c, r = _check_registration(dec, checkpoint["decoder"], cfg, d0, d1)
    assert np.abs(R.cpu().numpy() - golden_sample["R"]).max() < TOL
    assert np.abs(T.cpu().numpy() - golden_sample["T"]).max() < TOL
    assert len(c) == len(golden_sample["conf"]) == 251
    assert np.abs(c.cpu().numpy() - golden_sample["conf"]).max() < TOL
    assert abs(r - float(golden_sample["rmse"])) < TOL


def test_registration_real_weights_synthetic_golden(cfg, checkpoint, golden_synth):
    dec = _dec(cfg, checkpoint["decoder"])
    d0, d1 = torch.from_numpy(golden_synth["desc"]), torch.from_numpy(golden_synth["desc_moved"])
    R, T, c, r = _check_registration(dec, checkpoint["decoder"], cfg, d0, d1)
    assert np.abs(T.cpu().numpy() - golden_synth["T"]).max() < TOL * 2


def test_registration_scan_to_map_shape(cfg, checkpoint, golden_sample):
    """M = 1024 map tokens vs N = 256 (scan-to-map geometry, mapping.py:153)."""
    dec = _dec(cfg, checkpoint["decoder"])
    d0, d1 = torch.from_numpy(golden_sample["desc0"]), torch.from_numpy(golden_sample["desc1"])
    big = torch.cat([d0, _moved(d0, 1.0, (5, 5, 0), seed=1), _moved(d1, 0.0, (-9, 4, 0), seed=2), d1], dim=1)
    _check_registration(dec, checkpoint["decoder"], cfg, big, d1)


def test_registration_random_weights_batched_matches_single(cfg):
    sd = M.random_weights(M.decoder_shapes(cfg), seed=7)
    dec = _dec(cfg, sd)
    src = torch.stack([_descs(1, 256), _descs(2, 256), _descs(3, 256)])
    dst = torch.stack([_moved(src[0]), _moved(src[1], seed=5), _descs(9, 256)])
    res, conf = dec.registration_forward_batch(src.to(DEV), dst.to(DEV), 0.5)
    for p in range(3):
        R, T, c, r = dec.registration_forward(src[p].to(DEV), dst[p].to(DEV), num_sample=0.5)
        assert torch.equal(res[p, 0:9].view(3, 3), R) and torch.equal(res[p, 9:12].view(3, 1), T)
        assert int(res[p, 14]) == len(c) and torch.equal(conf[p, :len(c)], c)


def test_registration_batched_api_shapes(cfg):
    sd = M.random_weights(M.decoder_shapes(cfg), seed=8)
    dec = _dec(cfg, sd)
    s, d_ = _descs(1, 256, P=1).to(DEV), _descs(2, 256, P=1).to(DEV)
    R, T, c, r = dec.registration_forward(s, d_, num_sample=0.5)  # 3-D in -> batched out (decoder.py:121-126)
    assert R.shape == (1, 3, 3) and T.shape == (1, 3, 1) and c.dim() == 2 and isinstance(r, list)
    with pytest.raises(AssertionError):
        dec.registration_forward(_descs(1, 64, P=2).to(DEV), _descs(2, 64, P=2).to(DEV))
    with pytest.raises(ValueError):
        dec.registration_forward(s[0], d_[0], num_sample=-1.0)


def test_loop_detection(cfg, checkpoint, golden_sample):
    dec = _dec(cfg, checkpoint["decoder"])
    d0, d1 = torch.from_numpy(golden_sample["desc0"]), torch.from_numpy(golden_sample["desc1"])
    far = _descs(5, 256)
    S, D = torch.stack([d0, d1, d0]), torch.stack([d1, d1, far])
    want = M.loop_detection_forward(checkpoint["decoder"], cfg, S, D)
    got = dec.loop_detection_forward(S.to(DEV), D.to(DEV))
    assert got.shape == (3,) and (got.cpu() - want).abs().max() < TOL
    assert np.abs(got[:2].cpu().numpy() - golden_sample["loop"]).max() < TOL


def test_map_tile_and_scan_to_map_registration(cfg, checkpoint):
    """device-resident descriptor store -> map tile (vs oracle, 1e-5 of the coordinate range) ->
    scan-to-map registration (M = 256 scan descriptors against N = 4 x 256 map descriptors)."""
    import sys
    from oracle import maptile_ref
    from deeppointmap_b200 import Encoder, ops, sequence
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_oracle_pin import _map_tile_cases
    kps, poses, center = _map_tile_cases()
    store = torch.stack(kps).to(DEV)
    want = maptile_ref.map_tile(kps, poses, center)
    got = ops.map_tile(store, [0, 1, 2, 3, 4], torch.stack(poses), center).cpu()
    assert got.shape == want.shape == (131, 5 * 256)
    assert torch.equal(got[:128], want[:128])
    assert float((got[128:] - want[128:]).abs().max()) <= 1e-5 * float(want[128:].abs().max())
    sub = ops.map_tile(store, [3, 1], torch.stack([poses[3], poses[1]])).cpu()   # any subset / order, no centring
    assert float((sub - maptile_ref.map_tile([kps[3], kps[1]], [poses[3], poses[1]])).abs().max()) <= 1e-4

    # scan-to-map on real descriptors: 5 scans along a corridor, map = scans 0..3 in the frame of scan 3, query = scan 4
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    enc.load_state_dict(checkpoint["encoder"], strict=True)
    dec.load_state_dict(checkpoint["decoder"], strict=True)
    enc, dec = enc.to(DEV), dec.to(DEV)
    gt = sequence.trajectory(5)
    world = sequence.corridor_world(4, length_m=float(gt[-1, 0, 3]), device=DEV)
    frames = sequence.corridor_frames(world, gt, 16384, seed=5)
    desc = enc.descriptors(frames, None, coor_scale=cfg.coor_scale)
    tile = ops.map_tile(desc, [0, 1, 2, 3], gt[:4].float(), gt[3].float())
    R, T, conf, rmse = dec.registration_forward(desc[4], tile, num_sample=0.5)
    want_rel = torch.linalg.inv(gt[3]) @ gt[4]                    # scan 4 expressed in the map's (scan 3's) frame
    assert float((T.flatten().cpu().double() - want_rel[:3, 3]).norm()) < 0.6
    Rw, Tw, cw, rw = M.registration_forward(checkpoint["decoder"], cfg, desc[4].cpu(), tile.cpu(), 0.5)
    assert float((R.cpu() - Rw).abs().max()) < 1e-4 and float((T.cpu() - Tw).abs().max()) < 1e-4 * max(1.0, float(Tw.abs().max()))


def _map_of(blocks, seeds):
    """a map tile the way PoseGraph.__global_mapping concatenates key-point sets (pose_graph.py:373-409)"""
    out = []
    for i, (b, s) in enumerate(zip(blocks, seeds)):
        out.append(_moved(b, yaw_deg=0.7 * i, t=(1.5 * i, 0.2 * i, 0.0), noise=0.01, seed=s))
    return torch.cat(out, dim=1)


@pytest.mark.parametrize("m_blocks,n_blocks", [(16, 1), (16, 16), (1, 16), (7, 3)])
def test_registration_caller_sizes(cfg, checkpoint, golden_sample, m_blocks, n_blocks):
    """The sizes the SLAM callers use: graph_search(max_k=16) bounds a map tile at 16 x 256 = 4096 descriptors
    (pose_graph.py:513); scan-to-map is M=4096 vs N=256 (mapping.py:153), map-to-map M=N=4096
    (loop_closure.py:240) -> k = 1088 / 2048 pairs, 2k = 4096 Kabsch rows at the limit of the kernels."""
    dec = _dec(cfg, checkpoint["decoder"])
    d0, d1 = torch.from_numpy(golden_sample["desc0"]), torch.from_numpy(golden_sample["desc1"])
    src = _map_of([d0 if i % 2 == 0 else d1 for i in range(m_blocks)], range(100, 100 + m_blocks)) if m_blocks > 1 else d0
    dst = _map_of([d1 if i % 2 == 0 else d0 for i in range(n_blocks)], range(200, 200 + n_blocks)) if n_blocks > 1 else d1
    assert src.shape[1] == 256 * m_blocks and dst.shape[1] == 256 * n_blocks
    _check_registration(dec, checkpoint["decoder"], cfg, src, dst, conf_vs_fp64=True)


@pytest.mark.parametrize("C,Ms,Ns", [(8, 256, 256), (5, 256, 192), (13, 100, 256)])
def test_loop_detection_batched_unsaturated(cfg, C, Ms, Ns):
    """loop_closure.py:171: C candidates in one call.  Random weights keep the logits away from the sigmoid's flat
    ends, so the 1e-4 bar is a real check of the whole stack (the trained head saturates at 0.999999)."""
    sd = M.random_weights(M.decoder_shapes(cfg), seed=11)
    dec = _dec(cfg, sd)
    S = _descs(31, Ms, P=C)
    D = torch.stack([_moved(S[i], seed=i)[:, :Ns] if Ns <= Ms else _descs(50 + i, Ns) for i in range(C)])
    want = M.loop_detection_forward(sd, cfg, S, D)
    got = dec.loop_detection_forward(S.to(DEV), D.to(DEV))
    assert got.shape == (C,)
    assert float(want.min()) > 1e-3 and float(want.max()) < 1 - 1e-3, "test data saturates the sigmoid"
    assert (got.cpu() - want).abs().max() < TOL


def test_key_padding_masks(cfg, checkpoint, golden_sample):
    """MT-mode batches are padded (`padding_to`, system/core.py:141-169) and the decoder gets key-padding masks
    (descriptor_attention.py:33-42): padded descriptors are ignored as attention keys and nowhere else."""
    sd = checkpoint["decoder"]
    dec = _dec(cfg, sd)
    d0, d1 = torch.from_numpy(golden_sample["desc0"]), torch.from_numpy(golden_sample["desc1"])
    g = torch.Generator().manual_seed(2)
    sp = torch.zeros(256, dtype=torch.bool); sp[200:] = True                    # a padded tail ...
    dp = torch.rand(256, generator=g) < 0.2                                     # ... and scattered padding
    src, dst = d0.clone(), d1.clone()
    src[:, sp] = 0.0
    Rw, Tw, cw, rw = M.registration_forward(sd, cfg, src, dst, 0.5, s_pad=sp, d_pad=dp)
    R, T, c, r = dec.registration_forward(src.to(DEV), dst.to(DEV), sp.to(DEV), dp.to(DEV), num_sample=0.5)
    assert c.shape == cw.shape
    assert (R.cpu() - Rw).abs().max() < TOL and (T.cpu() - Tw).abs().max() < TOL * max(1.0, float(Tw.abs().max()))
    assert (c.cpu() - cw).abs().max() < TOL and abs(r - rw) < TOL * max(1.0, rw)
    # the masks matter: without them the answer differs
    R0, T0, c0, r0 = dec.registration_forward(src.to(DEV), dst.to(DEV), num_sample=0.5)
    assert c0.shape != c.shape or (c0 - c).abs().max() > 1e-3
    # loop head, batched, random weights (unsaturated), one side masked only
    sdr = M.random_weights(M.decoder_shapes(cfg), seed=11)
    decr = _dec(cfg, sdr)
    S, D = _descs(31, 256, P=4), _descs(32, 192, P=4)
    spm = torch.rand(4, 256, generator=g) < 0.3
    want = M.loop_detection_forward(sdr, cfg, S, D, spm, None)
    got = decr.loop_detection_forward(S.to(DEV), D.to(DEV), spm.to(DEV), None)
    assert (got.cpu() - want).abs().max() < TOL
    assert (decr.loop_detection_forward(S.to(DEV), D.to(DEV)).cpu() - want).abs().max() > 1e-5
    with pytest.raises(ValueError):
        dec.registration_forward(src.to(DEV), dst.to(DEV), sp[:100].to(DEV), None)


@pytest.mark.parametrize("Mq,Nk,mode", [(4096, 256, 0), (4096, 256, 1), (1000, 1500, 1), (130, 700, 0), (4096, 4096, 1),
                                        (256, 256, 1), (100, 37, 0), (33, 65, 1), (1, 5, 1)])
@pytest.mark.parametrize("impl", [1, 2])
def test_attention_pairs_long(Mq, Nk, mode, impl):
    """both attention kernels (1: mma.sync with fresh per-tile accumulators, 2: tcgen05 / TMEM flash attention) on the
    decoder's pair layout at map sizes, against fp64; with and without a key-padding mask"""
    g = torch.Generator().manual_seed(Mq + Nk + mode)
    H, R = 8, Mq + Nk
    q, k, v = (torch.randn(R, 256, generator=g) for _ in range(3))
    mask = torch.rand(R, generator=g) < 0.1
    for km in (None, mask):
        def ref(qs, ks, vs, kmask):
            qq, kk, vv = (t.view(-1, H, 32).transpose(0, 1).double() for t in (qs, ks, vs))
            att = qq @ kk.transpose(1, 2) / math.sqrt(32)
            if kmask is not None:
                att = att.masked_fill(kmask.view(1, 1, -1), float("-inf"))
            return (torch.softmax(att, -1) @ vv).transpose(0, 1).reshape(-1, 256)
        src, dst = slice(0, Mq), slice(Mq, R)
        ksrc, kdst = (dst, src) if mode else (src, dst)
        want = torch.cat([ref(q[src], k[ksrc], v[ksrc], None if km is None else km[ksrc]),
                          ref(q[dst], k[kdst], v[kdst], None if km is None else km[kdst])])
        qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
        out = torch.full((R, 256), float("nan"), device=DEV)
        kmd = None if km is None else km.to(DEV).to(torch.uint8)
        _C.check(_C.lib().dpm_attention_pairs_f32(qd.data_ptr(), 256, kd.data_ptr(), 256, vd.data_ptr(), 256, out.data_ptr(), 256,
                                                  1, Mq, Nk, mode, H, _C.ptr(kmd), impl, _C.stream_ptr()))
        assert rel_err(out, want) < 3e-6, (impl, km is not None, rel_err(out, want))
