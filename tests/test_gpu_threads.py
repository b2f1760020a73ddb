"""The reference's multi-thread mode calls the encoder from one host thread and the decoder from three others
(system/core.py:93-103), multi-agent mode runs four model pairs on four threads (infer_multiagents.py:100-113).
The library keeps its per-call state (pre-split weight registry, forked streams, workspaces, error string,
launch profile) thread-local: concurrent calls from several threads and streams must give bit-identical
results to the same calls made one after another."""
import copy
import threading

import pytest
import torch

from deeppointmap_b200 import Decoder, Encoder, data

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_concurrent_threads_match_sequential(cfg, checkpoint):
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    enc.load_state_dict(checkpoint["encoder"], strict=True)
    dec.load_state_dict(checkpoint["decoder"], strict=True)
    enc, dec = enc.to(DEV), dec.to(DEV)
    models = [(enc, dec)] + [(copy.deepcopy(enc), copy.deepcopy(dec)) for _ in range(3)]  # thread 0 shares, 1-3 own copies
    clouds = [torch.stack([data.kitti_shape_cloud(10 * t + i, 8192) for i in range(3)]).to(DEV) for t in range(4)]

    def work(t, out):
        e, d = models[t]
        st = torch.cuda.Stream(device=DEV)
        with torch.cuda.stream(st), torch.no_grad():
            res = []
            for rep in range(3):
                desc = e.descriptors(clouds[t], None, coor_scale=cfg.coor_scale)
                r, conf = d.registration_forward_batch(desc[:2], desc[1:], 0.5)
                prob = d.loop_detection_forward(desc[:2], desc[1:])
                res.append((desc.clone(), r.clone(), prob.clone()))
            st.synchronize()
        out[t] = res

    seq = {}
    for t in range(4):
        work(t, seq)
    par = {}
    threads = [threading.Thread(target=work, args=(t, par)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert sorted(par) == [0, 1, 2, 3]
    for t in range(4):
        for (d0, r0, p0), (d1, r1, p1) in zip(seq[t], par[t]):
            assert torch.equal(d0, d1), f"thread {t}: descriptors differ under concurrency"
            assert torch.equal(r0, r1), f"thread {t}: poses differ under concurrency"
            assert torch.equal(p0, p1), f"thread {t}: loop probabilities differ under concurrency"
    # and repeated calls are deterministic
    for t in range(4):
        assert torch.equal(seq[t][0][0], seq[t][2][0]) and torch.equal(seq[t][0][1], seq[t][2][1])


def test_step_is_cuda_graph_capturable(cfg, checkpoint):
    """No host sync, no allocation and no default-stream work inside the C-ABI calls: one frame's encoder (with its
    forked sampling chain) + registration can be captured into a CUDA graph and replayed on new inputs."""
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    enc.load_state_dict(checkpoint["encoder"], strict=True)
    dec.load_state_dict(checkpoint["decoder"], strict=True)
    enc, dec = enc.to(DEV), dec.to(DEV)
    a = torch.stack([data.kitti_shape_cloud(1, 16384), data.kitti_shape_cloud(2, 16384)]).to(DEV)
    b = torch.stack([data.kitti_shape_cloud(3, 16384), data.kitti_shape_cloud(4, 16384)]).to(DEV)
    static_in = a.clone()

    def step(x):
        desc = enc.descriptors(x, None, coor_scale=cfg.coor_scale)
        res, conf = dec.registration_forward_batch(desc[:1], desc[1:], 0.5)
        return desc, res

    with torch.no_grad():
        want_a = [t.clone() for t in step(a)]
        want_b = [t.clone() for t in step(b)]
        side = torch.cuda.Stream(device=DEV)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up on the capture stream: workspaces / forked streams exist
            step(static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            out = step(static_in)
        for inp, want in ((a, want_a), (b, want_b), (a, want_a)):
            static_in.copy_(inp)
            g.replay()
            torch.cuda.synchronize()
            assert torch.equal(out[0], want[0]) and torch.equal(out[1], want[1])


def test_split_weight_reuse_follows_weight_updates(cfg):
    """dpm_set_weights_epoch: a second call with unchanged weights into the same scratch skips the weight-split launch;
    an in-place weight update (version counter) or a new scratch buffer brings it back, and the results follow."""
    import copy
    from oracle import model_ref as M
    from deeppointmap_b200 import Encoder, _C, data
    sd = M.random_weights(M.encoder_shapes(cfg), seed=21)
    enc = Encoder(cfg).eval()
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(DEV)
    pts = data.kitti_shape_cloud(31, 8192)[None].to(DEV)

    def run():
        _C.launch_count_reset()
        with torch.no_grad():
            d = enc.descriptors(pts, None, 60.0)
        torch.cuda.synchronize()
        return d, _C.launch_count()

    d1, n1 = run()
    d2, n2 = run()
    assert torch.equal(d1, d2)
    if _C.REUSE_SPLIT:
        assert n2 == n1 - 1                                   # the split launch is gone
    with torch.no_grad():
        enc.downsampler[1].sa.mlp[0].weight.mul_(1.5)         # in place: same pointer, new version
    d3, n3 = run()
    assert n3 == n1 and not torch.equal(d3, d1)
    want = copy.deepcopy(sd)
    want["downsampler.1.sa.mlp.0.weight"] = want["downsampler.1.sa.mlp.0.weight"] * 1.5
    ref = M.descriptors(want, cfg, pts.cpu(), torch.zeros(1, 8192, dtype=torch.bool), "direct")
    assert float((d3.cpu() - ref).abs().max() / ref.abs().max()) < 1e-4
    _C.workspaces.release()                                   # new scratch buffer: split again
    d4, n4 = run()
    assert n4 == n1 and torch.equal(d4, d3)
