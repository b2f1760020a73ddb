"""Host-side logic that needs no GPU: module API, state_dict compatibility, helpers."""
import copy

import pytest
import torch

from oracle import model_ref as M


def test_state_dict_matches_reference_checkpoint(cfg, checkpoint):
    from deeppointmap_b200 import Encoder, Decoder
    e, d = Encoder(cfg), Decoder(cfg)
    assert list(e.state_dict().keys()) == list(checkpoint["encoder"].keys())
    assert list(d.state_dict().keys()) == list(checkpoint["decoder"].keys())
    e.load_state_dict(checkpoint["encoder"], strict=True)
    d.load_state_dict(checkpoint["decoder"], strict=True)


def test_state_dict_shapes_without_checkpoint(cfg):
    from deeppointmap_b200 import Encoder, Decoder
    e, d = Encoder(cfg), Decoder(cfg)
    assert [(k, tuple(v.shape)) for k, v in e.state_dict().items()] == [(k, tuple(s)) for k, s in M.encoder_shapes(cfg)]
    assert [(k, tuple(v.shape)) for k, v in d.state_dict().items()] == [(k, tuple(s)) for k, s in M.decoder_shapes(cfg)]
    assert len(e.state_dict()) == 110 and len(d.state_dict()) == 82
    assert sum(v.numel() for v in e.state_dict().values()) == 3_824_224  # the shipped checkpoint's encoder (110 tensors)


def test_modules_deepcopy_and_eval(cfg):
    from deeppointmap_b200 import Encoder, Decoder
    e, d = Encoder(cfg).eval(), Decoder(cfg).eval()
    e2, d2 = copy.deepcopy(e), copy.deepcopy(d)  # pipeline/infer_multiagents.py:100-113
    assert e2._desc.npoint[0] == 4096 and d2._desc.attention_layers == 3


def test_decoder_forward_is_training_only(cfg):
    from deeppointmap_b200 import Decoder
    d = Decoder(cfg).eval()
    with pytest.raises(AssertionError):  # decoder.py:37
        d(torch.zeros(1, 131, 4), torch.zeros(1, 131, 4))


def test_num_pairs_follows_reference_rule():
    from deeppointmap_b200 import Decoder
    assert Decoder.num_pairs(0.5, 256, 256) == 128
    assert Decoder.num_pairs(0.5, 4096, 256) == 1088
    assert Decoder.num_pairs(300, 256, 256) == 150
    assert Decoder.num_pairs(300.0, 256, 256) == 150
    for bad in (0.0, -1.0, "x"):
        with pytest.raises(ValueError):
            Decoder.num_pairs(bad, 256, 256)
    for ns in (0.5, 1.0, 7, 300.0):
        assert Decoder.num_pairs(ns, 100, 50) == M.num_pairs(ns, 100, 50)


def test_cpu_tensors_fail_loudly(cfg):
    """No CPU fallback: the product path refuses CPU tensors instead of silently computing."""
    from deeppointmap_b200 import Encoder, ops
    e = Encoder(cfg).eval()
    with pytest.raises(RuntimeError):
        e(torch.zeros(1, 3, 5000), torch.zeros(1, 5000, dtype=torch.bool))
    with pytest.raises(RuntimeError):
        ops.sample_farthest_points(torch.zeros(1, 100, 3), K=4)
    with pytest.raises(RuntimeError):
        ops.knn_points(torch.zeros(1, 4, 3), torch.zeros(1, 100, 3), K=4)


def test_product_never_imports_oracle():
    import os
    import re
    from conftest import ROOT
    pkg = os.path.join(ROOT, "deeppointmap_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "dpm_oracle" not in txt, f


def test_synthetic_cloud_is_deterministic():
    from deeppointmap_b200 import data
    a, b = data.kitti_shape_cloud(5, 4096), data.kitti_shape_cloud(5, 4096)
    assert a.shape == (3, 4096) and a.dtype == torch.float32 and torch.equal(a, b)
    assert not torch.equal(a, data.kitti_shape_cloud(6, 4096))
    r = (a * 60).norm(dim=0)
    assert r.min() >= 1.0 - 1e-3 and r.max() <= 60.0 + 1e-3
    c = data.kitti_shape_cloud(0, 65536)
    assert c.shape == (3, 65536)


def test_encoder_refuses_training_with_grad(cfg):
    """the drop-in has no autograd path; pipeline/train.py must get an error, not detached features."""
    from deeppointmap_b200 import Encoder
    e = Encoder(cfg).train()
    with pytest.raises(NotImplementedError):
        e(torch.zeros(1, 3, 64), torch.zeros(1, 64, dtype=torch.bool))
    with torch.no_grad(), pytest.raises(RuntimeError):  # past the mode check: CPU tensors are refused (no fallback)
        e(torch.zeros(1, 3, 64), torch.zeros(1, 64, dtype=torch.bool))


def test_workspace_cache_is_bounded_and_releasable():
    from deeppointmap_b200 import _C
    ws = _C._Workspaces()
    ws.MAX_SLOTS = 3
    dev = torch.device("cpu")
    for i in range(5):
        ws.buf[(0, f"s{i}")] = torch.empty(10, dtype=torch.uint8)
        while len(ws.buf) > ws.MAX_SLOTS:
            ws.buf.pop(next(iter(ws.buf)))
    assert list(k[1] for k in ws.buf) == ["s2", "s3", "s4"] and ws.held_bytes() == 30
    assert ws.release() == 30 and ws.held_bytes() == 0
