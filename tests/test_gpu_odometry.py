"""Frame-to-frame odometry over a short synthetic corridor sequence (BASELINE config 5 in miniature):
batched encoder + batched registration with descriptors resident on the device, against ground truth and
against the one-pair-at-a-time module API."""
import os

import pytest
import torch

from conftest import ROOT
from deeppointmap_b200 import Decoder, Encoder, sequence

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_sequence_generator_is_deterministic_and_in_range():
    gt = sequence.trajectory(5)
    world = sequence.corridor_world(3, length_m=10.0, ground_density=2.0)
    a = sequence.corridor_frames(world, gt, 4096, seed=7)
    b = sequence.corridor_frames(world, gt, 4096, seed=7)
    assert torch.equal(a, b) and a.shape == (5, 3, 4096)
    d = (a * 60.0).norm(dim=1)
    assert float(d.max()) <= 60.5 and float(d.min()) >= 0.9


def test_batched_odometry_matches_pairwise_api_and_ground_truth(cfg, checkpoint):
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    enc.load_state_dict(checkpoint["encoder"], strict=True)
    dec.load_state_dict(checkpoint["decoder"], strict=True)
    enc, dec = enc.to(DEV), dec.to(DEV)
    n = 9
    gt = sequence.trajectory(n)
    world = sequence.corridor_world(1, length_m=float(gt[-1, 0, 3]), device=DEV)
    frames = sequence.corridor_frames(world, gt, 16384, seed=2)
    rel, est = sequence.run_odometry(enc, dec, frames, batch=4, coor_scale=cfg.coor_scale)  # 3 batches, ragged tail
    assert rel.shape == (n - 1, 16) and est.shape == (n, 4, 4)
    # the same pairs one at a time through the reference-shaped API: identical arithmetic
    desc = enc.descriptors(frames, None, coor_scale=cfg.coor_scale)
    for i in (0, 3, 4, 7):
        R, T, conf, rmse = dec.registration_forward(desc[i], desc[i + 1], num_sample=0.5)
        assert torch.allclose(R.flatten().cpu(), rel[i, 0:9].cpu(), atol=1e-5)
        assert torch.allclose(T.flatten().cpu(), rel[i, 9:12].cpu(), atol=1e-4)
    te, re = sequence.relative_errors(rel, gt)
    # sanity of the chain, not of the network: 1 m steps through a world of structureless random points
    assert float(te.median()) < 0.5 and float(re.median()) < 1.0, (te, re)
