#!/usr/bin/env python
"""1 000-frame synthetic odometry run (BASELINE.json config 5): world + trajectory -> 65 536-point scans ->
batched encoder + frame-to-frame registration on one B200 -> trajectory error against ground truth.

    python examples/odometry_sequence.py [--frames 1000] [--points 65536] [--batch 32]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deeppointmap_b200 import Decoder, Encoder, sequence  # noqa: E402
from deeppointmap_b200.config import dpm_b_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1000)
ap.add_argument("--points", type=int, default=65536)
ap.add_argument("--batch", type=int, default=32)
args = ap.parse_args()

dev = torch.device("cuda", 0)
cfg = dpm_b_config()
enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
ck = os.path.join(ROOT, "oracle", "_ref", "DeepPointMapAAAI.pth")
weights = "random-init"
if os.path.exists(ck):
    sd = torch.load(ck, map_location="cpu")
    enc.load_state_dict(sd["encoder"], strict=True)
    dec.load_state_dict(sd["decoder"], strict=True)
    weights = "DeepPointMapAAAI.pth"
enc, dec = enc.to(dev), dec.to(dev)

t0 = time.perf_counter()
gt = sequence.trajectory(args.frames)
world = sequence.corridor_world(0, length_m=float(gt[:, 0, 3].max()), half_width_m=70.0 + float(gt[:, 1, 3].abs().max()),
                                ground_density=8.0, device=dev)
frames = sequence.corridor_frames(world, gt, args.points, seed=1)
torch.cuda.synchronize()
t_gen = time.perf_counter() - t0

sequence.run_odometry(enc, dec, frames[:min(args.frames, 2 * args.batch)], args.batch, cfg.coor_scale)  # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
rel, est = sequence.run_odometry(enc, dec, frames, args.batch, cfg.coor_scale)
torch.cuda.synchronize()
dt = time.perf_counter() - t0

te, re = sequence.relative_errors(rel, gt)
gt0 = torch.linalg.inv(gt[0]) @ gt  # express ground truth in the first sensor frame, like the estimate
drift = float((est[-1, :3, 3] - gt0[-1, :3, 3]).norm())
print(json.dumps({
    "frames": args.frames, "points_per_frame": args.points, "batch": args.batch, "weights": weights,
    "frames_per_s": args.frames / dt, "seconds": dt, "scan_generation_s": t_gen,
    "rel_translation_err_m": {"mean": float(te.mean()), "median": float(te.median()), "max": float(te.max())},
    "rel_rotation_err_deg": {"mean": float(re.mean()), "median": float(re.median()), "max": float(re.max())},
    "path_length_m": float(args.frames - 1), "end_point_drift_m": drift,
    "mean_rmse": float(rel[:, 12].mean()), "mean_inliers": float(rel[:, 14].mean()),
}))
