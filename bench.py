#!/usr/bin/env python
"""bench.py -- encoder+matcher frames/s at 65 536 points per frame (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One STEP = one pass of the hot path over one batch of F synthetic frames on each GPU:
  Encoder (FPS, kNN+radius grouping, fused group MLP, FPN) -> F unified descriptor sets,
  then Decoder.registration_forward for the F consecutive pairs (frame i-1 -> frame i; the
  first frame of a batch pairs with the last frame of the previous one).
Frames are independent units: rank r works on its own F frames (weak scaling, no data-path
collective; the poses are all-gathered once per step, 64 B per frame).

The single JSON line printed by rank 0 follows the driver contract; see DESIGN.md section 6.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "encoder+matcher frames/sec @65536 pts"
UNIT = "frames/s"
L2_BYTES = 126 * 1024 * 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200, help="timed steps (default 200 = ~1 s of device time)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--frames", type=int, default=32, help="frames per GPU per step (reference EXTRACTOR_BATCHSIZE=32)")
    ap.add_argument("--points", type=int, default=65536)
    ap.add_argument("--streams", type=int, default=0,
                    help="CUDA streams the steps are issued on round-robin (independent frame sequences, like the "
                         "reference's multi-agent mode): the latency-bound FPS chain of one step overlaps the "
                         "throughput-bound kernels of another.  0 = auto: with the packed FPS mapping 8 (10 when only 10 "
                         "divides --steps), else 6, 5 or 4, whichever divides --steps (no partially filled last round "
                         "inside the timed region)")
    ap.add_argument("--fps-mode", type=int, default=3, choices=[0, 1, 2, 3],
                    help="dpm_set_fps_mode for the multi-stream legs (headline, sustained, e2e, kernel profile): 3 = packed, two "
                         "clouds per SM (less SM time per batch, longer FPS latency: pays with >= 8 streams in flight); 0 = the "
                         "library's automatic choice.  The one-stream legs (batch1, strong, caller sizes) always run on 0")
    ap.add_argument("--pack-min-steps", type=int, default=64,
                    help="steps from which a multi-stream leg takes the packed FPS mapping (profiling runs pass 1)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-batch1", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the extra legs (sustained run, strong scaling, caller sizes, uniform cube, reference on the GPU)")
    ap.add_argument("--pipeline-frames", type=int, default=200, help="synthetic scans of the pipeline/infer.py leg (0 = skip)")
    ap.add_argument("--sustain-s", type=float, default=1.5, help="minimum device time of the sustained leg")
    ap.add_argument("--strong-frames", type=int, default=32, help="GLOBAL frames of the strong-scaling leg (BASELINE configs[3])")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the CPU sample (0 = auto)")
    ap.add_argument("--kernels", type=int, default=16, help="entries of the per-kernel breakdown to print")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tensor_peak():
    """dense bf16 TFLOP/s: the sustained figure (the GEMMs are timed inside a long step)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d.get("bf16_tflops_sustained") or d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
        except Exception:
            pass
    return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


def load_weights(cfg):
    """The shipped checkpoint when its copy travelled (oracle/_ref/, made by build()); else the
    modules' seeded default initialisation of the same architecture."""
    ck = os.path.join(ROOT, "oracle", "_ref", "DeepPointMapAAAI.pth")
    if os.path.exists(ck):
        sd = torch.load(ck, map_location="cpu")
        return sd["encoder"], sd["decoder"], "DeepPointMapAAAI.pth"
    return None, None, "random-init"


# ---------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._thr = None
        try:
            if index < 0:
                raise RuntimeError("sampling disabled on this rank")
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                try:
                    phys = int(visible.split(",")[index])
                except Exception:
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.008)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------
# synthetic frames: a short trajectory per batch (SURVEY.md section 8d generator)
# ---------------------------------------------------------------------------------------------
def make_batch(seed: int, frames: int, n: int) -> torch.Tensor:
    from deeppointmap_b200 import data
    base = data.kitti_shape_cloud(seed, n)
    out = [base]
    for i in range(1, frames):
        moved, _, _ = data.rigid_move(base, yaw_deg=0.5 * i, t_m=(1.0 * i, 0.05 * i, 0.0), jitter_m=0.01,
                                      seed=seed * 1000 + i)
        g = torch.Generator().manual_seed(seed * 1000 + i)
        out.append(moved[:, torch.randperm(n, generator=g)])
    return torch.stack(out).contiguous()  # (F, 3, n)


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_port_step(esd, dsd, cfg, pts, prev_desc):
    """One bounded step of the CPU path: encode `pts` (f,3,n) and register the f consecutive
    pairs.  Returns (descriptors of the last frame, poses)."""
    from oracle import model_ref as M
    pad = torch.zeros(pts.shape[0], pts.shape[2], dtype=torch.bool)
    desc = M.descriptors(esd, cfg, pts, pad, "direct")
    poses = []
    chain = [prev_desc] + list(desc) if prev_desc is not None else [desc[-1]] + list(desc)
    for i in range(1, len(chain)):
        R, T, conf, rmse = M.registration_forward(dsd, cfg, chain[i - 1], chain[i], 0.5)
        poses.append((R, T, rmse))
    return desc[-1], poses


def cpu_weights(cfg):
    from oracle import model_ref as M
    esd, dsd, name = load_weights(cfg)
    if esd is None:
        esd = M.random_weights(M.encoder_shapes(cfg), seed=1)
        dsd = M.random_weights(M.decoder_shapes(cfg), seed=2)
    return esd, dsd, name


def run_cpu_sample(cfg, n, frames, steps, warmup):
    from oracle import index_ops
    esd, dsd, wname = cpu_weights(cfg)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    index_ops.set_num_threads(cores)
    batches = [make_batch(100 + s, frames, n) for s in range(min(2, max(1, steps)))]
    prev = None
    for w in range(warmup):
        prev, _ = cpu_port_step(esd, dsd, cfg, batches[w % len(batches)], prev)
    t0 = time.perf_counter()
    for s in range(steps):
        prev, _ = cpu_port_step(esd, dsd, cfg, batches[s % len(batches)], prev)
    dt = time.perf_counter() - t0
    return {"value": frames * steps / dt, "seconds": dt, "cores": cores, "frames_per_step": frames, "weights": wname}


def run_reference_cpu(n, frames, steps, warmup):
    """the UNMODIFIED reference (oracle/_ref/reference, vendored by build()) on the host cores: its own Encoder /
    Decoder with its pure-torch sampler / querier (no pytorch3d in this image), 1 frame + 1 registration at a time as
    pipeline/infer.py does"""
    from oracle import ref_loader
    enc, dec, cfg = ref_loader.load_models("cpu", ops="fallback")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = make_batch(101, max(2, frames + 1), n)
    pad = torch.zeros(1, n, dtype=torch.bool)
    scale = float(cfg.slam_system.coor_scale)

    def one(i, prev):
        with torch.no_grad():
            c, f, _ = enc(batch[i % batch.shape[0]][None], pad)
            desc = torch.cat([f, c * scale], dim=1)[0]          # system/modules/odometry.py:46-49
            if prev is not None:
                dec.registration_forward(prev, desc, num_sample=0.5)
        return desc

    prev = None
    for w in range(warmup * frames):
        prev = one(w, prev)
    if prev is None:
        prev = one(0, None)
    t0 = time.perf_counter()
    for s_ in range(steps * frames):
        prev = one(s_ + 1, prev)
    dt = time.perf_counter() - t0
    return {"value": frames * steps / dt, "seconds": dt, "cores": cores, "frames_per_step": frames,
            "weights": "DeepPointMapAAAI.pth"}


def reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path, timed on the host cores.  When the
    reference's sources travelled with the snapshot (oracle/_ref/reference, made by build()) this is the unmodified
    reference itself (`kind: reference`); otherwise the oracle PORT (oracle/model_ref.py + oracle/dpm_oracle.c: same
    algorithm, C/OpenMP index ops instead of the Python loops, i.e. a FASTER CPU path).  Under torchrun only rank 0
    works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from deeppointmap_b200.config import dpm_b_config
    from oracle import ref_loader
    cfg = dpm_b_config()
    cores = os.cpu_count() or 1
    kind = "reference" if ref_loader.ref_root() is not None and not os.environ.get("DPM_BENCH_PORT") else "port"
    if kind == "reference":
        frames = args.cpu_frames or 1
        try:
            r = run_reference_cpu(args.points, frames, args.steps, min(args.warmup, 1))
            what = "UNMODIFIED reference (its own pure-torch FPS / kNN fallbacks), one frame at a time"
        except Exception as e:  # noqa: BLE001 -- fall back to the port, say so
            print(f"[bench] reference run failed ({e!r}); using the oracle port", file=sys.stderr)
            kind = "port"
    if kind == "port":
        frames = args.cpu_frames or max(1, min(8, cores // 4))
        r = run_cpu_sample(cfg, args.points, frames, args.steps, args.warmup)
        what = "oracle port (torch fp32 + C/OpenMP index ops)"
    sample = (f"{frames} frames x {args.points} pts per step (encoder + {frames} registrations), {what}, "
              f"{r['cores']} threads; {args.steps} steps in {r['seconds']:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": f"synthetic; weights {r['weights']}",
        "config": {"workload": f"synthetic KITTI-shape {args.points}-pt frames, DeepPointMap_B encoder fwd + pairwise "
                               f"registration (descriptor match + SVD pose), CPU bounded sample of {frames} frames/step"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": kind, "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ---------------------------------------------------------------------------------------------
# extra legs (VERDICT r1 "next round" item 1): each is device-timed with CUDA events on the stream it launches on
# ---------------------------------------------------------------------------------------------
def _timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _profile_call(fn):
    """per-kernel device ms of one call (CUDA event after every launch on the launching stream)"""
    from deeppointmap_b200 import _C
    torch.cuda.synchronize()
    _C.prof_begin()
    fn()
    tot = {}
    for tag, a, b, ms in _C.prof_end():
        if tag != "host_gap":
            tot[tag] = tot.get(tag, 0.0) + ms
    return dict(sorted(((k, round(v, 4)) for k, v in tot.items()), key=lambda kv: -kv[1]))


def caller_size_legs(dec, descs, tpeak):
    """The other shapes the SLAM callers push through the decoder (never the headline): scan-to-map M=4096 x N=256
    (mapping.py:153 with graph_search max_k=16, pose_graph.py:513), map-to-map 4096 x 4096 (loop_closure.py:240) and
    the batched loop head, C=8 candidates (loop_closure.py:171)."""
    out = {}
    S = descs.shape[2]
    nset = descs.shape[0]
    tile = lambda first, m: torch.cat([descs[(first + i) % nset] for i in range(m)], dim=1).unsqueeze(0).contiguous()
    C = dec.model_channel
    for name, mb, nb in (("scan_to_map_4096x256", 16, 1), ("map_to_map_4096x4096", 16, 16)):
        src, dst = tile(0, mb), tile(5, nb)
        Mm, Nn = src.shape[2], dst.shape[2]
        fn = lambda: dec.registration_forward_batch(src, dst, 0.5)
        ms = _timed(fn, 5)
        kern = _profile_call(fn)
        att_flops = dec.attention_layers * 4.0 * C * float(Mm + Nn) ** 2     # QK^T + PV, self + cross, both sides
        att_ms = kern.get("attention", 0.0) + kern.get("attention_tc5", 0.0)
        out[name] = {"ms_per_call": ms, "pairs_k": dec.num_pairs(0.5, Mm, Nn), "kernels_ms": kern,
                     "attention_TFLOPs": att_flops / (att_ms * 1e-3) / 1e12 if att_ms > 0 else None,
                     "attention_frac_of_bf16_peak": att_flops / (att_ms * 1e-3) / 1e12 / tpeak if att_ms > 0 else None}
    src = torch.stack([descs[i % nset] for i in range(8)]).contiguous()
    dst = descs[nset - 1].unsqueeze(0).repeat(8, 1, 1).contiguous()
    out["loop_detection_C8_256x256"] = {"ms_per_call": _timed(lambda: dec.loop_detection_forward(src, dst), 10),
                                        "kernels_ms": _profile_call(lambda: dec.loop_detection_forward(src, dst))}
    return out


def index_ops_on(cloud_kind, n, frames, dev, cfg):
    """stage-0 FPS + hybrid query alone on `frames` clouds of one kind (kitti shape / the adversarial uniform cube of
    SURVEY 8d), in both FPS mappings"""
    from deeppointmap_b200 import data, ops
    mk = data.kitti_shape_cloud if cloud_kind == "kitti" else data.uniform_cube_cloud
    e = cfg.encoder
    K, r, ns = int(e.npoint[0]), float(e.radius_list[0][0]), int(e.nsample_list[0][0])
    out = {}
    for B in sorted({1, frames}):
        pts = torch.stack([mk(900 + i, n).T.contiguous() for i in range(B)]).to(dev)
        pad = torch.zeros(B, n, dtype=torch.bool, device=dev)
        ctr = ops.sample_farthest_points(pts, K=K)[0]
        rec = {}
        for mode, label in ((1, "one_sm_per_cloud"), (2, "cluster_per_cloud")):
            if mode == 2 and B > 16:
                continue
            ops.set_fps_mode(mode)
            ms = _timed(lambda: ops.sample_farthest_points(pts, K=K), 3, 1)
            rec[f"fps_{label}_ms"] = ms
            rec[f"fps_{label}_us_per_pick"] = 1e3 * ms / (K - 1)
        ops.set_fps_mode(0)
        rec["knn_ms"] = _timed(lambda: ops.hybrid_query(r, ns, pts, ctr, pad), 3, 1)
        out[f"B{B}"] = rec
    return out


def reference_gpu_leg(cfg, dev, n, pts2):
    """north_star's >= 40x denominator: "the reference single-GPU PyTorch encoder+matcher" on THIS B200
    (BASELINE.md section 3).  The unmodified reference modules `.to(cuda)` with their own pure-torch sampler / querier
    when its sources travelled (oracle/_ref/reference); else the oracle restatement of the same fallbacks
    (oracle/model_ref.py: fps_torch + dense-distance topk).  One frame + one registration per step, as
    pipeline/infer.py runs it; CUDA-event timed."""
    from oracle import ref_loader
    pad = torch.zeros(1, n, dtype=torch.bool, device=dev)
    frames = [pts2[i:i + 1].contiguous() for i in range(pts2.shape[0])]
    try:
        if ref_loader.ref_root() is None:
            raise RuntimeError("reference sources not on this box")
        enc, dec, rcfg = ref_loader.load_models(dev, ops="fallback")
        scale = float(rcfg.slam_system.coor_scale)

        def one(i, prev):
            c, f, _ = enc(frames[i % len(frames)], pad)
            d = torch.cat([f, c * scale], dim=1)[0]
            if prev is not None:
                dec.registration_forward(prev, d, num_sample=0.5)
            return d
        kind = "unmodified reference modules on cuda (pure-torch FPS loop + dense-distance topk; no pytorch3d)"
    except Exception as e:  # noqa: BLE001
        from oracle import model_ref as M
        esd, dsd, _ = cpu_weights(cfg)
        esd = {k: v.to(dev) for k, v in esd.items()}
        dsd = {k: v.to(dev) for k, v in dsd.items()}

        def one(i, prev):
            d = M.descriptors(esd, cfg, frames[i % len(frames)], pad, "fallback", fps_mode="torch")[0]
            if prev is not None:
                M.registration_forward(dsd, cfg, prev, d, 0.5)
            return d
        kind = f"oracle restatement of the reference fallbacks on cuda ({e})"
    with torch.no_grad():
        prev = one(0, None)
        prev = one(1, prev)
        torch.cuda.synchronize()
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            prev = one(2 + i, prev)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"frames_per_s": 1e3 / ms, "ms_per_frame": ms, "what": kind, "frames_timed": reps}

def pipeline_leg(n_synth, n_points, dev):
    """BASELINE configs[4]: the reference's own pipeline/infer.py, unmodified, on the drop-in modules -- over
    `n_synth` synthetic 65 536-point .bin scans (a street world: undulating ground, kerbs, facades with recesses, poles,
    trees, parked boxes, seen from a 1 m / frame trajectory) and over the reference's real sample
    scans (51 frames forth and back) when they travelled.  frames/s from the reference's own per-stage timers
    (ResultLogger.log_time) and as whole-process wall clock (interpreter start, checkpoint load, file IO included)."""
    import shutil
    import tempfile
    from oracle import ref_loader
    from deeppointmap_b200 import pipeline
    ref, weight = ref_loader.ref_root(), ref_loader.checkpoint_path()
    if ref is None or weight is None:
        return {"unavailable": "reference sources not on this box (oracle/_ref/reference is made by build())"}
    work = tempfile.mkdtemp(prefix="dpm_pipeline_")
    out = {}
    try:
        jobs = []
        if n_synth > 0:
            seq = os.path.join(work, "synth", "0")
            pipeline.write_synthetic_sequence(seq, n_synth, n_points, seed=5, device=dev)
            jobs.append(("synthetic", seq, n_synth, None))   # shipped thresholds: the mapping thread drops the scans it distrusts
        real = ref_loader.sample_frames()
        if len(real) >= 2:
            seq = os.path.join(work, "real", "0")
            pipeline.write_boomerang_sequence(seq, real, 51)
            jobs.append(("real_kitti_sample", seq, 51, None))
        for name, seq, nf, over in jobs:
            y = pipeline.write_yaml(os.path.join(work, f"{name}.yaml"), ref, [seq], os.path.join(work, f"out_{name}"),
                                    transforms=pipeline.DEFAULT_TRANSFORMS, slam_overrides=over)
            r = pipeline.run_infer_subprocess(ref, "b200", y, weight, timeout=600)
            rec = {"frames": nf, "returncode": r["returncode"], "wall_s": r["wall_s"], "wall_frames_per_s": nf / r["wall_s"],
                   "loop_frames_per_s": r["loop_frames_per_s"], "stage_mean_s": r["stage_mean_s"],
                   "transforms": list(pipeline.DEFAULT_TRANSFORMS)}
            if r["returncode"] != 0:
                rec["stderr_tail"] = r["stderr_tail"][-300:]
            try:
                rec["scans_in_trajectory"] = int(len(pipeline.load_trajectory(os.path.join(work, f"out_{name}"))[0]))
            except Exception:  # noqa: BLE001
                rec["scans_in_trajectory"] = None
            out[name] = rec
    finally:
        shutil.rmtree(work, ignore_errors=True)
    out["what"] = ("python -m deeppointmap_b200.pipeline --impl b200 -> <reference>/pipeline/infer.py unmodified (single-thread "
                   "mode: SlamSystem.step per scan, batch 1, loop closure off); see tests/test_gpu_pipeline.py for the "
                   "trajectory comparison against the reference's own modules")
    return out


# ---------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch.distributed as dist
    from deeppointmap_b200 import Encoder, Decoder, _C
    from deeppointmap_b200.config import dpm_b_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; reporting n_gpus={world}", file=sys.stderr)

    cfg = dpm_b_config()
    F, n, K, W = args.frames, args.points, args.steps, args.warmup
    esd, dsd, wname = load_weights(cfg)
    torch.manual_seed(1234)
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    if esd is not None:
        enc.load_state_dict(esd, strict=True)
        dec.load_state_dict(dsd, strict=True)
    enc, dec = enc.to(dev), dec.to(dev)
    S = enc._out_points
    Cd = enc.final_channel + 3

    # input pool larger than L2 so no timed step finds its input cached
    bytes_per_batch = F * 3 * n * 4
    nslots = max(2, -(-int(1.25 * L2_BYTES) // bytes_per_batch))
    nslots = min(nslots, 64)
    host_pool = [make_batch((rank * 64 + s) * 7 + 1, F, n).pin_memory() for s in range(nslots)]
    dev_pool = [h.to(dev) for h in host_pool]
    pool_mb = nslots * bytes_per_batch / 2 ** 20

    # The packed FPS mapping (two clouds per SM) trades FPS latency for SM time: it needs >= 8 streams in flight AND a
    # pipeline long enough to leave the lock-step start behind (measured: 7480 against 7240 frames/s at 200 steps, but
    # 6910 against 7160 at 20 steps).  So a leg of >= 64 steps runs packed on 8 streams, a shorter one on the library's
    # automatic mapping with 6 / 5 / 4 streams.
    PACK_MIN_STEPS = args.pack_min_steps
    can_pack = args.fps_mode == 3 and F > int(_C.lib().dpm_fps_cluster_capacity())

    def leg_config(nsteps, flexible=False):
        """-> (packed?, streams) of a multi-stream leg of `nsteps` steps (flexible: the caller rounds the steps up to
        whole rounds, so the stream count need not divide them)"""
        if can_pack and nsteps >= PACK_MIN_STEPS:
            return True, (args.streams if args.streams > 0 else (8 if flexible or nsteps % 8 == 0 or nsteps % 10 else 10))
        return False, (args.streams if args.streams > 0 else next((c for c in (6, 5, 4) if nsteps % c == 0), 4))

    packed, NS = leg_config(K)
    NSMAX = max(NS, 10 if can_pack else NS)
    from deeppointmap_b200 import ops as _ops

    def fps_mode_multi(pk=None):   # the legs that keep several streams of batches in flight
        pk = packed if pk is None else pk
        _ops.set_fps_mode(3 if pk else (0 if args.fps_mode == 3 else args.fps_mode))

    def fps_mode_auto():    # one-stream legs: the library's own choice (cluster per cloud for small batches)
        _ops.set_fps_mode(0)
    streams = [torch.cuda.Stream(device=dev) for _ in range(NSMAX)]
    descbufs = [torch.zeros((F + 1, Cd, S), dtype=torch.float32, device=dev) for _ in range(NSMAX)]
    descbuf = descbufs[0]
    gathered = [torch.empty((world * F, _C.REG_STRIDE), dtype=torch.float32, device=dev) if world > 1 else None
                for _ in range(NSMAX)]
    k_pairs = dec.num_pairs(0.5, S, S)
    comm = torch.cuda.Stream(device=dev) if world > 1 else None

    def step(pts, q=0):
        """device-resident step of sequence q (on the current stream): F frames -> F descriptors -> F poses"""
        db = descbufs[q]
        enc.descriptors(pts, None, coor_scale=cfg.coor_scale, out=db[1:])
        result, conf = dec.registration_forward_batch(db[:F], db[1:], 0.5)
        db[0].copy_(db[F])
        if world > 1:
            # the pose all-gather (64 B per frame) runs on its own stream behind this step's event: the compute
            # streams never wait for the other ranks, so one slow rank does not stall the next step of the rest
            done = torch.cuda.Event()
            done.record()
            comm.wait_event(done)
            result.record_stream(comm)
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(gathered[q], result)
        return result, conf

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, first, fn, ns=None):
        """n steps round-robin over the first `ns` streams; returns device ms from a common start event to the
        last stream's end (events on the launching streams)"""
        ns = NS if ns is None else ns
        cur = torch.cuda.current_stream()
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(cur)
        for st in streams[:ns]:
            st.wait_event(ev0)
        for i in range(n):
            q = i % ns
            with torch.cuda.stream(streams[q]):
                fn(first + i, q)
        for st in streams[:ns] + ([comm] if comm is not None else []):  # the gathered poses are part of the job
            e = torch.cuda.Event()
            e.record(st)
            cur.wait_event(e)
        ev1.record(cur)
        return ev0, ev1

    with torch.no_grad():
        fps_mode_multi()
        # every stream is warmed (its workspace allocated) before the timed region; one extra step staggers them
        run_steps(max(W, NS) + (1 if NS > 1 else 0), 0, lambda i, q: step(dev_pool[i % nslots], q))
        # rank 0 reports the clocks (NVML polling stays off the other ranks); nvmlInit happens BEFORE the barrier so that
        # no rank enters the timed region late and makes the others wait in their first collective
        clk_obj = ClockSampler(local if rank == 0 else -1)
        barrier()
        # ---- timed region: K steps, CUDA events on the launching streams ---------------------
        _C.launch_count_reset()
        with clk_obj as clk:
            t_wall = time.perf_counter()
            ev0, ev1 = run_steps(K, W, lambda i, q: step(dev_pool[i % nslots], q))
            barrier()
            t_wall = time.perf_counter() - t_wall
        launches = _C.launch_count()
        total_ms = ev0.elapsed_time(ev1)
        step_ms = [total_ms / K]
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        value = world * F * K / (total_ms * 1e-3)

        # ---- sustained: the same loop for >= --sustain-s of device time, clocks sampled throughout --------------
        sustained = None
        if not args.no_extra:
            reps = max(1, int(-(-args.sustain_s * 1e3 // max(total_ms, 1e-3))))
            ks = reps * K
            spk, sns = leg_config(ks, flexible=True)
            ks = -(-ks // sns) * sns  # whole rounds
            fps_mode_multi(spk)
            if sns > NS or spk != packed:  # streams / kernels the headline did not warm
                run_steps(sns + 1, 0, lambda i, q: step(dev_pool[i % nslots], q), sns)
            clk2 = ClockSampler(local if rank == 0 else -1)
            barrier()
            with clk2 as c2:
                s0, s1 = run_steps(ks, W + K, lambda i, q: step(dev_pool[i % nslots], q), sns)
                barrier()
            sms = s0.elapsed_time(s1)
            ts = torch.tensor([sms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            sms = float(ts.item())
            sustained = {"steps": ks, "seconds": sms * 1e-3, "value": world * F * ks / (sms * 1e-3), "unit": UNIT,
                         "ms_per_step": sms / ks, "clocks": c2.summary(), "streams_per_gpu": sns,
                         "fps_mode": "packed: two clouds per SM (dpm_set_fps_mode(3))" if spk else "automatic",
                         "note": "same loop as the headline, run long enough for clocks / power to settle; from 64 steps "
                                 "on a leg runs the packed FPS mapping on 8 streams (see config.fps_mode)"}

        # ---- strong scaling (BASELINE configs[3] as written): --strong-frames GLOBAL frames sharded over the ranks
        # through FrameParallel.odometry (boundary-descriptor all-gather + pose all-gather over NCCL), one stream ----
        strong = None
        fps_mode_auto()
        if not args.no_extra:
            from deeppointmap_b200.frames import FrameParallel, shard
            G = args.strong_frames
            a0, a1 = shard(G, world, rank)
            gsets = [make_batch(7000 + s_, G, n)[a0:a1].to(dev) for s_ in range(max(2, min(nslots, 6)))]
            fp = FrameParallel(lambda p_: enc.descriptors(p_, None, coor_scale=cfg.coor_scale),
                               lambda s_, d_: dec.registration_forward_batch(s_, d_, 0.5)[0])
            prevd = torch.zeros((Cd, S), dtype=torch.float32, device=dev)
            for i in range(3):
                fp.odometry(gsets[i % len(gsets)], G, prevd, desc_shape=(Cd, S))
            barrier()
            nrep = 20
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for i in range(nrep):
                fp.odometry(gsets[i % len(gsets)], G, prevd, desc_shape=(Cd, S))
            g1.record()
            barrier()
            tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            gms = float(tg.item()) / nrep
            strong = {"scaling": "strong", "global_frames_per_step": G, "frames_on_rank0": a1 - a0, "ms_per_step": gms,
                      "value": G / (gms * 1e-3), "unit": UNIT, "steps": nrep, "streams_per_gpu": 1,
                      "api": "FrameParallel.odometry (frames.py): encode block -> all-gather boundary descriptors -> "
                             "register block -> all-gather poses",
                      "note": "total work fixed: divide by the n_gpus=1 figure of the same leg for the strong-scaling "
                              "speed-up; the FPS chain of a cloud is sequential, so the floor per step is one cloud's "
                              "FPS latency however few frames a rank holds"}
            del gsets

        # ---- end-to-end: pinned host frames in, poses out, through the module API ----------
        e2e = None
        fps_mode_multi()
        if not args.no_e2e:
            # two steps in flight per stream (double-buffered staging / result buffers): the host reads the poses of step
            # i - 2 NS while step i - NS still runs, so a stream never idles waiting for the host to notice its last step
            DEPTH = 2
            stage = [[torch.empty((F, 3, n), dtype=torch.float32, device=dev) for _ in range(DEPTH)] for _ in range(NS)]
            host_out = [[torch.empty((F, _C.REG_STRIDE), dtype=torch.float32).pin_memory() for _ in range(DEPTH)] for _ in range(NS)]
            done = [[None] * DEPTH for _ in range(NS)]
            turn = [0] * NS
            poses_seen = [0]

            h2d = torch.cuda.Stream(device=dev)      # input prefetch: the copy of a stream's NEXT step runs under its current one
            ready = [[None] * DEPTH for _ in range(NS)]  # (step index, event) of a prefetched input
            e2e_end = [0]                                # one past the last step of the current run: nothing is prefetched beyond it

            def e2e_step(i, q):
                j = turn[q]
                turn[q] = (j + 1) % DEPTH
                if done[q][j] is not None:   # the consumer reads step i - DEPTH * NS's poses before its buffers are reused
                    done[q][j].synchronize()
                    poses_seen[0] += int(host_out[q][j].shape[0])
                cur = torch.cuda.current_stream()
                if ready[q][j] is not None and ready[q][j][0] == i:
                    cur.wait_event(ready[q][j][1])
                else:
                    stage[q][j].copy_(host_pool[i % nslots], non_blocking=True)
                ready[q][j] = None
                result, _ = step(stage[q][j], q)
                host_out[q][j].copy_(result, non_blocking=True)
                done[q][j] = torch.cuda.Event()
                done[q][j].record()
                # prefetch the input of this stream's next step into the other staging buffer, once its last user is done
                jn, nxt = turn[q], i + NS
                if nxt < e2e_end[0]:
                    if done[q][jn] is not None:
                        h2d.wait_event(done[q][jn])
                    with torch.cuda.stream(h2d):
                        stage[q][jn].copy_(host_pool[nxt % nslots], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record()
                    ready[q][jn] = (nxt, ev)

            def drain():
                h2d.synchronize()
                for q in range(NS):
                    for j in range(DEPTH):
                        ready[q][j] = None
                        if done[q][j] is not None:
                            done[q][j].synchronize()
                            poses_seen[0] += int(host_out[q][j].shape[0])
                            done[q][j] = None

            e2e_end[0] = max(3, W // 2)
            run_steps(max(3, W // 2), 0, e2e_step)
            drain()
            barrier()
            e2e_end[0] = W + K
            t0 = time.perf_counter()
            run_steps(K, W, e2e_step)
            drain()
            barrier()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e = {"value": world * F * K / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": bytes_per_batch,
                   "d2h_bytes_per_step": F * _C.REG_STRIDE * 4, "ms_per_step": 1e3 * float(tt.item()) / K,
                   "timing": "host wall clock around K steps, all pose records read on the host",
                   "api": "Encoder.descriptors + Decoder.registration_forward_batch (one C-ABI call each), pinned host "
                          "input (prefetched on a copy stream under the stream's previous step), pose records copied back to pinned memory every "
                          "step; two steps in flight per stream"}

        # ---- batch 1 (BASELINE.json configs[1]/[2]): one frame + one registration at a time, one stream ----
        batch1 = None
        fps_mode_auto()
        if not args.no_batch1:
            db1 = torch.zeros((2, Cd, S), dtype=torch.float32, device=dev)
            one = [dev_pool[s % nslots][(3 * s) % F:(3 * s) % F + 1].contiguous() for s in range(min(16, 4 * nslots))]

            def step1(i):
                enc.descriptors(one[i % len(one)], None, coor_scale=cfg.coor_scale, out=db1[1:])
                r1, _ = dec.registration_forward_batch(db1[:1], db1[1:], 0.5)
                db1[0].copy_(db1[1])
                return r1

            for i in range(3):
                step1(i)
            torch.cuda.synchronize()
            n1 = 20
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            for i in range(n1):
                step1(3 + i)
            b1.record()
            torch.cuda.synchronize()
            ms1 = b0.elapsed_time(b1) / n1
            batch1 = {"frames_per_s": 1e3 / ms1, "ms_per_frame": ms1, "streams": 1, "frames_per_step": 1,
                      "note": "latency-bound: 4095 + 1023 + 255 + 63 + 15 sequential FPS picks per frame"}
            # the same frames PIPELINED over two streams: the registration of frame i (decoder, ~0.9 ms of small kernels)
            # runs while frame i+1 is being encoded (its FPS chain occupies one 8-SM cluster).  Per-frame LATENCY is
            # unchanged; this is the sustained rate of a streaming odometry front-end at batch 1.
            try:
                s_enc, s_dec = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
                ring = [torch.zeros((Cd, S), dtype=torch.float32, device=dev) for _ in range(4)]
                evs = [torch.cuda.Event() for _ in range(4)]
                dec_done = [torch.cuda.Event() for _ in range(4)]

                def piped(nfr, first):
                    for j in range(nfr):
                        i = first + j
                        slot = i % 4
                        with torch.cuda.stream(s_enc):
                            s_enc.wait_event(dec_done[slot])          # the decoder that last read this slot is done
                            enc.descriptors(one[i % len(one)], None, coor_scale=cfg.coor_scale, out=ring[slot].unsqueeze(0))
                            evs[slot].record(s_enc)
                        with torch.cuda.stream(s_dec):
                            s_dec.wait_event(evs[slot])
                            dec.registration_forward_batch(ring[(i - 1) % 4].unsqueeze(0), ring[slot].unsqueeze(0), 0.5)
                            dec_done[(i - 1) % 4].record(s_dec)

                for e_ in dec_done:
                    e_.record(s_dec)
                piped(4, 0)
                torch.cuda.synchronize()
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                s_enc.wait_event(p0); s_dec.wait_event(p0)
                piped(n1, 4)
                fin = torch.cuda.Event(); fin.record(s_dec)
                torch.cuda.current_stream().wait_event(fin)
                fin2 = torch.cuda.Event(); fin2.record(s_enc)
                torch.cuda.current_stream().wait_event(fin2)
                p1.record()
                torch.cuda.synchronize()
                batch1["pipelined_ms_per_frame"] = p0.elapsed_time(p1) / n1
                batch1["pipelined_note"] = ("two streams: registration of frame i overlaps the encoder of frame i+1; "
                                            "latency per frame stays ms_per_frame")
            except Exception as e:  # noqa: BLE001
                batch1["pipelined_error"] = str(e)[:200]
            # the same step captured once into a CUDA graph and replayed (the calls never sync or allocate)
            try:
                gin = one[0].clone()
                cs = torch.cuda.Stream(device=dev)
                cs.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(cs):
                    enc.descriptors(gin, None, coor_scale=cfg.coor_scale, out=db1[1:])
                    dec.registration_forward_batch(db1[:1], db1[1:], 0.5)
                torch.cuda.current_stream().wait_stream(cs)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=cs):
                    enc.descriptors(gin, None, coor_scale=cfg.coor_scale, out=db1[1:])
                    gres, _ = dec.registration_forward_batch(db1[:1], db1[1:], 0.5)
                    db1[0].copy_(db1[1])
                for i in range(3):
                    gin.copy_(one[i % len(one)])
                    graph.replay()
                torch.cuda.synchronize()
                b0.record()
                for i in range(n1):
                    gin.copy_(one[(3 + i) % len(one)])
                    graph.replay()
                b1.record()
                torch.cuda.synchronize()
                batch1["cuda_graph_ms_per_frame"] = b0.elapsed_time(b1) / n1
            except Exception as e:  # noqa: BLE001 -- report, the eager figure stands
                batch1["cuda_graph_ms_per_frame"] = None
                batch1["cuda_graph_error"] = str(e)[:200]

        # ---- the other caller shapes, the adversarial cloud, the reference on this GPU (rank 0 of a 1-GPU run) ----
        callers = cube = ref_gpu = pipe = None
        if not args.no_extra and world == 1:
            tpk, _ = tensor_peak()
            try:
                callers = caller_size_legs(dec, descbufs[0], tpk)
            except Exception as e:  # noqa: BLE001 -- an extra leg never takes the headline down
                callers = {"error": repr(e)[:300]}
            try:
                cube = {"kitti_shape": index_ops_on("kitti", n, F, dev, cfg), "uniform_cube": index_ops_on("cube", n, F, dev, cfg)}
            except Exception as e:  # noqa: BLE001
                cube = {"error": repr(e)[:300]}
            try:
                ref_gpu = reference_gpu_leg(cfg, dev, n, dev_pool[0][:5])
            except Exception as e:  # noqa: BLE001
                ref_gpu = {"error": repr(e)[:300]}
            try:
                pipe = pipeline_leg(args.pipeline_frames, n, dev)
            except Exception as e:  # noqa: BLE001
                pipe = {"error": repr(e)[:300]}

        # ---- per-kernel profile passes (a CUDA event after every launch, on the launching stream) ---------
        # "isolated": the step alone on one stream.  "concurrent": the profiled step on stream 0 while the other
        # NS-1 streams run unprofiled steps, i.e. the conditions of the timed region (a kernel's time then includes
        # what it loses to / hides behind the other streams' kernels).
        prof_steps = 3

        def profile_pass(concurrent):
            agg = {}
            for s_ in range(prof_steps):
                torch.cuda.synchronize()
                if concurrent:
                    for q in range(1, NS):
                        with torch.cuda.stream(streams[q]):
                            for j in range(3):
                                step(dev_pool[(s_ + q + j) % nslots], q)
                with torch.cuda.stream(streams[0]):
                    _C.prof_begin()
                    enc.descriptors(dev_pool[s_ % nslots], None, coor_scale=cfg.coor_scale, out=descbuf[1:])
                    dec.registration_forward_batch(descbuf[:F], descbuf[1:], 0.5)
                    recs = _C.prof_end()
                for tag, a, b, ms in recs:
                    if tag == "host_gap":  # stream idle between the two C-ABI calls of a step: not kernel time
                        continue
                    e = agg.setdefault((tag, a, b), [0.0, 0])
                    e[0] += ms
                    e[1] += 1
            torch.cuda.synchronize()
            k_ = sorted(((ms / prof_steps, cnt // prof_steps, tag, a, b) for (tag, a, b), (ms, cnt) in agg.items()),
                        reverse=True)
            tot = {}
            for ms, cnt, tag, a, b in k_:
                tot[tag] = round(tot.get(tag, 0.0) + ms, 4)
            return k_, dict(sorted(tot.items(), key=lambda kv: -kv[1]))

        fps_mode_multi()  # the kernels of the timed region
        kern, kern_totals = profile_pass(False)
        prof_total = sum(k[0] for k in kern)
        kern_conc, kern_totals_conc = profile_pass(True) if (NS > 1 and not args.no_extra) else (None, None)
        fps_mode_auto()

    # ---- roofline of the dominant kernel --------------------------------------------------
    peak, peak_src = peaks()
    top = kern[0]
    top_ms, top_cnt, top_tag, ta, tb = top
    per_launch_ms = top_ms / max(1, top_cnt)
    if top_tag == "fps":
        alg = 20.0 * ta * (tb - 1) * F  # 20 B per point per iteration (xyz 12 + min-dist r/w 8), SURVEY 8d
        what = f"fps_kernel N={ta} K={tb} x {F} clouds"
    elif top_tag == "knn":
        k0 = cfg.encoder.nsample_list[0][0]
        alg = (12.0 * ta * tb + 12.0 * ta * k0) * F  # every query streams every candidate xyz, SURVEY 8d
        what = f"knn_kernel S={ta} N={tb} x {F} clouds"
    else:
        alg = None
        what = f"{top_tag} ({ta},{tb})"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("fps_packed" if (top_tag == "fps" and packed) else top_tag, tj.get(top_tag))
        except Exception:
            traffic = None
    sm_total = torch.cuda.get_device_properties(dev).multi_processor_count
    fps_cluster = F <= int(_C.lib().dpm_fps_cluster_capacity()) and args.fps_mode in (0, 2)
    sms_used = (min(sm_total, 8 * F) if fps_cluster else min(sm_total, (F + 1) // 2 if packed else F)) if top_tag == "fps" else None
    conc_launch_ms = None
    if kern_conc:
        for ms_, cnt_, tag_, a_, b_ in kern_conc:
            if (tag_, a_, b_) == (top_tag, ta, tb):
                conc_launch_ms = ms_ / max(1, cnt_)
    roofline = {"bound": "hbm", "kernel": what, "achieved": (alg / (per_launch_ms * 1e-3) / 1e9) if alg else None,
                "peak": peak, "unit": "GB/s", "frac": (alg / (per_launch_ms * 1e-3) / 1e9 / peak) if alg else None,
                "traffic": traffic, "peak_source": peak_src, "launch_ms": per_launch_ms,
                "algorithmic_bytes_per_launch": alg, "share_of_step": top_ms / prof_total if prof_total else None,
                "dram_frac": (traffic / (per_launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "us_per_pick": (1e3 * per_launch_ms / max(1, tb - 1)) if top_tag == "fps" else None,
                "sms_used": sms_used, "sms_total": sm_total,
                "fps_mapping": ("cluster of 8 CTAs per cloud" if fps_cluster else
                                ("two clouds per CTA / SM (packed, dpm_set_fps_mode(3)): us_per_pick is per PAIR of clouds"
                                 if packed else "one CTA per cloud")) if top_tag == "fps" else None,
                "launch_ms_concurrent": conc_launch_ms,
                "note": "`frac` follows SURVEY 8d's streaming model (bytes a brute-force FPS would move) and is NOT an HBM "
                        "utilisation: the exact bucket-pruned FPS touches ~1 % of those bytes and they stay in L2 / shared "
                        "memory.  The hardware figures are `dram_frac` (real DRAM traffic / time / peak), `us_per_pick` (the "
                        "latency of the 4095-pick dependent chain that bounds the kernel) and `sms_used`"}
    # the tensor-core side of the path: every linear layer is a tcgen05 3xTF32 GEMM (algorithmic flops 2 M N K)
    tpeak, tpeak_src = tensor_peak()
    gemm = [(ms, cnt, a, b) for ms, cnt, tag, a, b in kern if tag in ("linear_tc", "linear_ln_tc") and a and b]
    roofline_tensor = None
    if gemm:
        g_ms = sum(x[0] for x in gemm)
        g_flops = sum(2.0 * x[2] * x[3] * x[1] for x in gemm)
        big = max(gemm, key=lambda x: x[0])
        roofline_tensor = {"bound": "tensor", "kernel": f"linear_tc_kernel, all {sum(x[1] for x in gemm)} launches of a step",
                           "achieved": g_flops / (g_ms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                           "frac": g_flops / (g_ms * 1e-3) / 1e12 / tpeak, "peak_source": tpeak_src, "ms_per_step": g_ms,
                           "algorithmic_flops_per_step": g_flops,
                           "largest": {"rows": big[2], "n_times_k": big[3], "launches": big[1],
                                       "TFLOPs": 2.0 * big[2] * big[3] * big[1] / (big[0] * 1e-3) / 1e12},
                           "note": "fp32-parity GEMMs run as 3 tf32 MMAs per product (3xTF32) against a dense-bf16 peak: "
                                   "1/6 of that peak is the ceiling; the launches are latency-bound single tiles (DESIGN 6c)"}
    # the second index kernel, for the FPS + ball-query figure of the headline metric
    index_kernels = {}
    for ms, cnt, tag, a, b in kern:
        if tag == "fps" and a == n:
            index_kernels["fps_stage0"] = {"ms": ms / max(1, cnt), "GBps": 20.0 * a * (b - 1) * F / (ms / max(1, cnt) * 1e-3) / 1e9}
        if tag == "knn" and b == n:
            k0 = cfg.encoder.nsample_list[0][0]
            index_kernels["knn_stage0"] = {"ms": ms / max(1, cnt),
                                           "GBps": (12.0 * a * b + 12.0 * a * k0) * F / (ms / max(1, cnt) * 1e-3) / 1e9}
    fps_ms = sum(ms for ms, cnt, tag, a, b in kern if tag == "fps")
    knn_ms = sum(ms for ms, cnt, tag, a, b in kern if tag == "knn")
    idx_bytes = 0.0
    e = cfg.encoder
    lv = [n] + list(e.npoint)
    for i, s_ in enumerate(e.npoint):
        idx_bytes += 20.0 * lv[i] * (s_ - 1)
        idx_bytes += 12.0 * s_ * lv[i] + 12.0 * s_ * e.nsample_list[i][0]
        seen = set()
        for j in range(1, len(e.radius_list[i])):
            key = (e.radius_list[i][j], e.nsample_list[i][j])
            if key in seen:
                continue
            seen.add(key)
            idx_bytes += 12.0 * s_ * s_ + 12.0 * s_ * e.nsample_list[i][j]
    idx_gbps = idx_bytes * F / ((fps_ms + knn_ms) * 1e-3) / 1e9 if fps_ms + knn_ms > 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": f"synthetic; weights {wname}",
        "config": {"workload": f"synthetic KITTI-shape {n}-pt clouds, DeepPointMap_B encoder fwd + pairwise registration "
                               f"(descriptor match + SVD pose, 256x256 descriptors, k={k_pairs}); {F} frames per GPU per step",
                   "frames_per_gpu_per_step": F, "points_per_frame": n, "global_frames_per_step": world * F,
                   "parallelism": f"frame-parallel x{world}", "streams_per_gpu": NS,
                   "fps_mode": ("packed: two clouds per SM (dpm_set_fps_mode(3)), chosen for legs of >= 64 steps" if packed else
                                ("automatic (one SM per cloud at this batch); the packed mapping is kept for legs of >= 64 "
                                 "steps, see `sustained`" if can_pack else f"dpm_set_fps_mode({args.fps_mode})")),
                   "l2": f"inputs rotate over a {pool_mb:.0f} MiB pool of {nslots} batches per GPU (> 126 MB L2)"},
        "e2e": e2e, "batch1": batch1, "gpu_launches": int(launches), "launches_per_step": launches / K,
        "clocks": clk.summary(), "roofline": roofline, "roofline_tensor": roofline_tensor,
        "index_ops": {"fps_plus_knn_GBps": idx_gbps, "frac_of_peak": idx_gbps / peak if idx_gbps else None,
                      "algorithmic_bytes_per_frame": idx_bytes, "fps_ms_per_step": fps_ms, "knn_ms_per_step": knn_ms,
                      **index_kernels},
        "kernels_ms_per_step": [{"kernel": tag, "a": a, "b": b, "launches": cnt, "ms": round(ms, 4)} for ms, cnt, tag, a, b in kern[:args.kernels]],
        "kernel_totals_ms_per_step": kern_totals,
        "kernel_totals_concurrent_ms_per_step": kern_totals_conc,
        "profiled_step_ms": prof_total, "wall_ms_per_step": 1e3 * t_wall / K,
        "sustained": sustained, "strong": strong, "caller_sizes": callers, "index_ops_by_cloud": cube,
        "reference_gpu": ref_gpu, "pipeline_infer": pipe,
        "vs_reference_gpu": (value / ref_gpu["frames_per_s"]) if (ref_gpu and ref_gpu.get("frames_per_s") and world == 1) else None,
    }

    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        frames = args.cpu_frames or max(1, min(8, cores // 4))
        cpu_steps = 16  # ~10-20 s of host work on a 16-core box
        r = run_cpu_sample(cfg, n, frames, cpu_steps, 1)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": f"{cpu_steps} steps x {frames} frames x {n} pts (encoder + {frames} registrations each) after 1 "
                                          f"warm-up step, oracle port (torch fp32 + C/OpenMP index ops), {r['seconds']:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
