"""Running the reference's OWN `pipeline/infer.py`, unmodified, on the B200 modules (BASELINE.json config 5,
SURVEY.md section 8f rank 1).

    python -m deeppointmap_b200.pipeline --reference /path/to/DeepPointMap --impl b200 \
           --yaml_file cfg.yaml --weight DeepPointMapAAAI.pth [any other infer.py flag]

`--impl b200`       sys.path = [repo, dropin/, compat/, reference, ..., compat_shims/]: `network.encoder.encoder.Encoder`
                    and `network.decoder.decoder.Decoder` resolve to the drop-in modules (seam #1) and `pytorch3d.ops`
                    to libdpm_b200.so (seam #3: the transforms' and the information matrix's `knn_points`).
`--impl reference`  the reference's own modules with its pure-torch sampler / querier (no pytorch3d): the oracle run the
                    trajectory is compared with.
Nothing of the reference is edited or copied: this file only prepares `sys.path` / `sys.argv`, restores the one
name Python 3.10 removed (`collections.Iterable`, pipeline/parameters.py:2) and hands over to `runpy`.

Helpers for the synthetic config-5 sequence: `write_synthetic_sequence` (KITTI-style `<n>.bin` float32 (N,4) files,
dataloader/heads/bin.py:16-17) and `write_yaml` (the shipped SemanticKITTI YAML with the data paths, a transform chain
this image can run, and the loop-closure optimiser -- open3d's pose-graph solver -- switched off).
"""
import os
import sys
from typing import List, Optional, Sequence

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)

#: what SURVEY.md section 8d prescribes for the synthetic config-5 sequence: identical input for both implementations
MINIMAL_TRANSFORMS = {
    "VoxelSample": {"voxel_size": 0.3, "retention": "first"},
    "DistanceSample": {"min_dis": 1.0, "max_dis": 60.0},
    "CoordinatesNormalization": {"ratio": 60.0},
    "ToTensor": {"padding_to": -1},
}

#: the shipped chain minus LowPassFilter (open3d normals) -- every step has a device implementation in ops.py as well
DEFAULT_TRANSFORMS = {
    "VoxelSample": {"voxel_size": 0.3, "retention": "first"},
    "ToGPU": {},
    "DistanceSample": {"min_dis": 1.0, "max_dis": 60.0},
    "OutlierFilter": {"nb_neighbors": 10, "std_ratio": 3.0},
    "CoordinatesNormalization": {"ratio": 60.0},
    "ToCPU": {},
    "ToTensor": {"padding_to": -1},
}


def write_synthetic_sequence(out_dir: str, n_frames: int, n_points: int = 65536, seed: int = 0, device="cpu",
                             world: str = "street"):
    """`n_frames` KITTI-style scans `<i>.bin` (float32 (N,4): x, y, z in metres, intensity 0) of a synthetic corridor
    world seen from a smooth SE(2) trajectory (1 m per frame, yaw rate <= 2 deg per frame).  Returns the ground-truth
    poses (n,4,4) fp64."""
    import numpy as np
    import torch
    from . import sequence
    os.makedirs(out_dir, exist_ok=True)
    gt = sequence.trajectory(n_frames)
    make = sequence.street_world if world == "street" else sequence.corridor_world
    world = make(seed, length_m=float(gt[-1, 0, 3]) + 1.0, device=device)
    for i0 in range(0, n_frames, 64):
        fr = sequence.corridor_frames(world, gt[i0:i0 + 64], n_points, seed=seed * 7919 + i0, scale=1.0, stable=True)  # metres
        for j in range(fr.shape[0]):
            rows = torch.cat([fr[j].T, torch.zeros(n_points, 1, device=fr.device)], dim=1).cpu().numpy().astype(np.float32)
            rows.tofile(os.path.join(out_dir, f"{i0 + j}.bin"))
    np.save(os.path.join(os.path.dirname(os.path.abspath(out_dir)), os.path.basename(out_dir.rstrip("/")) + "_gt.npy"),
            gt.numpy())
    return gt


def write_boomerang_sequence(out_dir: str, scans: Sequence[str], n_frames: int) -> List[int]:
    """A sequence of `n_frames` REAL scans out of a short recording: 0, 1, .., m-1, m-2, .., 0, 1, .. (driving forth
    and back), linked as `<i>.bin`.  Returns the source index of every frame."""
    os.makedirs(out_dir, exist_ok=True)
    m = len(scans)
    order, i, step = [], 0, 1
    for _ in range(n_frames):
        order.append(i)
        if m > 1 and not 0 <= i + step < m:
            step = -step
        i += step if m > 1 else 0
    for k, j in enumerate(order):
        dst = os.path.join(out_dir, f"{k}.bin")
        if os.path.lexists(dst):
            os.remove(dst)
        os.symlink(os.path.abspath(scans[j]), dst)
    return order


def write_yaml(path: str, reference_root: str, src_dirs: Sequence[str], out_dir: str, transforms: Optional[dict] = None,
               loop_closure: bool = False, num_workers: int = 0, base: str = "DeepPointMap_B_Main_SemanticKITTI.yaml",
               slam_overrides: Optional[dict] = None):
    """slam_overrides: entries of the YAML's `slam_system` block, e.g. {"edge_confidence_drop": 0.0, "edge_rmse_drop":
    1e9} so that no scan of a synthetic world (on which the trained network is less confident than on real LiDAR) is
    dropped by the mapping thread."""
    import yaml
    cfg = yaml.safe_load(open(os.path.join(reference_root, "configs", "infer", base)))
    cfg["infer_src"] = list(src_dirs)
    cfg["infer_tgt"] = out_dir
    cfg["num_workers"] = int(num_workers)
    cfg["transforms"] = dict(DEFAULT_TRANSFORMS if transforms is None else transforms)
    cfg["slam_system"]["enable_loop_closure"] = bool(loop_closure)
    cfg["slam_system"]["enable_global_optimization"] = bool(loop_closure)
    cfg["slam_system"].update(slam_overrides or {})
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f, sort_keys=False)
    return path


def setup_path(reference_root: str, impl: str = "b200") -> None:
    """sys.path for a run of the reference's pipeline.  The shims of packages this image lacks go LAST."""
    import collections
    import collections.abc
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable   # pipeline/parameters.py:2 predates Python 3.10
    front: List[str] = [ROOT]
    if impl == "b200":
        front += [os.path.join(PKG, "dropin"), os.path.join(PKG, "compat")]
    elif impl == "reference":
        sys.modules["pytorch3d"] = None   # not installed; make sure OUR ops package is not picked up either
    else:
        raise ValueError(f"impl must be 'b200' or 'reference', not {impl!r}")
    front += [reference_root, os.path.join(reference_root, "pipeline")]
    for p in reversed(front):
        while p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    shims = os.path.join(PKG, "compat_shims")
    if shims not in sys.path:
        sys.path.append(shims)


def main(argv: Optional[List[str]] = None) -> None:
    import argparse
    import runpy
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("--reference", required=True, help="root of the unmodified DeepPointMap checkout")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--script", default="infer.py", help="file under <reference>/pipeline/ to run")
    own, rest = ap.parse_known_args(argv)
    setup_path(own.reference, own.impl)
    script = os.path.join(own.reference, "pipeline", own.script)
    sys.argv = [script] + rest
    if "--thread_safety" not in rest:
        sys.argv.append("--thread_safety")   # infer.py:44-46: otherwise it forces the 'spawn' start method (num_workers 0 here)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()


def run_infer_subprocess(reference_root: str, impl: str, yaml_file: str, weight: str, log_path: Optional[str] = None,
                         timeout: Optional[float] = None) -> dict:
    """One run of the reference's pipeline/infer.py in a fresh interpreter (it parses sys.argv and configures
    multiprocessing at import time).  Returns {'returncode', 'wall_s', 'stage_mean_s': the reference's own per-stage
    timers (ResultLogger.log_time, printed by infer.py:119), 'loop_frames_per_s': 1 / their sum}."""
    import re
    import subprocess
    import time
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-m", "deeppointmap_b200.pipeline", "--reference", reference_root, "--impl", impl,
                        "--yaml_file", yaml_file, "--weight", weight], env=env, cwd=os.path.dirname(os.path.abspath(yaml_file)),
                       capture_output=True, text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    if log_path:
        with open(log_path, "w") as f:
            f.write(r.stdout[-40000:] + "\n==== stderr\n" + r.stderr[-40000:])
    stages = {}
    m = re.search(r"Sequence \d+ End, Time = (.*)", r.stdout + r.stderr)
    if m:
        for name, mean in re.findall(r"(\w+):([0-9.]+)/[0-9.]+s", m.group(1)):
            stages[name] = float(mean)
    tot = sum(stages.values())
    return {"returncode": r.returncode, "wall_s": wall, "stage_mean_s": stages,
            "loop_frames_per_s": (1.0 / tot) if tot > 0 else None, "stderr_tail": r.stderr[-1500:]}


def load_trajectory(out_dir: str, seq: int = 0):
    """(timesteps (n,), poses (n,3,4)) of a finished run (system/modules/recoder.py:76-97)"""
    import numpy as np
    d = os.path.join(out_dir, f"Seq{seq:02}")
    T = np.loadtxt(os.path.join(d, "trajectory.allframes.txt")).reshape(-1, 3, 4)
    steps = np.loadtxt(os.path.join(d, "trajectory.allsteps.txt")).astype(int).reshape(-1)
    return steps, T
