"""pipeline/infer.py (unmodified) over one synthetic sequence with both implementations; prints the trajectory
difference and the frames/s of each run.  python tools/run_pipeline_pair.py [frames] [points] [workdir]"""
import os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deeppointmap_b200 import pipeline
from oracle import ref_loader

def load_traj(d):
    T = np.loadtxt(os.path.join(d, "Seq00", "trajectory.allframes.txt")).reshape(-1, 3, 4)
    steps = np.loadtxt(os.path.join(d, "Seq00", "trajectory.allsteps.txt")).astype(int).reshape(-1)
    return steps, T

def run(impl, ref, yaml_file, weight, log):
    env = dict(os.environ, PYTHONPATH=ROOT)
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-m", "deeppointmap_b200.pipeline", "--reference", ref, "--impl", impl, "--yaml_file", yaml_file,
                        "--weight", weight], env=env, cwd=os.path.dirname(yaml_file), capture_output=True, text=True)
    open(log, "w").write(r.stdout[-20000:] + "\n==== stderr\n" + r.stderr[-20000:])
    if r.returncode != 0:
        tail = r.stderr.strip().splitlines()[-1][:300] if r.stderr.strip() else ""
        if not os.path.exists(os.path.join(os.path.dirname(yaml_file), f"out_{impl}", "Seq00", "trajectory.allframes.txt")):
            raise RuntimeError(f"{impl} run failed:\n{r.stderr[-3000:]}")
        print(f"[{impl}] infer.py stopped after the trajectory was saved: {tail}")
    return time.perf_counter() - t0

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    pts = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    work = sys.argv[3] if len(sys.argv) > 3 else "/tmp/dpm_pipeline"
    ref = ref_loader.ref_root()
    weight = ref_loader.checkpoint_path()
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    seq = os.path.join(work, "seq", "0")
    real = ref_loader.sample_frames()
    gt = None
    if pts <= 0 and real:   # points <= 0: the reference's real sample scans, forth and back
        order = pipeline.write_boomerang_sequence(seq, real, n)
    else:
        gt = pipeline.write_synthetic_sequence(seq, n, pts, seed=3, device=dev)
    out = {}
    for impl in ("b200", "reference"):
        y = pipeline.write_yaml(os.path.join(work, f"cfg_{impl}.yaml"), ref, [seq], os.path.join(work, f"out_{impl}"),
                                transforms=pipeline.MINIMAL_TRANSFORMS)
        dt = run(impl, ref, y, weight, os.path.join(work, f"log_{impl}.txt"))
        out[impl] = load_traj(os.path.join(work, f"out_{impl}"))
        print(f"{impl}: {n} frames in {dt:.1f} s wall (process start, model load, data loading and SLAM bookkeeping included) = {n / dt:.2f} frames/s; {len(out[impl][0])} scans in the trajectory")
    (sa, Ta), (sb, Tb) = out["b200"], out["reference"]
    print("same scans kept:", np.array_equal(sa, sb))
    if np.array_equal(sa, sb):
        dt_ = np.linalg.norm(Ta[:, :, 3] - Tb[:, :, 3], axis=1)
        dR = np.abs(Ta[:, :, :3] - Tb[:, :, :3]).max(axis=(1, 2))
        print("trajectory difference b200 vs reference: max |dt| %.4f m, max |dR| %.2e, final position %s vs %s" % (dt_.max(), dR.max(), Ta[-1, :, 3].round(3), Tb[-1, :, 3].round(3)))
        if gt is not None:
            g = gt.numpy()[sa][:, :3, 3] - gt.numpy()[sa][0, :3, 3]
            print("distance to ground truth at the end: b200 %.3f m, reference %.3f m (path %.1f m)" % (np.linalg.norm(Ta[-1, :, 3] - g[-1]), np.linalg.norm(Tb[-1, :, 3] - g[-1]), np.linalg.norm(g[-1])))
