"""Build libdpm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m deeppointmap_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libdpm_b200.so")
SOURCES = ["core.cu", "fps.cu", "fps_cluster.cu", "knn.cu", "grid.cu", "dense.cu", "gemm_tc.cu", "encoder.cu", "decoder.cu", "attention_tc5.cu", "pairing.cu", "infomat.cu", "frontend.cu", "maptile.cu", "outlier.cu", "lowpass.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dpm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    # developer builds: DPM_BUILD_DEFINES="-DDPM_FPS_PROFILE" DPM_BUILD_SO=/path/libx.so (instrumented copy
    # next to the product library; load it with DPM_LIB=/path/libx.so)
    global SO
    defines = os.environ.get("DPM_BUILD_DEFINES", "").split()
    alt = os.environ.get("DPM_BUILD_SO")
    if alt:
        SO, force = alt, True
    if not force and not needs_build():
        return SO
    objdir = os.path.join(HERE, "build" + ("_alt" if alt else ""))
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + defines + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"==== {src}\n{out}\n")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
