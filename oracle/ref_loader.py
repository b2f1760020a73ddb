"""TEST / BENCH INFRASTRUCTURE (never imported by the product package): locating and importing the UNMODIFIED
reference.

The reference is 32 loose .py files without packaging.  In the build container it lives at /root/reference;
`vendor()` (called by __graft_entry__.build()) copies its Python sources, YAML configs and checkpoint into the
git-ignored `oracle/_ref/reference/` so that the snapshot that travels to the GPU box carries them -- the same way
the oracle's compiled C library travels.  Nothing is copied into the tracked tree.

  ref_root()            -> directory of the reference tree, or None
  activate(ops=...)     -> puts it on sys.path; ops='fallback' hides pytorch3d so the reference degrades to its own
                           pure-torch sampler / querier (network/encoder/utils.py:29-38, 134-143), ops='b200' puts
                           deeppointmap_b200/compat first so its `-t3d` branches run on libdpm_b200.so
  load_models(device)   -> the reference's own Encoder / Decoder with the shipped checkpoint
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
VENDORED = os.path.join(ROOT, "oracle", "_ref", "reference")
PKG = os.path.join(ROOT, "deeppointmap_b200")
_KEEP_DIRS = ("network", "system", "dataloader", "pipeline", "utils", "configs")


def ref_root():
    for p in (SRC, VENDORED):
        if os.path.isdir(os.path.join(p, "network")):
            return p
    return None


def vendor() -> str:
    """copy the reference's sources (.py / .yaml) + checkpoint next to the oracle's build products"""
    if not os.path.isdir(os.path.join(SRC, "network")):
        return VENDORED if os.path.isdir(VENDORED) else ""
    for d in _KEEP_DIRS:
        for base, _, files in os.walk(os.path.join(SRC, d)):
            rel = os.path.relpath(base, SRC)
            for f in files:
                if f.endswith((".py", ".yaml", ".yml")):
                    os.makedirs(os.path.join(VENDORED, rel), exist_ok=True)
                    dst = os.path.join(VENDORED, rel, f)
                    if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(os.path.join(base, f)):
                        shutil.copyfile(os.path.join(base, f), dst)
    sample = os.path.join(SRC, "data", "sample")   # the 11 real KITTI scans the reference ships: the pipeline test's input
    for base, _, files in os.walk(sample):
        rel = os.path.relpath(base, SRC)
        for f in files:
            if f.endswith(".bin"):
                os.makedirs(os.path.join(VENDORED, rel), exist_ok=True)
                dst = os.path.join(VENDORED, rel, f)
                if not os.path.exists(dst):
                    shutil.copyfile(os.path.join(base, f), dst)
    ck, dst = os.path.join(SRC, "DeepPointMapAAAI.pth"), os.path.join(ROOT, "oracle", "_ref", "DeepPointMapAAAI.pth")
    if os.path.exists(ck) and not os.path.exists(dst):
        shutil.copyfile(ck, dst)
    return VENDORED


def sample_frames():
    """sorted paths of the real KITTI scans shipped with the reference (data/sample/seq06/velodyne/*.bin)"""
    root = ref_root()
    d = os.path.join(root, "data", "sample", "seq06", "velodyne") if root else ""
    if not os.path.isdir(d):
        return []
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".bin"))


def checkpoint_path():
    for p in (os.path.join(ROOT, "oracle", "_ref", "DeepPointMapAAAI.pth"), os.path.join(VENDORED, "DeepPointMapAAAI.pth"),
              os.path.join(SRC, "DeepPointMapAAAI.pth")):
        if os.path.exists(p):
            return p
    return None


def activate(ops: str = "fallback", dropin: bool = False) -> str:
    """sys.path for the reference.  The shims of packages this image lacks (colorlog, easydict, readerwriterlock,
    matplotlib, open3d) go LAST, so an installed package always wins."""
    root = ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (neither /root/reference nor oracle/_ref/reference)")
    import collections
    import collections.abc
    if not hasattr(collections, "Iterable"):  # pipeline/parameters.py:2 predates Python 3.10
        collections.Iterable = collections.abc.Iterable
    front = [ROOT]
    if dropin:
        front.append(os.path.join(PKG, "dropin"))
    if ops == "b200":
        front.append(os.path.join(PKG, "compat"))
        for k in [k for k in sys.modules if k == "pytorch3d" or k.startswith("pytorch3d.")]:
            if sys.modules[k] is None:
                del sys.modules[k]
    elif ops == "fallback":
        for k in [k for k in sys.modules if k == "pytorch3d" or k.startswith("pytorch3d.")]:
            del sys.modules[k]
        sys.modules["pytorch3d"] = None  # `import pytorch3d` raises ImportError -> the reference's own fallbacks
    else:
        raise ValueError(ops)
    front += [root, os.path.join(root, "pipeline")]
    for p in reversed(front):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    shims = os.path.join(PKG, "compat_shims")
    if shims not in sys.path:
        sys.path.append(shims)
    return root


def load_config(name: str = "DeepPointMap_B_Main_SemanticKITTI.yaml"):
    import yaml
    from easydict import EasyDict
    return EasyDict(yaml.safe_load(open(os.path.join(ref_root(), "configs", "infer", name))))


def load_models(device="cpu", ops: str = "fallback"):
    """(reference Encoder, reference Decoder, cfg) in eval mode with the shipped weights"""
    import torch
    activate(ops)
    from network.encoder.encoder import Encoder
    from network.decoder.decoder import Decoder
    cfg = load_config()
    ck = torch.load(checkpoint_path(), map_location="cpu")
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    enc.load_state_dict(ck["encoder"], strict=True)
    dec.load_state_dict(ck["decoder"], strict=False)  # pipeline/infer.py:64-65
    return enc.to(device), dec.to(device), cfg
