"""The three drop-in seams of SURVEY.md section 8b: (1) `network.encoder.encoder.Encoder` /
`network.decoder.decoder.Decoder`, (2) the reference's Sampler / Querier registry flipping to its
`-t3d` branches, (3) a `pytorch3d.ops` package -- all resolving to libdpm_b200.so."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import REF, ROOT
from oracle import index_ops as IO

HAS_REF = os.path.isdir(os.path.join(REF, "network"))
COMPAT = os.path.join(ROOT, "deeppointmap_b200", "compat")
DROPIN = os.path.join(ROOT, "deeppointmap_b200", "dropin")
SHIMS = os.path.join(ROOT, "deeppointmap_b200", "compat_shims")  # LAST on the path: installed packages win


def _run(code: str) -> str:
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, DROPIN, COMPAT, REF, SHIMS]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp", timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


@pytest.mark.reference
@pytest.mark.skipif(not HAS_REF, reason="/root/reference not present on this box")
def test_unmodified_reference_resolves_to_b200_modules_and_ops():
    out = _run(
        "from network.encoder.encoder import Encoder\n"
        "from network.decoder.decoder import Decoder\n"
        "from network.encoder import utils as RU\n"
        "import network.loss, network.encoder.pointnext as PN\n"
        "print(Encoder.__module__, Decoder.__module__)\n"
        "print(RU.__file__, PN.__file__)\n"
        "print(RU.Sampler('fps-t3d').sample_method.__name__, RU.Querier('hybrid-t3d').query_method.__name__,\n"
        "      RU.Querier('knn-t3d').query_method.__name__, RU.Querier('ball-t3d').query_method.__name__)\n"
        "print(RU.knn_points.__module__, RU.sample_farthest_points.__module__, RU.ball_query.__module__)\n"
        "from pytorch3d.ops.knn import knn_points as k2\n"
        "print(k2.__module__)\n")
    l = out.strip().splitlines()
    assert l[0] == "deeppointmap_b200.encoder deeppointmap_b200.decoder"
    assert l[1].startswith(REF) and l[1].count(REF) == 2           # the rest of `network` is the reference's own
    assert l[2] == "fps_t3d hybrid_query_t3d knn_query_t3d ball_query_t3d"  # no silent fallback to the torch versions
    assert l[3] == "deeppointmap_b200.ops deeppointmap_b200.ops deeppointmap_b200.ops"
    assert l[4] == "deeppointmap_b200.ops"


def test_compat_package_imports_without_reference():
    out = _run("import pytorch3d, pytorch3d.ops as o\nprint(sorted(o.__all__))\n"
               "from network.encoder.encoder import Encoder\nprint(Encoder.__module__)\n")
    assert "['ball_query', 'knn_gather', 'knn_points', 'sample_farthest_points']" in out
    assert "deeppointmap_b200.encoder" in out


# ---- GPU: the pytorch3d.ops contract as the reference's -t3d branches consume it -------------
@pytest.fixture()
def t3d():
    sys.path.insert(0, COMPAT)
    import pytorch3d.ops as o
    yield o
    sys.path.remove(COMPAT)


@pytest.mark.gpu
def test_registry_branches_on_b200_ops(t3d):
    """Restates Sampler.fps_t3d (utils.py:273-285) and Querier.hybrid_query_t3d / ball_query_t3d
    (:100-123) on top of the drop-in pytorch3d.ops and compares with the oracle."""
    from deeppointmap_b200 import data, ops
    dev = "cuda:0"
    B, N, S, K, r = 2, 5000, 300, 32, 0.12
    pts = torch.stack([data.kitti_shape_cloud(11, N).T, data.uniform_cube_cloud(12, N).T * 0.5]).contiguous()
    pad = torch.zeros(B, N, dtype=torch.bool)
    pad[1, 4000:] = True
    lengths = (~pad).sum(1)
    # Sampler.fps_t3d
    smp, idx = t3d.sample_farthest_points(points=pts.to(dev), lengths=lengths.to(dev), K=S, random_start_point=False)
    want = IO.fps(pts, lengths, S)
    assert torch.equal(idx.cpu(), want)
    assert torch.equal(smp.cpu(), torch.gather(pts, 1, want.unsqueeze(-1).expand(-1, -1, 3)))
    ctr = smp
    # Querier.hybrid_query_t3d
    res = t3d.knn_points(p1=ctr[..., :3], p2=pts.to(dev)[..., :3], lengths2=lengths.to(dev), K=K, return_nn=False,
                         return_sorted=False)
    gi, d = res.idx.clone(), res.dists
    m = d > (r ** 2)
    gi[m] = gi[:, :, :1].repeat(1, 1, K)[m]
    want_h = IO.hybrid(ctr.cpu(), pts, lengths, K, r)
    assert torch.equal(gi.cpu(), want_h)
    assert torch.equal(ops.hybrid_query(r, K, pts.to(dev), ctr, pad.to(dev)).cpu(), want_h)  # the fused form
    wd, wi = IO.knn(ctr.cpu(), pts, lengths, K)
    assert torch.equal(res.idx.cpu(), wi) and torch.equal(res.dists.cpu(), wd)
    # Querier.ball_query_t3d
    bq = t3d.ball_query(p1=ctr[..., :3], p2=pts.to(dev)[..., :3], lengths2=lengths.to(dev), K=K, radius=r, return_nn=False)
    wbd, wbi = IO.ball_query(ctr.cpu(), pts, lengths, K, r)
    assert torch.equal(bq.idx.cpu(), wbi) and torch.equal(bq.dists.cpu(), wbd)
    # knn_gather
    g = t3d.knn_gather(pts.to(dev), res.idx)
    assert torch.equal(g.cpu(), pts[torch.arange(B)[:, None, None], wi])


@pytest.mark.gpu
def test_reference_encoder_on_b200_ops_matches_dropin_encoder():
    """Seam #2 for real: the REFERENCE's own Encoder (its pointnext.py / utils.py, unmodified) on the GPU,
    with `pytorch3d.ops` resolving to libdpm_b200.so -- its Sampler / Querier registry picks the `-t3d` branches -- against
    the drop-in Encoder on the same frame.  Same index contract on both sides, so the descriptors agree to 1e-4."""
    from oracle import ref_loader
    if ref_loader.ref_root() is None or ref_loader.checkpoint_path() is None:
        pytest.skip("reference sources not on this box (oracle/_ref/reference is made by build())")
    import numpy as np
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("pytorch3d", "network")}
    try:
        for k in saved_mods:
            sys.modules.pop(k, None)
        renc, rdec, rcfg = ref_loader.load_models("cuda:0", ops="b200")
        import network.encoder.utils as RU
        assert RU.knn_points.__module__ == "deeppointmap_b200.ops"           # the seam is live, no pure-torch fallback
        assert type(renc).__module__ == "network.encoder.encoder"
        from deeppointmap_b200 import Encoder
        from oracle import model_ref as M
        cfg = M.default_config()
        ck = torch.load(ref_loader.checkpoint_path(), map_location="cpu")
        ours = Encoder(cfg).eval()
        ours.load_state_dict(ck["encoder"], strict=True)
        ours = ours.to("cuda:0")
        g = np.load(os.path.join(ROOT, "tests", "golden", "sample_pair.npz"))
        pts = torch.from_numpy(g["cloud0"])[None].to("cuda:0")
        pad = torch.zeros(1, pts.shape[2], dtype=torch.bool, device="cuda:0")
        # the reference's 1x1 convolutions go through cuDNN, whose default on this GPU is TF32 (1e-3 relative):
        # switch it off so that the reference itself computes in fp32
        tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            with torch.no_grad():
                rc, rf, rp = renc(pts, pad)
                oc, of, op = ours(pts, pad)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        assert torch.equal(rc, oc) and torch.equal(rp, op)                    # same FPS picks, bit for bit
        assert float((rf - of).abs().max() / rf.abs().max()) < 1e-4
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k.split(".")[0] in ("pytorch3d", "network")]:
            sys.modules.pop(k, None)
        sys.modules.update(saved_mods)
