"""Element-wise error of the per-stage encoder features: GPU vs fp32 oracle vs an fp64 run of the oracle (same indices
injected), in units of the bound 1e-4 * max(|b|, channel rms, 1 % stage rms)."""
import copy, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import model_ref as M
from deeppointmap_b200 import Encoder
cfg = M.default_config()
sd = torch.load("oracle/_ref/DeepPointMapAAAI.pth", map_location="cpu")["encoder"]
g = np.load("tests/golden/sample_pair.npz")
c0 = torch.from_numpy(g["cloud0"])[None]
pad = torch.zeros(1, c0.shape[2], dtype=torch.bool)
tr = {}
M.encoder_forward(sd, cfg, c0, pad, "direct", trace=tr)
sd64 = {k: v.double() for k, v in sd.items()}
tr64 = {}
M.encoder_forward(sd64, cfg, c0.double(), pad, "direct", trace=tr64, inject={"fps_idx": tr["fps_idx"], "knn_idx": tr["knn_idx"]})
for n in range(1, 6):
    sub = copy.deepcopy(cfg); e = sub.encoder
    e.npoint, e.radius_list, e.nsample_list, e.upsample_layers = e.npoint[:n], e.radius_list[:n], e.nsample_list[:n], 0
    enc = Encoder(sub).eval(); enc.load_state_dict(sd, strict=False); enc = enc.to("cuda:0")
    with torch.no_grad():
        fea = enc(c0.cuda(), pad.cuda())[1].cpu().transpose(1, 2).double()
    w32, w64 = tr["fea"][n - 1].double(), tr64["fea"][n - 1]
    rms_c = w64.pow(2).mean(dim=(0, 1), keepdim=True).sqrt()
    bound = 1e-4 * torch.maximum(torch.maximum(w64.abs(), rms_c.expand_as(w64)), (1e-2 * w64.pow(2).mean().sqrt()).expand_as(w64))
    print(f"stage {n}: C={w64.shape[2]}  gpu vs fp64: max {float(((fea - w64).abs() / bound).max()):.2f} bounds   "
          f"oracle32 vs fp64: max {float(((w32 - w64).abs() / bound).max()):.2f}   gpu vs oracle32: max {float(((fea - w32).abs() / bound).max()):.2f}")
