"""Where does the confidence error at map sizes come from?  CUDA vs the fp32 oracle vs an fp64 run of the oracle."""
import math, sys, torch, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import model_ref as M
import test_gpu_decoder as T
from deeppointmap_b200 import Decoder, _C
DEV = "cuda:0"
cfg = M.default_config()
ck = torch.load("oracle/_ref/DeepPointMapAAAI.pth", map_location="cpu")
g = np.load("tests/golden/sample_pair.npz")
d0, d1 = torch.from_numpy(g["desc0"]), torch.from_numpy(g["desc1"])
sd = ck["decoder"]; sd64 = {k: v.double() for k, v in sd.items()}
dec = Decoder(cfg).eval(); dec.load_state_dict(sd, strict=True); dec = dec.to(DEV)
# attention core at long key lengths
for Lq, Lk in ((256, 4096), (4096, 4096)):
    gg = torch.Generator().manual_seed(1)
    q, k, v = torch.randn(Lq, 256, generator=gg), torch.randn(Lk, 256, generator=gg), torch.randn(Lk, 256, generator=gg)
    qq, kk, vv = (t.view(-1, 8, 32).transpose(0, 1).double() for t in (q, k, v))
    ref = (torch.softmax(qq @ kk.transpose(1, 2) / math.sqrt(32), -1) @ vv).transpose(0, 1).reshape(Lq, 256)
    ref32 = (torch.softmax(qq.float() @ kk.float().transpose(1, 2) / math.sqrt(32), -1) @ vv.float()).transpose(0, 1).reshape(Lq, 256)
    out = torch.empty(Lq, 256, device=DEV)
    prob = torch.tensor([0, Lq, 0, Lk], dtype=torch.int32, device=DEV)
    qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
    _C.check(_C.lib().dpm_attention_f32(qd.data_ptr(), 256, kd.data_ptr(), 256, vd.data_ptr(), 256, out.data_ptr(), 256, prob.data_ptr(), 1, Lq, 8, _C.stream_ptr()))
    print("attention", Lq, Lk, "cuda vs fp64", float((out.cpu().double() - ref).abs().max() / ref.abs().max()), "torch fp32 vs fp64", float((ref32.double() - ref).abs().max() / ref.abs().max()))
for mb, nb in ((1, 1), (4, 1), (16, 1), (16, 16)):
    src = T._map_of([d0 if i % 2 == 0 else d1 for i in range(mb)], range(100, 100 + mb)) if mb > 1 else d0
    dst = T._map_of([d1 if i % 2 == 0 else d0 for i in range(nb)], range(200, 200 + nb)) if nb > 1 else d1
    tr, tr6 = {}, {}
    R, Tt, c, r = M.registration_forward(sd, cfg, src, dst, 0.5, trace=tr)
    s6, d6, sx6, dx6 = M.attention_forward(sd64, cfg, src[None].double(), dst[None].double())
    si6, di6, conf6, P6 = M.pairing(sd64, cfg, s6, d6, 0.5)
    Rg, Tg, cg, rg = dec.registration_forward(src.to(DEV), dst.to(DEV), num_sample=0.5)
    # top-k confidences sorted (k values): compare the sorted lists
    k = conf6.shape[0]
    res, confg = dec.registration_forward_batch(src[None].to(DEV), dst[None].to(DEV), 0.5)
    print(mb, nb, "oracle32 conf vs fp64", float((tr["conf"].double() - conf6).abs().max()),
          "| inlier conf cuda vs oracle32", float((cg.cpu() - c).abs().max()) if cg.shape == c.shape else (cg.shape, c.shape),
          "| R", float((Rg.cpu() - R).abs().max()), "T", float((Tg.cpu() - Tt).abs().max()))
    # fp64 conf of the inliers selected by the fp32 oracle
    keep, inl = tr["keep"], tr["inlier"]
    P6m = P6[0] if P6.dim() == 3 else P6
    c64 = P6m[tr["src_index"], tr["dst_index"]].repeat(2)[keep][inl]
    print("     inlier conf: oracle32 vs fp64", float((c.double() - c64).abs().max()), " cuda vs fp64", float((cg.cpu().double() - c64).abs().max()) if cg.shape == c64.shape else None)
