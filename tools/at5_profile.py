import ctypes, sys, torch
sys.path.insert(0, ".")
from deeppointmap_b200 import _C
lib = _C.lib()
st = torch.cuda.current_stream().cuda_stream
P, M, N = 1, 4096, 4096
R = P * (M + N)
q, k, v = (torch.randn(R, 256, device="cuda") for _ in range(3))
out = torch.empty(R, 256, device="cuda")
for _ in range(3):
    lib.dpm_attention_pairs_f32(q.data_ptr(), 256, k.data_ptr(), 256, v.data_ptr(), 256, out.data_ptr(), 256, P, M, N, 0, 8, None, 2, st)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 16)()
lib.dpm_debug_at5_profile(buf)
nt = buf[8]
names = ["loop top", "wait s_full", "tmem ld S", "mask+max+exchange+exp", "P write+fence+arrive", "wait o_full", "O read+fold"]
tot = sum(buf[i] for i in range(7))
print("tiles", nt, "cycles per tile", tot / nt)
for i, n in enumerate(names):
    print(f"  {n:28s} {buf[i] / nt:8.0f}")
