"""attention core at map sizes: mma.sync kernel (impl 1) vs tcgen05 / TMEM flash attention (impl 2).
flops = 4 * C * sum over (query range x key range) (QK^T + PV), C = 256."""
import sys
import torch
sys.path.insert(0, ".")
from deeppointmap_b200 import _C
lib = _C.lib()
st = torch.cuda.current_stream().cuda_stream
for (P, M, N, mode) in [(1, 4096, 256, 0), (1, 4096, 256, 1), (1, 4096, 4096, 0), (1, 4096, 4096, 1), (4, 2048, 2048, 0), (32, 256, 256, 0), (1, 1024, 1024, 0)]:
    R = P * (M + N)
    q, k, v = (torch.randn(R, 256, device="cuda") for _ in range(3))
    out = torch.empty(R, 256, device="cuda")
    fl = 4.0 * 256 * P * ((M * M + N * N) if mode == 0 else 2.0 * M * N)
    res = []
    for impl in (1, 2):
        def run():
            rc = lib.dpm_attention_pairs_f32(q.data_ptr(), 256, k.data_ptr(), 256, v.data_ptr(), 256, out.data_ptr(), 256, P, M, N, mode, 8, None, impl, st)
            assert rc == 0, lib.dpm_last_error()
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res.append((ms, fl / ms / 1e9))
    print(f"P={P} M={M} N={N} mode={mode}: mma.sync {res[0][0]:.3f} ms ({res[0][1]:.1f} TFLOP/s)   tcgen05 {res[1][0]:.3f} ms ({res[1][1]:.1f} TFLOP/s)   x{res[0][0] / res[1][0]:.2f}", flush=True)
