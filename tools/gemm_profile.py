"""Phase timeline of the tcgen05 GEMM CTAs (developer build -DDPM_TC_PROFILE):
   DPM_BUILD_DEFINES=-DDPM_TC_PROFILE DPM_BUILD_SO=$PWD/deeppointmap_b200/libdpm_prof.so python -m deeppointmap_b200.build
   DPM_LIB=$PWD/deeppointmap_b200/libdpm_prof.so python tools/gemm_profile.py"""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, ".")
from deeppointmap_b200 import _C
lib = _C.lib()
st = torch.cuda.current_stream().cuda_stream
for (M, N, K, ln) in [(16384, 256, 256, False), (16384, 256, 256, True), (16384, 768, 256, False), (131072, 64, 64, True), (512, 2048, 512, False), (16384, 256, 32, False)]:
    X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda"); Y = torch.empty(M, N, device="cuda"); g = torch.ones(N, device="cuda")
    nb = lib.dpm_linear_ln_workspace_bytes(M, N, K); ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    def run():
        if ln:
            rc = lib.dpm_linear_ln_ws_f32(X.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), R.data_ptr(), N, g.data_ptr(), b.data_ptr(), None, 0, Y.data_ptr(), N, M, N, K, 0, ws.data_ptr(), nb, st)
        else:
            rc = lib.dpm_linear_ws_f32(X.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), R.data_ptr(), N, Y.data_ptr(), N, M, N, K, 0, ws.data_ptr(), nb, st)
        assert rc == 0, lib.dpm_last_error()
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    nct = ((M + 127) // 128) * ((N + 255) // 256 if N > 128 else 1)
    buf = (ctypes.c_ulonglong * (6 * min(nct, 2048)))()
    assert lib.dpm_debug_tc_profile(buf, nct) == 0
    t = np.array(buf, dtype=np.int64).reshape(-1, 6).astype(np.float64)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    d = np.diff(t, axis=1) / 1e3
    print(f"M={M} N={N} K={K} ln={ln}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (split + gemm); CTAs {len(t)}")
    print("   start spread (us): max", rel[:, 0].max().round(2), " end: median", np.median(rel[:, 5]).round(2), "max", rel[:, 5].max().round(2))
    print("   per-CTA phase us (median): setup %.2f | first stage full %.2f | mainloop (to accum ready) %.2f | epilogue %.2f | teardown %.2f" % tuple(np.median(d, axis=0)))
