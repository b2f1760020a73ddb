#!/usr/bin/env python
"""Micro-benchmark of the tensor-core linear (dpm_linear_ws_f32) on the decoder / encoder shapes.
Round-1 findings (B200): ~19 us of every call is size-independent (split launch + one 128-row tile's
prologue, 8 K blocks at ~1.2 us each through 2 stages, TMEM round trip); the epilogue adds ~6 us.  A K block
costs max(W stage fill ~1 us, 12 tf32 MMAs ~0.8 us): with hi + lo copies a 128 x 256 tile leaves room for two stages
only, so neither hides behind the other; prefetching the X tile into registers one block ahead changes nothing."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ROOT)
    from deeppointmap_b200 import _C
    lib = _C.lib()
    st = torch.cuda.current_stream().cuda_stream
    for (M, N, K, res) in [(16384, 256, 256, True), (16384, 256, 256, False), (16384, 768, 256, False), (512, 2048, 512, False),
                           (131072, 128, 32, False), (8192, 256, 256, True), (2048, 1024, 256, False), (32768, 64, 64, False)]:
        X = torch.randn(M, K, device="cuda")
        W = torch.randn(N, K, device="cuda") / K ** 0.5
        b = torch.randn(N, device="cuda")
        R = torch.randn(M, N, device="cuda") if res else None
        Y = torch.empty(M, N, device="cuda")
        nb = lib.dpm_linear_workspace_bytes(N, K)
        ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
        def run():
            rc = lib.dpm_linear_ws_f32(X.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), R.data_ptr() if res else None, N,
                                       Y.data_ptr(), N, M, N, K, 0, ws.data_ptr(), nb, st)
            assert rc == 0, lib.dpm_last_error()
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        want = X.double() @ W.double().T + b.double() + (R.double() if res else 0)
        err = float((Y.double() - want).abs().max() / want.abs().max())
        print(f"  M={M} N={N} K={K} res={res}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (split + gemm), rel err {err:.1e}", flush=True)
else:
    subprocess.run([sys.executable, __file__, "child"])
