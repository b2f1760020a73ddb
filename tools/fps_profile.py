"""phase timing of the cluster FPS (developer build with -DDPM_FC_PROFILE):
   DPM_BUILD_DEFINES=-DDPM_FC_PROFILE DPM_BUILD_SO=$PWD/gpurun_out/libprof.so python -m deeppointmap_b200.build
   DPM_LIB=$PWD/gpurun_out/libprof.so python tools/fps_profile.py"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from deeppointmap_b200 import _C, data, ops  # noqa: E402

lib = _C.lib()
lib.dpm_debug_fc_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 16)()
names = ["box test + ballot", "tile update (phase A)", "warp arg-max", "publish", "bucket maxima (phase B)", "wait + collect", "-", "loop / writer"]
ops.set_fps_mode(2)
for kind, n, k in (("kitti", 65536, 4096), ("cube", 65536, 4096), ("kitti", 16384, 4096)):
    mk = data.kitti_shape_cloud if kind == "kitti" else data.uniform_cube_cloud
    pts = mk(1, n).T.contiguous()[None].to("cuda:0")
    ops.sample_farthest_points(pts, K=k)
    torch.cuda.synchronize()
    lib.dpm_debug_fc_profile(buf, 1)
    ops.sample_farthest_points(pts, K=k)
    torch.cuda.synchronize()
    lib.dpm_debug_fc_profile(buf, 1)
    picks = buf[8]
    tot = sum(buf[i] for i in range(8)) / 32.0 / max(picks, 1)
    print(f"{kind} N={n} K={k}: {tot:.0f} cycles per pick per warp (avg over 32 warps); {picks} picks in {buf[9]} rounds")
    for i in range(8):
        print(f"   {names[i]:28s} {buf[i] / 32.0 / max(picks, 1):8.1f}")
