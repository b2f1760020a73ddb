"""ctypes front-end of oracle/dpm_oracle.c -- TEST INFRASTRUCTURE ONLY.

The CPU parity checker for the index ops (FPS, kNN, kNN+radius "hybrid", ball
query).  Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product path never touches it.

Reference lines each function follows are listed in dpm_oracle.c's header.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdpm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off (no FMA: parity depends on it) -fopenmp."""
    src = os.path.join(_HERE, "dpm_oracle.c")
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
               "-shared", "-fPIC", "-o", _SO, src, "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        c_fp, c_i64p = ctypes.c_void_p, ctypes.c_void_p
        L.oracle_fps.argtypes = [c_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i64p, ctypes.c_int, c_i64p]
        L.oracle_knn.argtypes = [c_fp, ctypes.c_int, c_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 c_i64p, ctypes.c_int, c_i64p, c_fp]
        L.oracle_hybrid.argtypes = [c_fp, ctypes.c_int, c_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, c_i64p, ctypes.c_int, ctypes.c_float, c_i64p]
        L.oracle_ball_query.argtypes = [c_fp, ctypes.c_int, c_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, c_i64p, ctypes.c_int, ctypes.c_float, c_i64p, c_fp]
        L.oracle_d2.argtypes = [c_fp, c_fp]
        L.oracle_d2.restype = ctypes.c_float
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to("cpu", torch.float32).contiguous()


def _len(lengths, B):
    if lengths is None:
        return None, None
    l = lengths.detach().to("cpu", torch.int64).contiguous()
    assert l.shape == (B,)
    return l, l.data_ptr()


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


def fps(points: torch.Tensor, lengths, K: int) -> torch.Tensor:
    """(B,N,D>=3) fp32 -> idx (B,K) int64, -1 padded.  utils.py:209-270."""
    p = _f32(points)
    B, N, D = p.shape
    idx = torch.empty((B, K), dtype=torch.int64)
    l, lp = _len(lengths, B)
    rc = lib().oracle_fps(p.data_ptr(), B, N, D, lp, K, idx.data_ptr())
    if rc != 0:
        raise ValueError(f"oracle_fps rc={rc}")
    return idx


def knn(p1: torch.Tensor, p2: torch.Tensor, lengths2, K: int):
    """-> (d2 (B,S,K) fp32 ascending, idx (B,S,K) int64), zero padded when lengths2 < K."""
    a, b = _f32(p1), _f32(p2)
    B, S, D1 = a.shape
    _, N, D2 = b.shape
    idx = torch.empty((B, S, K), dtype=torch.int64)
    d2 = torch.empty((B, S, K), dtype=torch.float32)
    l, lp = _len(lengths2, B)
    rc = lib().oracle_knn(a.data_ptr(), D1, b.data_ptr(), D2, B, S, N, lp, K, idx.data_ptr(), d2.data_ptr())
    if rc != 0:
        raise ValueError(f"oracle_knn rc={rc}")
    return d2, idx


def radius2_f32(radius: float) -> float:
    """`dists > radius ** 2` (utils.py:119) compares an fp32 tensor with a Python
    double; torch casts the scalar to the tensor dtype, i.e. to fp32."""
    return float(np.float32(float(radius) ** 2))


def hybrid(p1: torch.Tensor, p2: torch.Tensor, lengths2, K: int, radius: float) -> torch.Tensor:
    """Querier.hybrid_query_t3d (utils.py:112-123) -> idx (B,S,K) int64."""
    a, b = _f32(p1), _f32(p2)
    B, S, D1 = a.shape
    _, N, D2 = b.shape
    idx = torch.empty((B, S, K), dtype=torch.int64)
    l, lp = _len(lengths2, B)
    rc = lib().oracle_hybrid(a.data_ptr(), D1, b.data_ptr(), D2, B, S, N, lp, K,
                             ctypes.c_float(radius2_f32(radius)), idx.data_ptr())
    if rc != 0:
        raise ValueError(f"oracle_hybrid rc={rc}")
    return idx


def ball_query(p1: torch.Tensor, p2: torch.Tensor, lengths2, K: int, radius: float):
    """pytorch3d ball_query contract -> (d2, idx) with -1 / 0 padding."""
    a, b = _f32(p1), _f32(p2)
    B, S, D1 = a.shape
    _, N, D2 = b.shape
    idx = torch.empty((B, S, K), dtype=torch.int64)
    d2 = torch.empty((B, S, K), dtype=torch.float32)
    l, lp = _len(lengths2, B)
    rc = lib().oracle_ball_query(a.data_ptr(), D1, b.data_ptr(), D2, B, S, N, lp, K,
                                 ctypes.c_float(radius2_f32(radius)), idx.data_ptr(), d2.data_ptr())
    if rc != 0:
        raise ValueError(f"oracle_ball_query rc={rc}")
    return d2, idx


def d2(a, b) -> float:
    x = np.asarray(a, dtype=np.float32)
    y = np.asarray(b, dtype=np.float32)
    return float(lib().oracle_d2(x.ctypes.data, y.ctypes.data))
