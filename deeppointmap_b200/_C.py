"""ctypes binding of libdpm_b200.so (include/dpm_b200.h).

There is no fallback: if the library is missing or the tensors are not on a CUDA device the
calls raise.  Build it with `python -m deeppointmap_b200.build` (nvcc, sm_100a).
"""
import ctypes
import itertools
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DPM_LIB") or os.path.join(_HERE, "libdpm_b200.so")  # DPM_LIB: instrumented developer build

MAX_STAGES, MAX_BLOCKS = 8, 4
REG_R, REG_T, REG_RMSE, REG_NCORR, REG_NINLIER, REG_ITERS, REG_STRIDE = 0, 9, 12, 13, 14, 15, 16
ACT_NONE, ACT_RELU = 0, 1

_ERR = {-1: "shape", -2: "unsupported", -3: "workspace", -4: "cuda", -5: "argument"}


class EncoderDesc(ctypes.Structure):
    _fields_ = [("n_stages", ctypes.c_int), ("in_channel", ctypes.c_int), ("width", ctypes.c_int),
                ("expansion", ctypes.c_int), ("out_channel", ctypes.c_int), ("upsample_layers", ctypes.c_int),
                ("npoint", ctypes.c_int * MAX_STAGES), ("n_blocks", ctypes.c_int * MAX_STAGES),
                ("radius", (ctypes.c_double * MAX_BLOCKS) * MAX_STAGES),
                ("nsample", (ctypes.c_int * MAX_BLOCKS) * MAX_STAGES)]


class DecoderDesc(ctypes.Structure):
    _fields_ = [("in_channel", ctypes.c_int), ("model_channel", ctypes.c_int), ("attention_layers", ctypes.c_int),
                ("heads", ctypes.c_int), ("tau", ctypes.c_float), ("eps_offset", ctypes.c_float)]


_lib = None
_lock = threading.Lock()

_vp, _i, _f, _sz, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_longlong

_SIGS = {
    "dpm_version": ([], _i),
    "dpm_last_error": ([], ctypes.c_char_p),
    "dpm_launch_count": ([], _ll),
    "dpm_launch_count_reset": ([], None),
    "dpm_prof_begin": ([_vp], _i),
    "dpm_prof_end": ([ctypes.c_char_p, _sz], _i),
    "dpm_set_weights_epoch": ([ctypes.c_ulonglong], None),
    "dpm_fps_f32": ([_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp], _i),
    "dpm_fps_workspace_bytes": ([_i, _i, _i, _i], _sz),
    "dpm_set_fps_mode": ([_i], None),
    "dpm_fps_cluster_capacity": ([], _i),
    "dpm_knn_f32": ([_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp], _i),
    "dpm_knn_radius_f32": ([_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _f, _vp, _vp, _sz, _vp], _i),
    "dpm_ball_query_f32": ([_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _f, _vp, _vp, _vp, _sz, _vp], _i),
    "dpm_knn_workspace_bytes": ([_i, _i, _i, _i], _sz),
    "dpm_linear_f32": ([_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "dpm_linear_workspace_bytes": ([_i, _i], _sz),
    "dpm_linear_ws_f32": ([_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp], _i),
    "dpm_linear_ln_workspace_bytes": ([_i, _i, _i], _sz),
    "dpm_linear_ln_ws_f32": ([_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp], _i),
    "dpm_layernorm_f32": ([_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp], _i),
    "dpm_group_ln_relu_max_f32": ([_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _f, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "dpm_fp_interp_f32": ([_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "dpm_encoder_forward": ([ctypes.POINTER(EncoderDesc), _vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _f, _vp,
                             _vp, _vp, _sz, _vp], _i),
    "dpm_encoder_workspace_bytes": ([ctypes.POINTER(EncoderDesc), _i, _i], _sz),
    "dpm_encoder_num_weights": ([ctypes.POINTER(EncoderDesc)], _i),
    "dpm_encoder_out_points": ([ctypes.POINTER(EncoderDesc)], _i),
    "dpm_registration_forward": ([ctypes.POINTER(DecoderDesc), _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp,
                                  _sz, _vp], _i),
    "dpm_registration_workspace_bytes": ([ctypes.POINTER(DecoderDesc), _i, _i, _i, _i], _sz),
    "dpm_loop_detection_forward": ([ctypes.POINTER(DecoderDesc), _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _sz,
                                    _vp], _i),
    "dpm_loop_detection_workspace_bytes": ([ctypes.POINTER(DecoderDesc), _i, _i, _i], _sz),
    "dpm_decoder_num_weights": ([ctypes.POINTER(DecoderDesc)], _i),
    "dpm_posenc_f32": ([_vp, _i, _vp, _i, _vp, _i, _i, _vp], _i),
    "dpm_attention_f32": ([_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp], _i),
    "dpm_attention_pairs_f32": ([_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp], _i),
    "dpm_information_matrix_workspace_bytes": ([_i, _i], _sz),
    "dpm_information_matrix_f32": ([_vp, _i, _vp, _i, _vp, _f, _vp, _vp, _vp, _sz, _vp], _i),
    "dpm_frontend_workspace_bytes": ([_ll], _sz),
    "dpm_frontend_f32": ([_vp, _i, _i, _f, _f, _f, _f, _ll, _vp, _vp, _vp, _sz, _vp], _i),
    "dpm_map_tile_f32": ([_vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp], _i),
    "dpm_outlier_filter_workspace_bytes": ([_i, _i], _sz),
    "dpm_outlier_filter_f32": ([_vp, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _sz, _vp], _i),
    "dpm_low_pass_filter_workspace_bytes": ([_i, _i], _sz),
    "dpm_low_pass_filter_f32": ([_vp, _i, _i, _f, _i, _f, _i, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp], _i),
    "dpm_kabsch_f32": ([_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp], _i),
}

EXPORTS = tuple(_SIGS)


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: the CUDA library has not been built "
                        "(run `python -m deeppointmap_b200.build`). There is no CPU fallback.")
                L = ctypes.CDLL(LIB_PATH)
                for name, (args, res) in _SIGS.items():
                    fn = getattr(L, name)
                    fn.argtypes = args
                    fn.restype = res
                _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().dpm_last_error().decode("utf-8", "replace")
        kind = _ERR.get(rc, str(rc))
        exc = ValueError if rc in (-1, -5) else (NotImplementedError if rc == -2 else RuntimeError)
        raise exc(f"libdpm_b200 {what}: {kind} error: {msg}")


def stream_ptr(device=None) -> int:
    """raw handle of torch's current stream ON `device` (default: the current device)"""
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("deeppointmap_b200 ops run on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")


def ptr(t):
    return 0 if t is None else t.data_ptr()


_buffer_ids = itertools.count(1)
_epoch_ids = itertools.count(1)
_epochs = {}


REUSE_SPLIT = not os.environ.get("DPM_NO_SPLIT_REUSE")


def weights_epoch(fingerprint, ws: torch.Tensor) -> int:
    """A non-zero token that changes whenever the weights' fingerprint or the scratch buffer's identity does (see
    dpm_set_weights_epoch): lets a call skip re-splitting weights it already split into this very buffer."""
    key = (fingerprint, getattr(ws, "_dpm_id", 0))
    # never while a CUDA graph is being captured: a replay must re-split (it is the only way an in-place weight update
    # can reach a captured call)
    if key[1] == 0 or not REUSE_SPLIT or torch.cuda.is_current_stream_capturing():
        return 0
    e = _epochs.get(key)
    if e is None:
        if len(_epochs) > 256:
            _epochs.clear()
        e = _epochs[key] = next(_epoch_ids)
    return e


class _Workspaces(threading.local):
    """One growable scratch buffer per (host thread, device, slot); a slot is "<module><stream handle>", so calls on
    different streams never share scratch.  At most MAX_SLOTS buffers are kept per thread (least recently used goes
    first: a process that rotates over torch's stream pool does not accumulate them) and `release()` drops them all;
    a dropped buffer goes back to torch's caching allocator, which keeps it alive for the kernels already queued on
    the stream it was last used on (record_stream)."""
    MAX_SLOTS = 24

    def __init__(self):
        self.buf = {}  # insertion-ordered: least recently used first

    def get(self, device, nbytes: int, slot: str = "default") -> torch.Tensor:
        key = (device.index if device.index is not None else torch.cuda.current_device(), slot)
        b = self.buf.pop(key, None)
        if b is None or b.numel() < nbytes:
            b = torch.empty(int(nbytes) + (int(nbytes) >> 4) + 4096, dtype=torch.uint8, device=device)
            b.record_stream(torch.cuda.current_stream(device))
            b._dpm_id = next(_buffer_ids)   # identity of THIS allocation (an address can be reused by a later one)
        self.buf[key] = b
        while len(self.buf) > self.MAX_SLOTS:
            self.buf.pop(next(iter(self.buf)))
        return b

    def release(self, device=None) -> int:
        """drop this thread's scratch buffers (of one device, or all); returns the bytes handed back"""
        idx = None if device is None else torch.device(device).index
        keys = [k for k in self.buf if idx is None or k[0] == idx]
        freed = sum(self.buf.pop(k).numel() for k in keys)
        return freed

    def held_bytes(self) -> int:
        return sum(b.numel() for b in self.buf.values())


workspaces = _Workspaces()


def launch_count() -> int:
    return int(lib().dpm_launch_count())


def launch_count_reset() -> None:
    lib().dpm_launch_count_reset()


def prof_begin() -> None:
    check(lib().dpm_prof_begin(stream_ptr()), "prof_begin")


def prof_end():
    """-> list of (kernel, a, b, ms), one per launch since prof_begin (synchronises)."""
    buf = ctypes.create_string_buffer(1 << 20)
    n = lib().dpm_prof_end(buf, len(buf))
    if n < 0:
        check(n, "prof_end")
    out = []
    for line in buf.value.decode().splitlines():
        t, a, b, ms = line.split()
        out.append((t, int(a), int(b), float(ms)))
    return out
