"""Error behaviour of the C-ABI boundary: negative return code + thread-local message, mapped by the Python layer to
the exception types the reference's callers see (ValueError for shapes / arguments, NotImplementedError for
unsupported configurations, RuntimeError for workspace / CUDA problems); nothing is written on failure."""
import ctypes

import pytest
import torch

from deeppointmap_b200 import Decoder, Encoder, _C, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _err(rc):
    return rc, _C.lib().dpm_last_error().decode()


def test_error_codes_and_messages():
    lib = _C.lib()
    x = torch.zeros(4, 8, 3, device=DEV)
    idx = torch.zeros(4, 4, dtype=torch.int64, device=DEV)
    ws = torch.zeros(1 << 24, dtype=torch.uint8, device=DEV)
    st = _C.stream_ptr()
    rc, msg = _err(lib.dpm_fps_f32(None, 4, 8, 3, None, 4, idx.data_ptr(), None, ws.data_ptr(), ws.numel(), st))
    assert rc == -5 and "null" in msg
    rc, msg = _err(lib.dpm_fps_f32(x.data_ptr(), 4, 8, 2, None, 4, idx.data_ptr(), None, ws.data_ptr(), ws.numel(), st))
    assert rc == -1 and "D=2" in msg
    rc, msg = _err(lib.dpm_fps_f32(x.data_ptr(), 4, 8, 3, None, 4, idx.data_ptr(), None, ws.data_ptr(), 16, st))
    assert rc == -3 and "workspace" in msg
    rc, msg = _err(lib.dpm_knn_f32(x.data_ptr(), 3, x.data_ptr(), 3, 4, 8, 8, None, None, 0, idx.data_ptr(), None, ws.data_ptr(),
                                   ws.numel(), st))
    assert rc == -1 and "K=0" in msg
    # DPM_ERR_UNSUPPORTED: a leading dimension the attention kernels cannot take (checked before any launch)
    rc, msg = _err(lib.dpm_attention_f32(x.data_ptr(), 3, x.data_ptr(), 3, x.data_ptr(), 3, x.data_ptr(), 3, idx.data_ptr(), 1, 8,
                                         1, st))
    assert rc == -2 and "multiples of 4" in msg
    rc, msg = _err(lib.dpm_information_matrix_f32(x.data_ptr(), 0, x.data_ptr(), 8, x.data_ptr(), 1.0, x.data_ptr(), None,
                                                  ws.data_ptr(), ws.numel(), st))
    assert rc == -1
    rc, msg = _err(lib.dpm_frontend_f32(x.data_ptr(), 8, 2, 0.3, 1.0, 60.0, 60.0, 1 << 20, x.data_ptr(), idx.data_ptr(),
                                        ws.data_ptr(), ws.numel(), st))
    assert rc == -1 and "stride" in msg
    rc, msg = _err(lib.dpm_linear_ws_f32(x.data_ptr(), 3, x.data_ptr(), 3, None, None, 0, x.data_ptr(), 3, 8, 3, 3, 0,
                                         ws.data_ptr(), 8, st))
    assert rc == -3


def test_python_layer_exception_types(cfg):
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.sample_farthest_points(torch.zeros(1, 8, 3), K=2)
    with pytest.raises(ValueError):
        ops.sample_farthest_points(torch.zeros(1, 8, 2, device=DEV), K=2)
    with pytest.raises(NotImplementedError):
        ops.sample_farthest_points(torch.zeros(1, 8, 3, device=DEV), K=2, random_start_point=True)
    with pytest.raises(ValueError):
        ops.knn_points(torch.zeros(1, 8, 3, device=DEV), torch.zeros(1, 8, 3, device=DEV), K=0)
    with pytest.raises(ValueError):
        ops.information_matrix(torch.zeros(2, 8, device=DEV), torch.zeros(3, 8, device=DEV), torch.eye(4))
    enc, dec = Encoder(cfg).eval().to(DEV), Decoder(cfg).eval().to(DEV)
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 2, 64, device=DEV), torch.zeros(1, 64, dtype=torch.bool, device=DEV))     # fewer than 3 channels
    with pytest.raises(AssertionError):                                                               # decoder.py:169
        dec.registration_forward(torch.zeros(2, 131, 256, device=DEV), torch.zeros(2, 131, 256, device=DEV))
    with pytest.raises(AssertionError):                                                               # decoder.py:37
        dec.eval()(torch.zeros(1, 131, 256, device=DEV), torch.zeros(1, 131, 256, device=DEV))
    with pytest.raises(RuntimeError):                                                                 # parameters left on the CPU
        Encoder(cfg).eval()(torch.zeros(1, 3, 64, device=DEV), torch.zeros(1, 64, dtype=torch.bool, device=DEV))
