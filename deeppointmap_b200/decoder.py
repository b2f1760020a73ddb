"""Drop-in for the reference's `network.decoder.decoder.Decoder` (seam #1).

Same constructor (`Decoder(args)` reading `args.decoder.*`, `args.loss.tau`, `args.loss.eps_offset`,
network/decoder/decoder.py:12-32), same inference methods and return conventions
(`registration_forward` decoder.py:91-127, `loop_detection_forward` :129-143) and the same
state_dict keys, so the shipped checkpoint loads (`strict=False` as in pipeline/infer.py:65, and
also strict).  Modules only hold parameters; each call is one entry into libdpm_b200.so.
The training-only `forward` raises exactly like the reference does in eval mode.
"""
import ctypes
from typing import List, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from . import _C
from .encoder import _cfg_get


def _conv_head(cin: int, cout: int) -> nn.Sequential:  # Conv1d - ReLU - Conv1d (heads.py:6-19)
    return nn.Sequential(nn.Conv1d(cin, cout, 1), nn.Identity(), nn.Conv1d(cout, cout, 1))


class _AttnLayer(nn.Module):  # parameter layout of DescriptorAttentionLayer (descriptor_attention.py:9-23)
    def __init__(self, c: int, heads: int):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(embed_dim=c, num_heads=heads, batch_first=True, dropout=0)
        self.cross_attn = nn.MultiheadAttention(embed_dim=c, num_heads=heads, batch_first=True, dropout=0)
        self.mlp = nn.Sequential(nn.Linear(c, c), nn.Identity(), nn.Linear(c, c))
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(c), nn.LayerNorm(c), nn.LayerNorm(c)


class _OffsetHead(nn.Module):  # heads.py:22-33
    def __init__(self, e: int, coor_dim: int = 3):
        super().__init__()
        self.mlp = nn.Sequential(nn.Conv1d(e, e // 2, 1), nn.Identity(), nn.Conv1d(e // 2, e // 4, 1), nn.Identity(),
                                 nn.Conv1d(e // 4, e // 8, 1))
        self.downsample = nn.Conv1d(e, e // 8, 1)
        self.head = nn.Conv1d(e // 8, coor_dim, 1)


class _OverlapHead(nn.Module):  # heads.py:45-58
    def __init__(self, c: int):
        super().__init__()
        self.mlp = _conv_head(c, c)
        self.projection = nn.Sequential(nn.Linear(2 * c, 2 * c), nn.Identity(), nn.Linear(2 * c, 1), nn.Identity())


class Decoder(nn.Module):
    HEADS = 8  # descriptor_attention.py:14

    def __init__(self, args):
        super().__init__()
        self.args = args
        cfg = _cfg_get(args, "decoder")
        loss = _cfg_get(args, "loss")
        self.decoder_cfg = cfg
        self.in_channel = int(_cfg_get(cfg, "in_channel"))
        self.model_channel = int(_cfg_get(cfg, "model_channel"))
        self.attention_layers = int(_cfg_get(cfg, "attention_layers"))
        self.tau = float(_cfg_get(loss, "tau"))
        self.eps_offset = float(_cfg_get(loss, "eps_offset"))
        c = self.model_channel
        if c != self.HEADS * 32:
            raise NotImplementedError("libdpm_b200 attention is built for head_dim 32 (model_channel 256, 8 heads)")

        self.projection = nn.Conv1d(self.in_channel, c, kernel_size=1)
        self.descriptor_attention = nn.ModuleList([_AttnLayer(c, self.HEADS) for _ in range(self.attention_layers)])
        self.similarity_head = _conv_head(c, c)
        self.offset_head = _OffsetHead(2 * c)
        self.loop_head = _OverlapHead(c)
        self.coarse_pairing_head = _conv_head(self.in_channel, self.in_channel)

        d = _C.DecoderDesc()
        d.in_channel, d.model_channel, d.attention_layers, d.heads = self.in_channel, c, self.attention_layers, self.HEADS
        d.tau, d.eps_offset = self.tau, self.eps_offset
        self._desc = d
        self._wcache = None

    # ---- weights -------------------------------------------------------------------------------
    def _dim_t(self, device) -> Tensor:
        npf = self.model_channel // 3 // 2 * 2  # descriptor_attention.py:61, 70-71
        t = torch.arange(npf, dtype=torch.float32)
        t = 10000 ** (2 * torch.div(t, 2, rounding_mode="trunc") / npf)
        return t.to(device).contiguous()

    def _ordered_params(self) -> List[Tensor]:
        sd = dict(self.named_parameters())
        names = ["projection.weight", "projection.bias"]
        for l in range(self.attention_layers):
            p = f"descriptor_attention.{l}"
            for a in ("self_attn", "cross_attn"):
                names += [f"{p}.{a}.in_proj_weight", f"{p}.{a}.in_proj_bias", f"{p}.{a}.out_proj.weight",
                          f"{p}.{a}.out_proj.bias"]
            names += [p + ".mlp.0.weight", p + ".mlp.0.bias", p + ".mlp.2.weight", p + ".mlp.2.bias"]
            for n in (1, 2, 3):
                names += [f"{p}.norm{n}.weight", f"{p}.norm{n}.bias"]
        names += ["similarity_head.0.weight", "similarity_head.0.bias", "similarity_head.2.weight", "similarity_head.2.bias"]
        for m in ("mlp.0", "mlp.2", "mlp.4", "downsample", "head"):
            names += [f"offset_head.{m}.weight", f"offset_head.{m}.bias"]
        for m in ("mlp.0", "mlp.2", "projection.0", "projection.2"):
            names += [f"loop_head.{m}.weight", f"loop_head.{m}.bias"]
        names += ["coarse_pairing_head.0.weight", "coarse_pairing_head.0.bias", "coarse_pairing_head.2.weight",
                  "coarse_pairing_head.2.bias"]
        return [sd[n] for n in names]

    def _weights(self, device):
        c = self._wcache
        if c is None or c[0] != device:
            ps = self._ordered_params()
            for p in ps:
                if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("Decoder parameters must be contiguous fp32 tensors on the input's CUDA device "
                                       "(call .to(device) first)")
            dim_t = self._dim_t(device)
            arr = (ctypes.c_void_p * (len(ps) + 1))(*([p.data_ptr() for p in ps] + [dim_t.data_ptr()]))
            c = (device, arr, len(ps) + 1, ps, dim_t, next(_C._epoch_ids))
            self._wcache = c
        return c[1], c[2]

    def _fingerprint(self):
        c = self._wcache
        return (id(self), c[5], sum(p._version for p in c[3]))

    def _apply(self, fn, *a, **kw):
        self._wcache = None
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self._wcache = None
        return super().load_state_dict(*a, **kw)

    def __getstate__(self):  # copy.deepcopy / pickle: the pointer table is rebuilt on first use
        s = self.__dict__.copy()
        s["_wcache"] = None
        return s

    # ---- API -------------------------------------------------------------------------------------
    def forward(self, src_descriptor, dst_descriptor, src_padding_mask=None, dst_padding_mask=None, gt_Rt=None):
        assert self.training, 'forward is not available during inference!'
        raise NotImplementedError("training forward is out of scope of the B200 inference path (decoder.py:40-89)")

    @staticmethod
    def num_pairs(num_sample: Union[int, float], M: int, N: int) -> int:
        """decoder.py:170-178"""
        if isinstance(num_sample, int):
            k = num_sample
        elif isinstance(num_sample, float) and num_sample > 1:
            k = int(num_sample)
        elif isinstance(num_sample, float) and 0 < num_sample <= 1:
            k = int(num_sample * (M + N))
        else:
            raise ValueError(f'Argument `num_sample` with value {num_sample} is not supported')
        return k // 2

    @staticmethod
    def _prep(x: Tensor) -> Tensor:
        return x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()

    @staticmethod
    def _mask(mask: Optional[Tensor], P: int, L: int, dev) -> Optional[Tensor]:
        """key-padding mask (P, L), True = padded -> contiguous uint8 on the device, or None"""
        if mask is None:
            return None
        if mask.dim() == 1:
            mask = mask.unsqueeze(0)
        if tuple(mask.shape) != (P, L):
            raise ValueError(f"padding mask must be ({P}, {L}), got {tuple(mask.shape)}")
        return mask.to(device=dev, dtype=torch.bool).to(torch.uint8).contiguous()

    @torch.no_grad()
    def registration_forward_batch(self, src: Tensor, dst: Tensor, num_sample: Union[int, float] = 0.5,
                                   src_padding_mask: Tensor = None, dst_padding_mask: Tensor = None):
        """P independent pairs in one call, no host sync.  src (P,Cd,M), dst (P,Cd,N) ->
        result (P,16) [R(9) T(3) rmse K' K'' iters], conf (P,2k) with the K'' inlier confidences first."""
        _C.require_cuda(src, dst)
        src, dst = self._prep(src), self._prep(dst)
        P, Cd, M = src.shape
        N = dst.shape[2]
        sm, dm = self._mask(src_padding_mask, P, M, src.device), self._mask(dst_padding_mask, P, N, src.device)
        if dst.shape[0] != P or dst.shape[1] != Cd or Cd != self.in_channel + 3:
            raise ValueError("descriptors must be (P, in_channel+3, L) with matching P")
        k = self.num_pairs(num_sample, M, N)
        dev = src.device
        lib = _C.lib()
        warr, nw = self._weights(dev)
        result = torch.empty((P, _C.REG_STRIDE), dtype=torch.float32, device=dev)
        conf = torch.zeros((P, 2 * max(k, 1)), dtype=torch.float32, device=dev)
        nb = lib.dpm_registration_workspace_bytes(ctypes.byref(self._desc), P, M, N, k)
        if nb == 0:
            _C.check(-1, "registration workspace")
        ws = _C.workspaces.get(dev, nb, f"dec{_C.stream_ptr(dev)}")
        lib.dpm_set_weights_epoch(_C.weights_epoch(self._fingerprint(), ws))
        with torch.cuda.device(dev):
            rc = lib.dpm_registration_forward(ctypes.byref(self._desc), warr, nw, src.data_ptr(), dst.data_ptr(),
                                              _C.ptr(sm), _C.ptr(dm), P, M, N, k, result.data_ptr(), conf.data_ptr(),
                                              ws.data_ptr(), ws.numel(), _C.stream_ptr())
        lib.dpm_set_weights_epoch(0)
        _C.check(rc, "registration_forward")
        return result, conf

    @torch.no_grad()
    def registration_forward(self, src_descriptor: Tensor, dst_descriptor: Tensor,
                             src_padding_mask: Tensor = None, dst_padding_mask: Tensor = None,
                             num_sample: Union[int, float] = 0.5) \
            -> Tuple[Tensor, Tensor, Tensor, Union[List[float], float]]:
        batch = not (src_descriptor.ndim == 2 and dst_descriptor.ndim == 2)
        s = src_descriptor if batch else src_descriptor.unsqueeze(0)
        d = dst_descriptor if batch else dst_descriptor.unsqueeze(0)
        assert s.shape[0] == 1, 'batch size in inference must be 1'
        result, conf = self.registration_forward_batch(s, d, num_sample, src_padding_mask, dst_padding_mask)
        host = result[0].cpu()  # the one sync the API demands (rmse is a Python float)
        R = result[0, _C.REG_R:_C.REG_R + 9].view(3, 3).clone()
        T = result[0, _C.REG_T:_C.REG_T + 3].view(3, 1).clone()
        n_inl = int(host[_C.REG_NINLIER])
        rmse = float(host[_C.REG_RMSE])
        c = conf[0, :n_inl].clone()
        if not batch:
            return R, T, c, rmse
        return R.unsqueeze(0), T.unsqueeze(0), c.unsqueeze(0), [rmse]

    @torch.no_grad()
    def loop_detection_forward(self, src_descriptor: Tensor, dst_descriptor: Tensor,
                               src_padding_mask: Tensor = None, dst_padding_mask: Tensor = None) -> Tensor:
        if src_descriptor.ndim == 2 and dst_descriptor.ndim == 2:
            src_descriptor, dst_descriptor = src_descriptor.unsqueeze(0), dst_descriptor.unsqueeze(0)
        _C.require_cuda(src_descriptor, dst_descriptor)
        src, dst = self._prep(src_descriptor), self._prep(dst_descriptor)
        P, Cd, M = src.shape
        N = dst.shape[2]
        dev = src.device
        lib = _C.lib()
        warr, nw = self._weights(dev)
        sm, dm = self._mask(src_padding_mask, P, M, dev), self._mask(dst_padding_mask, P, N, dev)
        prob = torch.empty((P,), dtype=torch.float32, device=dev)
        nb = lib.dpm_loop_detection_workspace_bytes(ctypes.byref(self._desc), P, M, N)
        ws = _C.workspaces.get(dev, nb, f"dec{_C.stream_ptr(dev)}")
        lib.dpm_set_weights_epoch(_C.weights_epoch(("loop",) + self._fingerprint(), ws))
        with torch.cuda.device(dev):
            rc = lib.dpm_loop_detection_forward(ctypes.byref(self._desc), warr, nw, src.data_ptr(), dst.data_ptr(),
                                                _C.ptr(sm), _C.ptr(dm), P, M, N, prob.data_ptr(), ws.data_ptr(), ws.numel(),
                                                _C.stream_ptr())
        lib.dpm_set_weights_epoch(0)
        _C.check(rc, "loop_detection_forward")
        return prob
