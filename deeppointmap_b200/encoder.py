"""Drop-in for the reference's `network.encoder.encoder.Encoder` (seam #1).

Same constructor (`Encoder(args)` reading `args.encoder.*`, network/encoder/encoder.py:11-49),
same `forward(points (B,C>=3,N), points_padding (B,N) bool) -> [coor (B,3,S), fea (B,Cout,S),
pad (B,S)]` (encoder.py:51-69) and the same state_dict keys/shapes, so
`load_state_dict(torch.load('DeepPointMapAAAI.pth')['encoder'], strict=True)` works.  The modules
below only HOLD parameters; the whole forward pass is one call into libdpm_b200.so
(`dpm_encoder_forward`), enqueued on the current CUDA stream with no host synchronisation.
"""
import ctypes
from typing import List, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _C


def _cfg_get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class _LNHolder(nn.Module):
    """state_dict keys '<idx>.ln.weight/bias' (LayerNorm1d/2d, network/encoder/utils.py:392-413)."""

    def __init__(self, channels: int):
        super().__init__()
        self.ln = nn.LayerNorm(channels)


def _mlp_holder(cin: int, channels: List[int], dim: int, drop_last_act: bool = False, bias: bool = True) -> nn.Sequential:
    """Parameter layout of build_mlp (utils.py:358-389): conv, norm, act triplets (acts hold nothing)."""
    conv = nn.Conv1d if dim == 1 else nn.Conv2d
    mods = []
    for c in channels:
        mods += [conv(cin, c, kernel_size=1, bias=bias), _LNHolder(c), nn.Identity()]
        cin = c
    if drop_last_act:
        mods = mods[:-1]
    return nn.Sequential(*mods)


class _SA(nn.Module):
    def __init__(self, cin, bias=True):
        super().__init__()
        self.mlp = _mlp_holder(cin + 3, [2 * cin], 2, bias=bias)


class _LA(nn.Module):
    def __init__(self, c, bias=True):
        super().__init__()
        self.mlp = _mlp_holder(c + 3, [c], 2, bias=bias)


class _IRM(nn.Module):
    def __init__(self, c, expansion, bias=True):
        super().__init__()
        self.la = _LA(c, bias)
        self.pw_conv = _mlp_holder(c, [c * expansion, c], 1, drop_last_act=True, bias=bias)


class _Stage(nn.Module):
    def __init__(self, cin, n_blocks, expansion, bias=True):
        super().__init__()
        self.sa = _SA(cin, bias)
        irm = [_IRM(2 * cin, expansion, bias) for _ in range(n_blocks - 1)]
        self.irm = nn.Sequential(*irm) if irm else nn.Identity()


class _FP(nn.Module):
    def __init__(self, cin, cout, bias=True):
        super().__init__()
        self.mlp = _mlp_holder(cin, [cout, cout], 1, bias=bias)


class Encoder(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        cfg = _cfg_get(args, "encoder")
        self.encoder_cfg = cfg
        self.in_channel = int(_cfg_get(cfg, "in_channel"))
        self.out_channel = int(_cfg_get(cfg, "out_channel"))
        npoint = list(_cfg_get(cfg, "npoint"))
        radius_list = [list(r) for r in _cfg_get(cfg, "radius_list")]
        nsample_list = [list(n) for n in _cfg_get(cfg, "nsample_list")]
        self.downsample_layers = len(npoint)
        self.upsample_layers = int(_cfg_get(cfg, "upsample_layers"))
        width = int(_cfg_get(cfg, "width"))
        expansion = int(_cfg_get(cfg, "expansion"))
        norm = str(_cfg_get(cfg, "norm", "LN")).lower()
        bias = bool(_cfg_get(cfg, "bias", True))
        if norm != "ln":
            raise NotImplementedError("libdpm_b200 implements LayerNorm blocks (norm: LN, the shipped configuration); "
                                      "the BatchNorm / InstanceNorm variants of build_mlp are not on the CUDA path")
        self._bias = bias
        for s in (_cfg_get(cfg, "sample", None) or []):
            t = _cfg_get(s, "type", "fps-t3d")
            if t not in ("fps", "fps-t3d"):
                raise NotImplementedError(f"sampler '{t}' is not on the CUDA path (only fps / fps-t3d)")
        if self.downsample_layers > _C.MAX_STAGES or any(len(r) > _C.MAX_BLOCKS for r in radius_list):
            raise NotImplementedError("too many stages / blocks for dpm_encoder_desc")

        d = _C.EncoderDesc()
        d.n_stages, d.in_channel, d.width, d.expansion = self.downsample_layers, self.in_channel, width, expansion
        d.out_channel, d.upsample_layers = self.out_channel, self.upsample_layers
        for i in range(self.downsample_layers):
            assert len(radius_list[i]) == len(nsample_list[i])
            d.npoint[i] = int(npoint[i])
            d.n_blocks[i] = len(radius_list[i])
            for j, (r, k) in enumerate(zip(radius_list[i], nsample_list[i])):
                d.radius[i][j] = float(r)
                d.nsample[i][j] = int(k)
        self._desc = d

        self.point_mlp0 = nn.Conv1d(self.in_channel, width, kernel_size=1)
        self.downsampler = nn.ModuleList()
        self.upsampler = nn.ModuleList()
        w = width
        for i in range(self.downsample_layers):
            self.downsampler.append(_Stage(w, len(radius_list[i]), expansion, bias))
            w *= 2
        up_in = w
        for _ in range(self.upsample_layers):
            up_out = max(self.out_channel, w // 2)
            self.upsampler.append(_FP(up_in + w // 2, up_out, bias))
            w //= 2
            up_in = up_out
        self.final_channel = up_in if self.upsample_layers > 0 else w
        lvl = self.downsample_layers - self.upsample_layers
        self._out_points = None if lvl <= 0 else int(npoint[lvl - 1])
        self._wcache = None
        self.trace = False          # when True, forward() records FPS / group indices in last_trace
        self.last_trace: Optional[dict] = None

    # ---- weight pointer table (canonical order = reference state_dict order) ---------------
    def _ordered_params(self) -> List[Tensor]:
        sd = dict(self.named_parameters())
        names = ["point_mlp0.weight", "point_mlp0.bias"]

        def ln(p):
            return [p + ".ln.weight", p + ".ln.bias"]

        for i, stage in enumerate(self.downsampler):
            p = f"downsampler.{i}"
            names += [p + ".sa.mlp.0.weight", p + ".sa.mlp.0.bias"] + ln(p + ".sa.mlp.1")
            for j in range(self._desc.n_blocks[i] - 1):
                q = f"{p}.irm.{j}"
                names += [q + ".la.mlp.0.weight", q + ".la.mlp.0.bias"] + ln(q + ".la.mlp.1")
                names += [q + ".pw_conv.0.weight", q + ".pw_conv.0.bias"] + ln(q + ".pw_conv.1")
                names += [q + ".pw_conv.3.weight", q + ".pw_conv.3.bias"] + ln(q + ".pw_conv.4")
        for i in range(self.upsample_layers):
            p = f"upsampler.{i}"
            names += [p + ".mlp.0.weight", p + ".mlp.0.bias"] + ln(p + ".mlp.1")
            names += [p + ".mlp.3.weight", p + ".mlp.3.bias"] + ln(p + ".mlp.4")
        if self._bias:
            return [sd[n] for n in names]
        # bias: False (utils.py:358-389 builds the convolutions without one): the kernels take a pointer per slot of
        # the canonical table, so the missing conv biases are zero vectors (x + 0 is exact) kept as non-persistent
        # buffers -- they follow .to(device) and stay out of the state_dict
        out = []
        for i, n in enumerate(names):
            if n in sd:
                out.append(sd[n])
                continue
            c = sd[names[i - 1]].shape[0]  # the conv weight right before it
            key = f"_zero_bias_{c}"
            if not hasattr(self, key):
                self.register_buffer(key, torch.zeros(c, dtype=torch.float32, device=sd[names[i - 1]].device), persistent=False)
            out.append(getattr(self, key))
        return out

    def _weights(self, device):
        c = self._wcache
        if c is None or c[0] != device:
            ps = self._ordered_params()
            for p in ps:
                if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("Encoder parameters must be contiguous fp32 tensors on the input's CUDA device "
                                       "(call .to(device) first)")
            arr = (ctypes.c_void_p * len(ps))(*[p.data_ptr() for p in ps])
            c = (device, arr, len(ps), ps, next(_C._epoch_ids))
            self._wcache = c
        return c[1], c[2]

    def _fingerprint(self):
        """identity of the current weight VALUES as far as torch can tell: this module, this pointer table, and the
        parameters' in-place version counters (an optimiser step or `p.mul_()` bumps them)"""
        c = self._wcache
        return (id(self), c[4], sum(p._version for p in c[3]))

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

    # ---- forward -----------------------------------------------------------------------------
    def _run(self, points: Tensor, points_padding: Optional[Tensor], want_desc: bool, coor_scale: float,
             desc_out: Optional[Tensor] = None):
        if self.training and torch.is_grad_enabled():
            # pipeline/train.py would otherwise get detached features without any notice
            raise NotImplementedError("deeppointmap_b200.Encoder is inference-only (no autograd path): call .eval() "
                                      "and / or run under torch.no_grad()")
        _C.require_cuda(points, points_padding)
        if points.dim() != 3 or points.shape[1] < 3:
            raise ValueError("points must be (B, C>=3, N)")
        pts = points if (points.dtype == torch.float32 and points.is_contiguous()) else points.float().contiguous()
        B, C, N = pts.shape
        pad = None
        if points_padding is not None:
            if points_padding.shape != (B, N):
                raise ValueError("points_padding must be (B, N)")
            pad = points_padding.to(torch.bool).contiguous()
        dev = pts.device
        S = self._out_points if self._out_points is not None else N
        Cout = self.final_channel
        lib = _C.lib()
        warr, nw = self._weights(dev)
        coor = torch.empty((B, 3, S), dtype=torch.float32, device=dev)
        fea = torch.empty((B, Cout, S), dtype=torch.float32, device=dev)
        opad = torch.empty((B, S), dtype=torch.bool, device=dev)
        desc = None
        if want_desc:
            desc = desc_out if desc_out is not None else torch.empty((B, Cout + 3, S), dtype=torch.float32, device=dev)
            if (desc.shape != (B, Cout + 3, S) or desc.dtype != torch.float32 or desc.device != dev
                    or not desc.is_contiguous()):
                raise ValueError(f"`out` must be a contiguous fp32 ({B}, {Cout + 3}, {S}) tensor on {dev}")
        tf = tk = None
        if self.trace:
            nf = sum(self._desc.npoint[i] for i in range(self._desc.n_stages))
            nk = sum(self._desc.npoint[i] * self._desc.nsample[i][j] for i in range(self._desc.n_stages)
                     for j in range(self._desc.n_blocks[i]))
            tf = torch.empty((B * nf,), dtype=torch.int64, device=dev)
            tk = torch.empty((B * nk,), dtype=torch.int32, device=dev)
        nb = lib.dpm_encoder_workspace_bytes(ctypes.byref(self._desc), B, N)
        if nb == 0:
            _C.check(-1, "encoder workspace")
        ws = _C.workspaces.get(dev, nb, f"enc{_C.stream_ptr(dev)}")
        lib.dpm_set_weights_epoch(_C.weights_epoch(self._fingerprint(), ws))  # unchanged weights + same scratch: no re-split
        with torch.cuda.device(dev):
            rc = lib.dpm_encoder_forward(ctypes.byref(self._desc), warr, nw, pts.data_ptr(), C, _C.ptr(pad), B, N,
                                         coor.data_ptr(), fea.data_ptr(), opad.data_ptr(), _C.ptr(desc),
                                         float(coor_scale), _C.ptr(tf), _C.ptr(tk), ws.data_ptr(), ws.numel(),
                                         _C.stream_ptr())
        lib.dpm_set_weights_epoch(0)
        _C.check(rc, "encoder_forward")
        if self.trace:
            fps_idx, knn_idx, fo, ko = [], [], 0, 0
            for i in range(self._desc.n_stages):
                s = self._desc.npoint[i]
                fps_idx.append(tf[fo:fo + B * s].view(B, s))
                fo += B * s
                for j in range(self._desc.n_blocks[i]):
                    k = self._desc.nsample[i][j]
                    knn_idx.append(tk[ko:ko + B * s * k].view(B, s, k))
                    ko += B * s * k
            self.last_trace = {"fps_idx": fps_idx, "knn_idx": knn_idx}
        return coor, fea, opad, desc

    def forward(self, points: Tensor, points_padding: Tensor) -> List[Tensor]:
        coor, fea, pad, _ = self._run(points, points_padding, False, 1.0)
        return [coor, fea, pad]

    @torch.no_grad()
    def descriptors(self, points: Tensor, points_padding: Optional[Tensor] = None, coor_scale: float = 60.0,
                    out: Optional[Tensor] = None) -> Tensor:
        """Encoder + the glue of ExtractionThread.process (system/modules/odometry.py:46-49):
        (B, Cout+3, S) = [fea ; coor * coor_scale], written by the same call (into `out` if given)."""
        return self._run(points, points_padding, True, coor_scale, out)[3]
