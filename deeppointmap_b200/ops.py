"""Tensor-level front-ends of the index ops, with pytorch3d.ops-compatible signatures.

These are what the drop-in `pytorch3d` package (deeppointmap_b200/compat/pytorch3d) exports,
so the unmodified reference's `-t3d` branches (network/encoder/utils.py:11-14, 91-123, 273-285)
run on libdpm_b200.so.  CUDA tensors only; no CPU fallback.
"""
from collections import namedtuple
from typing import Optional, Union

import numpy as np
import torch

from . import _C

_KNN = namedtuple("KNN", "dists idx knn")
_BallQuery = namedtuple("BallQuery", "dists idx knn")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _lengths(lengths, B: int, P: int, device):
    if lengths is None:
        return None
    if lengths.shape != (B,):
        raise ValueError("points and lengths must have same batch dimension.")
    return lengths.to(device=device, dtype=torch.int64).contiguous()


def _ws(device, nbytes):
    return _C.workspaces.get(device, nbytes, f"ops{_C.stream_ptr(device)}")


def set_fps_mode(mode: int = 0) -> None:
    """How every FPS of the library maps a cloud onto the chip (`dpm_set_fps_mode`): 0 auto (a cluster of 8 SMs per
    cloud while the batch is <= 16 clouds, else one SM per cloud), 1 always one SM, 2 always a cluster, 3 "packed": two
    clouds per SM (less SM time, longer latency: for >= 8 streams of 32-frame batches in flight).  Same picks."""
    _C.lib().dpm_set_fps_mode(int(mode))


def sample_farthest_points(points: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                           K: Union[int, list, torch.Tensor] = 50, random_start_point: bool = False):
    """pytorch3d.ops.sample_farthest_points: (N,P,D) -> (sampled (N,K,D), idx (N,K) int64, -1 padded)."""
    _C.require_cuda(points)
    if random_start_point:
        raise NotImplementedError("random_start_point=True is not on the DeepPointMap path (utils.py:273)")
    if not isinstance(K, int):
        ks = torch.as_tensor(K).flatten().tolist()
        if len(set(ks)) != 1:
            raise NotImplementedError("per-cloud K is not supported")
        K = int(ks[0])
    p = _f32c(points)
    B, N, D = p.shape
    if D < 3:
        raise ValueError("points must have at least 3 columns")
    L = _lengths(lengths, B, N, p.device)
    idx = torch.empty((B, K), dtype=torch.int64, device=p.device)
    out = torch.empty((B, K, D), dtype=torch.float32, device=p.device)
    lib = _C.lib()
    nb = lib.dpm_fps_workspace_bytes(B, N, D, K)
    ws = _ws(p.device, nb)
    with torch.cuda.device(p.device):
        _C.check(lib.dpm_fps_f32(p.data_ptr(), B, N, D, _C.ptr(L), K, idx.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                 ws.numel(), _C.stream_ptr()), "fps")
    return out, idx


def knn_gather(x: torch.Tensor, idx: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pytorch3d.ops.knn_gather: x (N,M,U), idx (N,L,K) -> (N,L,K,U)."""
    N, M, U = x.shape
    _, L, K = idx.shape
    out = x[:, :, None].expand(-1, -1, K, -1).gather(1, idx[:, :, :, None].expand(-1, -1, -1, U))
    if lengths is not None:
        mask = torch.arange(K, device=x.device)[None].expand(N, -1) >= lengths[:, None]
        out = out.masked_fill(mask[:, None, :, None].expand(-1, L, -1, U), 0.0)
    return out


def knn_points(p1: torch.Tensor, p2: torch.Tensor, lengths1=None, lengths2=None, norm: int = 2, K: int = 1,
               version: int = -1, return_nn: bool = False, return_sorted: bool = True):
    """pytorch3d.ops.knn_points: squared L2, ascending by (d2, index); idx int64; zero padded."""
    _C.require_cuda(p1, p2)
    if norm != 2:
        raise NotImplementedError("only the L2 norm is on the DeepPointMap path")
    a, b = _f32c(p1), _f32c(p2)
    if a.shape[0] != b.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    B, S, D1 = a.shape
    _, N, D2 = b.shape
    L1, L2 = _lengths(lengths1, B, S, a.device), _lengths(lengths2, B, N, a.device)
    idx = torch.empty((B, S, K), dtype=torch.int64, device=a.device)
    d2 = torch.empty((B, S, K), dtype=torch.float32, device=a.device)
    lib = _C.lib()
    ws = _ws(a.device, lib.dpm_knn_workspace_bytes(B, S, N, K))
    with torch.cuda.device(a.device):
        _C.check(lib.dpm_knn_f32(a.data_ptr(), D1, b.data_ptr(), D2, B, S, N, _C.ptr(L1), _C.ptr(L2), K,
                                 idx.data_ptr(), d2.data_ptr(), ws.data_ptr(), ws.numel(), _C.stream_ptr()), "knn")
    nn = knn_gather(p2, idx, lengths2) if return_nn else None
    return _KNN(dists=d2, idx=idx, knn=nn)


def ball_query(p1: torch.Tensor, p2: torch.Tensor, lengths1=None, lengths2=None, K: int = 500, radius: float = 0.2,
               return_nn: bool = True):
    """pytorch3d.ops.ball_query: first K points (index order) with d2 < radius**2; -1 / 0 padded."""
    _C.require_cuda(p1, p2)
    a, b = _f32c(p1), _f32c(p2)
    B, S, D1 = a.shape
    _, N, D2 = b.shape
    L1, L2 = _lengths(lengths1, B, S, a.device), _lengths(lengths2, B, N, a.device)
    idx = torch.empty((B, S, K), dtype=torch.int64, device=a.device)
    d2 = torch.empty((B, S, K), dtype=torch.float32, device=a.device)
    lib = _C.lib()
    ws = _ws(a.device, lib.dpm_knn_workspace_bytes(B, S, N, K))
    r2 = float(np.float32(float(radius) ** 2))
    with torch.cuda.device(a.device):
        _C.check(lib.dpm_ball_query_f32(a.data_ptr(), D1, b.data_ptr(), D2, B, S, N, _C.ptr(L1), _C.ptr(L2), K, r2,
                                        idx.data_ptr(), d2.data_ptr(), ws.data_ptr(), ws.numel(), _C.stream_ptr()),
                 "ball_query")
    nn = None
    if return_nn:
        nn = knn_gather(p2, idx.clamp(min=0), None)
        nn = nn.masked_fill((idx < 0)[..., None], 0.0)
    return _BallQuery(dists=d2, idx=idx, knn=nn)


def hybrid_query(radius: float, K: int, points: torch.Tensor, centers: torch.Tensor,
                 points_padding: torch.Tensor) -> torch.Tensor:
    """Querier.hybrid_query_t3d (network/encoder/utils.py:112-123) as one fused call -> idx (B,S,K) int64."""
    _C.require_cuda(points, centers)
    a, b = _f32c(centers), _f32c(points)
    B, S, D1 = a.shape
    _, N, D2 = b.shape
    L2 = (~points_padding).sum(1).to(torch.int64).contiguous()
    idx = torch.empty((B, S, K), dtype=torch.int64, device=a.device)
    lib = _C.lib()
    ws = _ws(a.device, lib.dpm_knn_workspace_bytes(B, S, N, K))
    r2 = float(np.float32(float(radius) ** 2))
    with torch.cuda.device(a.device):
        _C.check(lib.dpm_knn_radius_f32(a.data_ptr(), D1, b.data_ptr(), D2, B, S, N, L2.data_ptr(), K, r2,
                                        idx.data_ptr(), ws.data_ptr(), ws.numel(), _C.stream_ptr()), "knn_radius")
    return idx


def information_matrix(pointcloud_1: torch.Tensor, pointcloud_2: torch.Tensor, SE3: torch.Tensor, radius: float = 1.0,
                       return_count: bool = False):
    """calculate_information_matrix_from_pcd (system/modules/utils.py:60-104, pytorch3d branch) as one native
    call: (3,N1), (3,N2) CUDA fp32 clouds and a (4,4) pose -> (6,6) information matrix on the same device
    (no host sync; the reference's caller does `.cpu()` itself).  SURVEY.md section 8f rank 2."""
    _C.require_cuda(pointcloud_1, pointcloud_2)
    if pointcloud_1.dim() != 2 or pointcloud_2.dim() != 2 or pointcloud_1.shape[0] != 3 or pointcloud_2.shape[0] != 3:
        raise ValueError("point clouds must be (3, N)")
    if tuple(SE3.shape) != (4, 4):
        raise ValueError("SE3 must be (4, 4)")
    p1, p2 = _f32c(pointcloud_1), _f32c(pointcloud_2)
    dev = p1.device
    T = SE3.to(device=dev, dtype=torch.float32).contiguous()
    n1, n2 = p1.shape[1], p2.shape[1]
    info = torch.empty((6, 6), dtype=torch.float32, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    if n1 == 0 or n2 == 0:
        info.zero_()
        cnt.zero_()
        return (info, cnt) if return_count else info
    lib = _C.lib()
    nb = lib.dpm_information_matrix_workspace_bytes(n1, n2)
    ws = _ws(dev, nb)
    with torch.cuda.device(dev):
        _C.check(lib.dpm_information_matrix_f32(p1.data_ptr(), n1, p2.data_ptr(), n2, T.data_ptr(), float(radius),
                                                info.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(),
                                                _C.stream_ptr()), "information_matrix")
    return (info, cnt) if return_count else info


def preprocess_frame(raw: torch.Tensor, voxel_size: float = 0.3, min_dis: float = 1.0, max_dis: float = 60.0,
                     ratio: float = 60.0, max_voxels: int = 1 << 25, outlier=None, lowpass=None) -> torch.Tensor:
    """Raw frame -> encoder input on the device: BinReader's NaN-row drop, VoxelSample(voxel_size, 'first'),
    DistanceSample(min_dis, max_dis), [OutlierFilter(*outlier), e.g. outlier=(10, 3.0) as in the shipped YAML],
    CoordinatesNormalization(ratio) (dataloader/heads/bin.py:16-17, dataloader/transforms.py:230-246, 331-356,
    387-407).  raw (N, C>=3) CUDA fp32 rows (a KITTI .bin is (N,4)) -> (3, n) fp32, points in the reference's
    order (ascending voxel id).  One host sync per data-dependent size.  max_voxels bounds the dense first-index table
    (4 bytes per cell of the frame's bounding grid, 134 MB of scratch at the default: a 160 x 160 x 30 m frame at 0.3 m
    has 28 M cells); a frame whose grid is larger raises and asks for a higher bound."""
    if outlier is not None or lowpass is not None:
        # the whole shipped YAML chain on the device: VoxelSample -> DistanceSample -> [OutlierFilter(*outlier)] ->
        # [LowPassFilter(*lowpass), e.g. lowpass=(0.5, 16, 2.0, 4)] -> CoordinatesNormalization
        rows = preprocess_frame(raw, voxel_size, min_dis, max_dis, 1.0, max_voxels).T.contiguous()   # metres (x / 1.0 is exact)
        # pcd.xyz /= ratio as an IEEE division inside the LAST filter's emit pass (torch's CUDA `/ scalar` multiplies
        # by the reciprocal, which is 1 ulp off the reference's CPU arithmetic)
        if outlier is not None:
            rows = outlier_filter(rows, int(outlier[0]), float(outlier[1]), out_divisor=1.0 if lowpass is not None else float(ratio))
        if lowpass is not None:
            rows = low_pass_filter(rows, *lowpass, out_divisor=float(ratio))
        return rows.T.contiguous()
    _C.require_cuda(raw)
    if raw.dim() != 2 or raw.shape[1] < 3:
        raise ValueError("raw must be (N, C>=3)")
    r = _f32c(raw)
    n, stride = r.shape
    dev = r.device
    if n == 0:
        return torch.empty((3, 0), dtype=torch.float32, device=dev)
    out = torch.empty((n, 3), dtype=torch.float32, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    lib = _C.lib()
    nb = lib.dpm_frontend_workspace_bytes(int(max_voxels))
    if nb == 0:
        raise ValueError("max_voxels out of range")
    ws = _ws(dev, nb)
    with torch.cuda.device(dev):
        _C.check(lib.dpm_frontend_f32(r.data_ptr(), n, stride, float(voxel_size), float(min_dis), float(max_dis),
                                      float(ratio), int(max_voxels), out.data_ptr(), cnt.data_ptr(), ws.data_ptr(),
                                      ws.numel(), _C.stream_ptr()), "frontend")
    k = int(cnt.item())
    if k < 0:
        raise ValueError(f"the voxel grid of this frame exceeds max_voxels={max_voxels}; crop the frame or raise it")
    return out[:k].T.contiguous()


def map_tile(store: torch.Tensor, ids, poses: torch.Tensor, center: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Scan-to-map input stage on the device (PoseGraph.__global_mapping + the centring of
    global_map_query_graph, system/modules/pose_graph.py:373-409, 499-511).  store (n, Cd, S) device-resident
    descriptor sets, ids (m,) key-frame slots, poses (m,4,4) their SE3_pred, center (4,4) or None ->
    (Cd, m*S) map tile = the `dst_descriptor` of Decoder.registration_forward.  SURVEY.md section 8f rank 3."""
    _C.require_cuda(store)
    if store.dim() != 3 or store.shape[1] < 4:
        raise ValueError("store must be (n, Cd>=4, S)")
    st = _f32c(store)
    dev = st.device
    idt = torch.as_tensor(ids, dtype=torch.int32).to(dev).contiguous().view(-1)
    m = int(idt.numel())
    P = poses.to(device=dev, dtype=torch.float32).contiguous().view(-1, 16)
    if m == 0 or P.shape[0] != m:
        raise ValueError("ids and poses must have the same non-zero length")
    C = None if center is None else center.to(device=dev, dtype=torch.float32).contiguous().view(16)
    n, Cd, S = st.shape
    tile = torch.empty((Cd, m * S), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _C.check(_C.lib().dpm_map_tile_f32(st.data_ptr(), n, Cd, S, idt.data_ptr(), P.data_ptr(), _C.ptr(C), m,
                                           tile.data_ptr(), _C.stream_ptr()), "map_tile")
    return tile


def outlier_filter(rows: torch.Tensor, nb_neighbors: int = 10, std_ratio: float = 3.0, return_mask: bool = False,
                   out_divisor: float = 1.0):
    """OutlierFilter, the reference's CUDA branch (dataloader/transforms.py:230-246), in one native call:
    rows (N, C>=3) CUDA fp32 -> (n, 3) rows kept (original order, IEEE-divided by out_divisor)
    [, (N,) bool mask].  One host sync (n)."""
    _C.require_cuda(rows)
    if rows.dim() != 2 or rows.shape[1] < 3:
        raise ValueError("rows must be (N, C>=3)")
    r = _f32c(rows)
    n, stride = r.shape
    dev = r.device
    if n == 0:
        e = torch.empty((0, 3), dtype=torch.float32, device=dev)
        return (e, torch.empty((0,), dtype=torch.bool, device=dev)) if return_mask else e
    out = torch.empty((n, 3), dtype=torch.float32, device=dev)
    mask = torch.empty((n,), dtype=torch.bool, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    lib = _C.lib()
    nb = lib.dpm_outlier_filter_workspace_bytes(n, int(nb_neighbors))
    if nb == 0:
        raise NotImplementedError("nb_neighbors must be in 1..31")
    ws = _ws(dev, nb)
    with torch.cuda.device(dev):
        _C.check(lib.dpm_outlier_filter_f32(r.data_ptr(), n, stride, int(nb_neighbors), float(std_ratio), float(out_divisor),
                                            out.data_ptr(),
                                            mask.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(), _C.stream_ptr()),
                 "outlier_filter")
    kept = out[:int(cnt.item())]
    return (kept, mask) if return_mask else kept


def low_pass_filter(rows: torch.Tensor, normals_radius: float = 0.5, normals_num: int = 16, filter_std: float = 2.0,
                    flux: int = 4, max_remain: int = -1, return_mask: bool = False, out_divisor: float = 1.0):
    """LowPassFilter (dataloader/transforms.py:256-297) in one native call: rows (N, C>=3) CUDA fp32, metres ->
    (n, 3) rows kept (original order, IEEE-divided by out_divisor) [, (N,) bool mask].  One host sync (n).
    max_remain > 0 and smaller than the kept count: the reference then keeps the max_remain points of largest
    similarity IN THAT ORDER (transforms.py:284-285) -- done here from the kernel's statistic."""
    _C.require_cuda(rows)
    if rows.dim() != 2 or rows.shape[1] < 3:
        raise ValueError("rows must be (N, C>=3)")
    r = _f32c(rows)
    n, stride = r.shape
    dev = r.device
    if n == 0:
        e = torch.empty((0, 3), dtype=torch.float32, device=dev)
        return (e, torch.empty((0,), dtype=torch.bool, device=dev)) if return_mask else e
    out = torch.empty((n, 3), dtype=torch.float32, device=dev)
    mask = torch.empty((n,), dtype=torch.bool, device=dev)
    sim = torch.empty((n,), dtype=torch.float32, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    lib = _C.lib()
    nb = lib.dpm_low_pass_filter_workspace_bytes(n, int(normals_num))
    if nb == 0:
        raise NotImplementedError("normals_num must be in 1..31")
    ws = _ws(dev, nb)
    with torch.cuda.device(dev):
        _C.check(lib.dpm_low_pass_filter_f32(r.data_ptr(), n, stride, float(normals_radius), int(normals_num),
                                             float(filter_std), int(flux), float(out_divisor), out.data_ptr(),
                                             mask.data_ptr(), sim.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(),
                                             _C.stream_ptr()), "low_pass_filter")
    k = int(cnt.item())
    if 0 < int(max_remain) < k:
        top = torch.topk(sim, k=int(max_remain)).indices
        kept = r[top, :3] / float(out_divisor)
        if return_mask:
            m = torch.zeros((n,), dtype=torch.bool, device=dev)
            m[top] = True
            return kept, m
        return kept
    kept = out[:k]
    return (kept, mask) if return_mask else kept
