"""Frame-parallel multi-GPU driver for the hot path.

The reference has no multi-GPU inference (pipeline/infer.py:7 asserts DDP off); frames are
independent units -- Encoder.forward has no cross-batch interaction and registration_forward is
per (src, dst) pair -- so a batch of F frames is sharded in contiguous blocks over the ranks of a
`torch.distributed` group (NCCL over NVLink on the B200 box, gloo in the CPU tests):

  1. every rank encodes its block                               (no communication)
  2. ONE all-gather of each rank's LAST descriptor set (134 KB) so rank r can register its first
     frame against frame (start_r - 1), which lives on rank r-1
  3. every rank registers its consecutive pairs                 (no communication)
  4. ONE all-gather of the pose records (64 B per frame) so every rank holds the trajectory

Both collectives are latency-sized (<= 1 MB at 8 ranks); there is nothing to overlap or fuse.
"""
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

REG_STRIDE = 16  # dpm_b200.h DPM_REG_STRIDE


def shard(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rank `rank`: sizes differ by at most one, earlier ranks
    take the larger blocks; ranks beyond n_frames get an empty block."""
    if n_frames < 0 or world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad shard request n_frames={n_frames} world={world} rank={rank}")
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n_frames: int, world: int) -> List[int]:
    return [b - a for a, b in (shard(n_frames, world, r) for r in range(world))]


class FrameParallel:
    """encode_fn(points (f,3,N)) -> descriptors (f,Cd,S); register_fn(src (p,Cd,S), dst (p,Cd,S))
    -> pose records (p, REG_STRIDE).  On the GPU box these are `Encoder.descriptors` and
    `Decoder.registration_forward_batch(...)[0]`."""

    def __init__(self, encode_fn: Callable[[Tensor], Tensor], register_fn: Callable[[Tensor, Tensor], Tensor],
                 group: Optional[dist.ProcessGroup] = None):
        self.encode_fn, self.register_fn, self.group = encode_fn, register_fn, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def _ag(self, out: Tensor, x: Tensor) -> None:
        """all_gather_into_tensor; a gloo group (CPU tests, or several ranks sharing one GPU) is fed through host
        copies, NCCL takes the device tensors as they are"""
        if x.is_cuda and dist.get_backend(self.group) == "gloo":
            h = torch.empty(out.shape, dtype=out.dtype)
            dist.all_gather_into_tensor(h, x.cpu(), group=self.group)
            out.copy_(h)
        else:
            dist.all_gather_into_tensor(out, x, group=self.group)

    def _all_gather_padded(self, x: Tensor, sizes: List[int]) -> Tensor:
        """all-gather of per-rank blocks with different leading sizes (padded to the largest)."""
        if self.world == 1:
            return x
        m = max(sizes)
        buf = x.new_zeros((m,) + tuple(x.shape[1:]))
        buf[: x.shape[0]] = x
        out = x.new_empty((self.world * m,) + tuple(x.shape[1:]))
        self._ag(out, buf)
        if len(set(sizes)) == 1:
            return out
        return torch.cat([out[r * m: r * m + s] for r, s in enumerate(sizes)], dim=0)

    @torch.no_grad()
    def odometry(self, local_points: Tensor, n_frames: int, prev_desc: Optional[Tensor] = None,
                 desc_shape: Optional[Tuple[int, int]] = None):
        """local_points: this rank's block (shard(n_frames, world, rank)) of the global batch.
        Returns (poses (n_frames, REG_STRIDE) for the pairs (i-1 -> i), i = 0 pairs with
        `prev_desc` (the last frame of the previous batch) or is zero-filled; local descriptors).
        desc_shape = (Cd, S) of a descriptor set when the caller knows it: skips the one collective (and host
        sync) that otherwise lets ranks without frames learn it."""
        start, stop = shard(n_frames, self.world, self.rank)
        f = stop - start
        if local_points.shape[0] != f:
            raise ValueError(f"rank {self.rank} owns frames [{start},{stop}) but got {local_points.shape[0]} frames")
        sizes = shard_sizes(n_frames, self.world)
        desc = self.encode_fn(local_points) if f > 0 else None
        # boundary exchange: last descriptor of every rank
        if self.world > 1:
            # every rank must contribute the same shape: learn it from whoever has frames
            if desc_shape is not None:
                cd, s = int(desc_shape[0]), int(desc_shape[1])
            else:
                if desc is not None:
                    shape = torch.tensor(list(desc.shape[1:]), dtype=torch.int64)
                else:
                    shape = torch.zeros(2, dtype=torch.int64)
                if dist.get_backend(self.group) != "gloo":
                    shape = shape.to(local_points.device)
                dist.all_reduce(shape, op=dist.ReduceOp.MAX, group=self.group)
                cd, s = int(shape[0]), int(shape[1])
            last = desc[-1:] if desc is not None else torch.zeros((1, cd, s), device=local_points.device)
            lasts = torch.empty((self.world, cd, s), dtype=last.dtype, device=last.device)
            self._ag(lasts, last.contiguous())
        else:
            lasts = None
        poses = torch.zeros((f, REG_STRIDE), dtype=torch.float32, device=local_points.device)
        if f > 0:
            # predecessor of the block's first frame
            pred = None
            if start > 0:
                owner = max(r for r in range(self.world) if sizes[r] > 0 and shard(n_frames, self.world, r)[1] <= start)
                pred = lasts[owner: owner + 1]
            elif prev_desc is not None:
                pred = prev_desc.unsqueeze(0) if prev_desc.dim() == 2 else prev_desc
            if pred is not None:
                src = torch.cat([pred, desc[:-1]], dim=0)
                poses = self.register_fn(src.contiguous(), desc)
            elif f > 1:
                poses[1:] = self.register_fn(desc[:-1].contiguous(), desc[1:].contiguous())
        all_poses = self._all_gather_padded(poses, sizes)
        return all_poses, desc
