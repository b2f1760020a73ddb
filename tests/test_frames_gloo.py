"""Frame-parallel host logic (deeppointmap_b200/frames.py) on CPU: world_size-2 and -3 gloo
process groups with stub encode / register functions (the CUDA kernels are covered by the -m gpu
tests; here only sharding, the boundary exchange and the pose gather are under test)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deeppointmap_b200 import frames as FP


def test_shard_partitions_everything_contiguously():
    for n in (0, 1, 5, 32, 33):
        for w in (1, 2, 3, 8):
            blocks = [FP.shard(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and (b - a) >= (d - c) >= 0
            assert max(FP.shard_sizes(n, w)) - min(FP.shard_sizes(n, w)) <= 1
    with pytest.raises(ValueError):
        FP.shard(4, 2, 2)


def _encode(points):  # (f,3,N) -> (f,4,2): a descriptor that identifies the frame
    f = points.shape[0]
    fid = points[:, 0, 0]
    return torch.stack([fid, fid * 2, fid * 3, fid * 4], dim=1).unsqueeze(-1).expand(f, 4, 2).contiguous()


def _register(src, dst):  # pose record: [src id, dst id, ...]
    out = torch.zeros(src.shape[0], FP.REG_STRIDE)
    out[:, 0] = src[:, 0, 0]
    out[:, 1] = dst[:, 0, 0]
    out[:, 2] = 1.0
    return out


def _expected(n, with_prev):
    want = torch.zeros(n, FP.REG_STRIDE)
    for i in range(n):
        if i == 0 and not with_prev:
            continue
        want[i, 0] = i - 1 if i > 0 else -1.0
        want[i, 1] = i
        want[i, 2] = 1.0
    return want


def _worker(rank, world, port, n, with_prev, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = FP.shard(n, world, rank)
        pts = torch.zeros(b - a, 3, 7)
        pts[:, 0, 0] = torch.arange(a, b, dtype=torch.float32)
        fp = FP.FrameParallel(_encode, _register)
        prev = torch.full((4, 2), -1.0) * torch.tensor([1.0, 2.0, 3.0, 4.0]).view(4, 1) if with_prev else None
        poses, desc = fp.odometry(pts, n, prev)
        ok = torch.equal(poses, _expected(n, with_prev)) and (desc is None) == (b == a)
        q.put((rank, bool(ok), poses[:, :3].tolist()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n,with_prev", [(2, 8, False), (2, 7, True), (3, 2, True), (2, 1, False)])
def test_frame_parallel_gloo(world, n, with_prev):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, with_prev, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, poses in res:
        assert ok, (rank, poses)


def test_single_process_matches_expected():
    pts = torch.zeros(5, 3, 7)
    pts[:, 0, 0] = torch.arange(5, dtype=torch.float32)
    poses, desc = FP.FrameParallel(_encode, _register).odometry(pts, 5)
    assert torch.equal(poses, _expected(5, False)) and desc.shape == (5, 4, 2)
