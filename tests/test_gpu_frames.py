"""Frame-parallel path on the device (BASELINE configs[3]): `FrameParallel.odometry` with the real
encoder / decoder across 2 ranks -- NCCL with one GPU per rank when the box has two, else two ranks sharing cuda:0
over a gloo group (collectives staged through the host) -- against the same frames run in one process."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
N_PTS, N_FRAMES = 8192, 5


def _frames():
    from deeppointmap_b200 import data
    base = data.kitti_shape_cloud(21, N_PTS)
    out = [base]
    for i in range(1, N_FRAMES):
        out.append(data.rigid_move(base, yaw_deg=0.6 * i, t_m=(0.9 * i, 0.05 * i, 0.0), jitter_m=0.01, seed=40 + i)[0])
    return torch.stack(out).contiguous()


def _models(dev):
    from oracle import model_ref as M   # weights only (tests may use the oracle)
    from deeppointmap_b200 import Encoder, Decoder
    cfg = M.default_config()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ck = os.path.join(root, "oracle", "_ref", "DeepPointMapAAAI.pth")
    if os.path.exists(ck):
        sd = torch.load(ck, map_location="cpu")
        esd, dsd = sd["encoder"], sd["decoder"]
    else:
        esd = M.random_weights(M.encoder_shapes(cfg), seed=1)
        dsd = M.random_weights(M.decoder_shapes(cfg), seed=2)
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    enc.load_state_dict(esd, strict=True)
    dec.load_state_dict(dsd, strict=True)
    return cfg, enc.to(dev), dec.to(dev)


def _worker(rank, world, port, backend, q):
    import torch.distributed as dist
    from deeppointmap_b200 import frames as FP
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dev = torch.device("cuda", rank if backend == "nccl" else 0)
        torch.cuda.set_device(dev)
        if backend == "nccl":
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
        cfg, enc, dec = _models(dev)
        a, b = FP.shard(N_FRAMES, world, rank)
        pts = _frames()[a:b].to(dev)
        fp = FP.FrameParallel(lambda p: enc.descriptors(p, None, coor_scale=cfg.coor_scale),
                              lambda s, d: dec.registration_forward_batch(s, d, 0.5)[0])
        poses, desc = fp.odometry(pts, N_FRAMES)
        poses2, _ = fp.odometry(pts, N_FRAMES, desc_shape=(131, 256))   # known shape: no shape collective
        torch.cuda.synchronize()
        q.put((rank, poses.cpu(), bool(torch.equal(poses, poses2)), None))
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001 -- the parent asserts on it
        import traceback
        q.put((rank, None, False, traceback.format_exc()))


def test_frame_parallel_two_ranks_on_device():
    dev = torch.device("cuda:0")
    cfg, enc, dec = _models(dev)
    pts = _frames().to(dev)
    with torch.no_grad():
        desc = enc.descriptors(pts, None, coor_scale=cfg.coor_scale)
        want, _ = dec.registration_forward_batch(desc[:-1].contiguous(), desc[1:].contiguous(), 0.5)
    want = torch.cat([torch.zeros(1, want.shape[1], device=dev), want]).cpu()
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, backend, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, poses, same, err in res:
        assert err is None, err
        assert same, "desc_shape shortcut changed the result"
        assert poses.shape == want.shape
        # every frame runs through the same kernels whatever the batch it sits in: identical records
        assert torch.equal(poses[:, :13], want[:, :13]), (rank, backend, (poses - want).abs().max())
    # sanity of the chain itself: ~0.9 m steps along x
    assert abs(float(want[1, 9]) - 0.9) < 0.3
