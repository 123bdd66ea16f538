"""world_size-2 gloo test of the N>1 path's host logic: ray sharding + the single flat-bucket gradient all-reduce."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from neuraludf_b200 import dp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.arange(6.0).reshape(2, 3))
    b = torch.nn.Parameter(torch.ones(3))
    frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
    rays = torch.arange(10.0).reshape(5, 2)
    (mine, same) = dp.shard_rays(rays, rays)
    assert torch.equal(mine, same)
    loss = (mine.sum() * w).sum() + (rank + 1) * b.sum()     # per-shard loss
    loss.backward()
    bucket = dp.GradBucket([w, b, frozen])
    assert bucket.flat.numel() == 9
    bucket.allreduce_mean()
    out[rank] = (mine.clone(), w.grad.clone(), b.grad.clone())
    dist.destroy_process_group()


def test_ray_sharding_and_gradient_allreduce_world2():
    from neuraludf_b200 import dp
    assert [dp.shard_bounds(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [dp.shard_bounds(4096, r, 8) for r in range(8)][-1] == (3584, 4096)
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    rays = torch.arange(10.0).reshape(5, 2)
    assert torch.equal(out[0][0], rays[:3]) and torch.equal(out[1][0], rays[3:])
    # mean over ranks of d/dw [(sum of my rays) * sum(w)] and of d/db [(rank+1) * sum(b)]
    expect_w = torch.full((2, 3), float((rays[:3].sum() + rays[3:].sum()) / 2))
    for r in range(world):
        assert torch.allclose(out[r][1], expect_w)
        assert torch.allclose(out[r][2], torch.full((3,), 1.5))
