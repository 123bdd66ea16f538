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


class _FakeHandle:
    """stands in for ops.UdfHandle on the CPU: parameters, the bucket layout the library wants, and a backward that writes
    its gradients straight into the sink views the way ops._UdfFunction.backward does"""

    def __init__(self, params):
        self.ps = params
        self.grad_sink = None

    def sink_layout(self):
        return [[self.ps[2]], [self.ps[0], self.ps[1]]]          # biases first, then the rest

    def backward(self, values):
        sink = self.grad_sink if (self.grad_sink is not None and self.grad_sink.begin()) else None
        outs = []
        for p, v in zip(self.ps, values):
            g = sink.view(p) if sink is not None else torch.empty_like(p)
            g.copy_(v)                                            # "kernel" writes in place
            outs.append(g)
        if sink is not None:
            sink.ready()
        for p, g in zip(self.ps, outs):                           # what autograd's AccumulateGrad does with a fresh gradient
            p.grad = g if p.grad is None else p.grad + g


class _FakeModule:
    def __init__(self, h):
        self._handle = h


def _sink_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from neuraludf_b200 import dp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3))]
    b = [torch.nn.Parameter(torch.zeros(2, 2)), torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(2))]
    loose = torch.nn.Parameter(torch.zeros(1))
    ha, hb = _FakeHandle(a), _FakeHandle(b)
    bucket = dp.GradBucket(a + b + [loose], modules=[_FakeModule(ha), _FakeModule(hb)], overlap=True)
    assert bucket.flat.numel() == 6 + 4 + 3 + 4 + 1 + 2 + 1 and len(bucket.regions) == 2
    res = []
    for step in range(2):                                         # two steps: the regions must re-arm after zero_grad
        for p in a + b + [loose]:
            p.grad = None
        hb.backward([torch.full_like(p, float(rank + 1 + step)) for p in b])     # colour-net-like region first (async)
        ha.backward([torch.full_like(p, float(10 * (rank + 1))) for p in a])
        loose.grad = torch.full((1,), float(rank))
        # gradients must live inside the flat bucket (no pack / unpack copies)
        assert a[0].grad.data_ptr() >= bucket.flat.data_ptr()
        assert a[0].grad.data_ptr() < bucket.flat.data_ptr() + 4 * bucket.flat.numel()
        bucket.allreduce_mean()
        res.append((a[0].grad.clone(), b[2].grad.clone(), loose.grad.clone()))
    out[rank] = res
    dist.destroy_process_group()


def test_gradient_sinks_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sink_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        for step in range(2):
            ga, gb, gl = out[r][step]
            assert torch.allclose(ga, torch.full((3, 2), 15.0))                  # mean of 10 and 20
            assert torch.allclose(gb, torch.full((2,), 1.5 + step))              # mean of (1 + step) and (2 + step)
            assert torch.allclose(gl, torch.full((1,), 0.5))


def _replay_worker(rank, world, port, out):
    """the CUDA-graph arrangement of bench.py at N > 1: forward + backward are replayed (the kernels write the flat bucket and the
    static loose gradients, no Python bookkeeping runs), then allreduce_replayed() + the optimiser follow eagerly"""
    sys.path.insert(0, ROOT)
    from neuraludf_b200 import dp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3))]
    loose = torch.nn.Parameter(torch.zeros(2))
    ha = _FakeHandle(a)
    bucket = dp.GradBucket(a + [loose], modules=[_FakeModule(ha)], overlap=False)
    # "capture": one ordinary step establishes the static gradient tensors (views of the flat bucket + the loose gradient)
    ha.backward([torch.full_like(p, 1.0) for p in a])
    loose.grad = torch.zeros(2)
    static = [p.grad for p in a] + [loose.grad]
    res = []
    for step in range(2):
        # "replay": the kernels overwrite the static tensors in place; begin() / ready() do not run
        for g in static[:-1]:
            g.fill_(float((rank + 1) * (step + 1)))
        static[-1].fill_(float(rank + 10 * step))
        bucket.allreduce_replayed()
        assert all(p.grad is g for p, g in zip(a + [loose], static))             # nothing was re-allocated or zeroed
        res.append((a[0].grad.clone(), a[2].grad.clone(), loose.grad.clone()))
    out[rank] = res
    dist.destroy_process_group()


def test_allreduce_after_graph_replay_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_replay_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        for step in range(2):
            ga, gc, gl = out[r][step]
            assert torch.allclose(ga, torch.full((3, 2), 1.5 * (step + 1)))      # mean of (step+1) and 2 (step+1)
            assert torch.allclose(gc, torch.full((3,), 1.5 * (step + 1)))
            assert torch.allclose(gl, torch.full((2,), 0.5 + 10.0 * step))       # mean of 10 step and 1 + 10 step
