"""Data-parallel plumbing for the ray-sharded renderer: rays are independent, parameters are replicated,
so the only exchange per training step is ONE all-reduce of the flat gradient bucket (NCCL over NVLink/NVSwitch on the
GPU box; the same code runs on gloo/CPU tensors in the unit tests).  No kernels here."""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous split of n_items over `world` ranks (first ranks get the remainder): [lo, hi)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays_o, rays_d, *more, rank=None, world=None):
    """This rank's contiguous slice of every per-ray tensor."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(rays_o.shape[0], rank, world)
    return tuple(t[lo:hi] for t in (rays_o, rays_d) + more)


class GradBucket:
    """Flat fp32 bucket over the trainable parameters; `allreduce_mean()` sums the per-rank gradients with one
    collective and writes the mean back into each `.grad` (DDP semantics: mean of the per-shard losses)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)

    def allreduce_mean(self, group=None):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / world)
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
