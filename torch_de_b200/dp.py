"""Data parallelism over collocation points (the reference has none, ReadMe.md:155).

Points are independent, so every rank holds a shard of each group and the weights are replicated.  The
only exchange is ONE all-reduce of the flat parameter gradient per step.  `FlatGrads` makes every
`p.grad` a view of one contiguous FP32 buffer, so autograd accumulates straight into it and the
collective needs no flatten/unflatten copies.
"""
from typing import Iterable

import torch
import torch.distributed as dist


class FlatGrads:
    def __init__(self, net_or_params):
        params = list(net_or_params.parameters()) if hasattr(net_or_params, 'parameters') else list(net_or_params)
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=self.params[0].dtype, device=self.params[0].device)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        """Sum over ranks, divide by world size: with equal shards the mean of local means is the global mean
        (loss.py:26,57 `mean` over the global point count)."""
        world = dist.get_world_size(group)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if world > 1:
            self.flat.mul_(1.0 / world)

    def check_aliasing(self) -> bool:
        base = self.flat.data_ptr()
        o = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != base + o * self.flat.element_size():
                return False
            o += p.numel()
        return True


def shard(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous 1/world slice of a point set (equal sizes; the tail remainder is dropped like the reference's
    sampler drops a ragged last batch, utils/data.py:233-243)."""
    n = x.shape[0] // world
    return x[rank * n:(rank + 1) * n]
