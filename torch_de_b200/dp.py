"""Data parallelism over collocation points (the reference has none, ReadMe.md:155).

Points are independent, so every rank holds a shard of each group and the weights are replicated.  The
only exchange is ONE all-reduce of the flat parameter gradient per step.  `FlatGrads` makes every
`p.grad` a view of one contiguous FP32 buffer, so autograd accumulates straight into it and the
collective needs no flatten/unflatten copies.
"""
from typing import Iterable, List, Sequence

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
        """Use this instead of `optimizer.zero_grad()`: the default `set_to_none=True` drops the views, autograd then
        allocates fresh gradients and the flat buffer is no longer what the parameters accumulate into."""
        self.flat.zero_()

    def _rebind(self):
        """Re-install the views (after an `optimizer.zero_grad(set_to_none=True)` by the caller), keeping whatever gradients
        the parameters hold at the moment."""
        o = 0
        for p in self.params:
            view = self.flat[o:o + p.numel()].view_as(p)
            if p.grad is None:
                view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view
            o += p.numel()

    def all_reduce_mean(self, group=None):
        """Sum over ranks, divide by world size: with equal shards the mean of local means is the global mean
        (loss.py:26,57 `mean` over the global point count)."""
        world = dist.get_world_size(group)
        if not self.check_aliasing():       # somebody reset the gradients to None: reducing a stale buffer would go unnoticed
            self._rebind()
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
    """Contiguous 1/world slice of a point set.  Shards are EQUAL (the mean of the ranks' means is then the global mean): the
    `x.shape[0] % world` points at the tail are dropped, like the reference's sampler drops a ragged last batch
    (utils/data.py:233-243)."""
    n = x.shape[0] // world
    return x[rank * n:(rank + 1) * n]


def proportional_shards(step_ms: Sequence[float], sizes: Sequence[int], quantum: int = 32) -> List[int]:
    """Shard sizes proportional to each rank's measured speed (points per millisecond), same total.

    Synchronous data parallelism ends a step when the SLOWEST rank is ready; GPUs under their power cap settle at different
    clocks (1.66-1.73 GHz on one 8 x B200 box), so equal shards leave the faster ones waiting in the all-reduce.  With shards
    n_r proportional to n_r / t_r every rank needs the same time.  The loss is a mean over ALL points, so a rank's residual
    rows then weigh n_r / N instead of 1 / world: scale them by world * n_r / N before `all_reduce_mean`.
    Sizes are rounded to `quantum` points (whole tiles); the rounding remainder goes to the fastest rank.
    Host-side helper only: bench.py keeps equal shards (measured on 8 GPUs, 20 steps: mean compute 43.1 ms per step, slowest
    rank's mean 44.2 ms, mean of the per-step maximum 45.8 ms -- 1.1 ms of the 2.7 ms that synchronous DP pays is rank skew,
    which this addresses; the other 1.6 ms is step-to-step jitter of power-capped clocks, which it cannot)."""
    assert len(step_ms) == len(sizes) and all(t > 0 for t in step_ms)
    total = int(sum(sizes))
    speed = [n / t for n, t in zip(sizes, step_ms)]
    tot_speed = sum(speed)
    out = [max(quantum, int(total * v / tot_speed) // quantum * quantum) for v in speed]
    out[max(range(len(out)), key=lambda r: speed[r])] += total - sum(out)
    return out
