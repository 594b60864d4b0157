"""Loss weighting on top of the jet engine.

`GradNorm` keeps the reference's signature and update rule (src/torch_DE/utils/Loss_weighting.py:30-50): one
`torch.autograd.grad(loss_k, params, retain_graph=True, allow_unused=True)` per loss row -- each is one reverse jet sweep on the
stash kept by the forward (the engine supports repeated backward calls on one forward) -- then
w_k <- alpha * w_k + (1 - alpha) * min(sum_j |g_j| / |g_k|, max_weight).

`Causal_weighting` mirrors Loss_weighting.py:53-67: weights exp(-eps * cumsum of the time-ordered mean residual), first one 1.
"""
import torch


def GradNorm(net: torch.nn.Module, global_weights: torch.Tensor, *losses: torch.Tensor, alpha: float = 0.9,
             max_weight=None, eps: float = 1e-5) -> torch.Tensor:
    assert isinstance(global_weights, torch.Tensor), \
        f'global_weights needs to be of type Torch.Tensor!!! Got {type(global_weights)} type instead'
    assert len(global_weights) == len(losses), \
        f'the number of weights {len(global_weights)} given do not match the number of losses ({len(losses)})'
    params = list(net.parameters())
    grads = [torch.autograd.grad(loss, params, torch.ones_like(loss), retain_graph=True, allow_unused=True) for loss in losses]
    cap = float('inf') if max_weight is None else max_weight
    with torch.no_grad():
        mags = torch.stack([torch.sqrt(sum(g.pow(2).sum() for g in gs if g is not None)) for gs in grads])
        total = mags.sum()
        new = torch.where(mags > eps, (total / mags.clamp_min(1e-30)).clamp(max=cap), global_weights.to(mags.dtype))
        return alpha * global_weights + (1 - alpha) * new.to(global_weights.dtype)


def Causal_weighting(loss, eps: float = 1.0) -> torch.Tensor:
    """Causal-training weights for time-ordered residual points (reference: utils/Loss_weighting.py:53-67).

    `loss` is the dict the loss handler returns for point losses: loss['residual'][group][name] -> per-point residual loss,
    points sorted by ascending time.  The rows are averaged, w_i = exp(-eps * sum_{j<=i} mean_loss_j) and w_0 = 1."""
    with torch.no_grad():
        rows = [row for group in loss['residual'].values() for row in group.values()]
        mean = torch.stack(list(rows)).mean(dim=0) if len(rows) > 1 else rows[0]
        weights = torch.exp(-eps * torch.cumsum(mean, dim=0))
        weights[0] = 1
    return weights
