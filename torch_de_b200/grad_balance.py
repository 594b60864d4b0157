"""Loss weighting on top of the jet engine.

`GradNorm` keeps the reference's signature and update rule (src/torch_DE/utils/Loss_weighting.py:30-50):
w_k <- alpha * w_k + (1 - alpha) * min(sum_j |g_j| / |g_k|, max_weight),  g_k = d loss_k / d net.parameters().

The reference obtains every g_k with its own `torch.autograd.grad(loss_k, params, retain_graph=True)` -- k full reverse passes
over the whole batch.  Here the rows produced by this package's `Loss_handler` on a `Jet_engine` output carry where they came
from (the stream block of one stash-keeping forward, and the point range of their group), and the fused path
  1. takes d loss_k / d block from autograd (the fused residual kernel's seed launch -- no network work),
  2. runs ONE C-ABI call per row (`tde_jet_backward_range`) that sweeps only the tiles of that row's own group on the shared
     stash: a boundary row costs a sweep over its few boundary points, not over the batch.  Rows that share a group are swept
     one after the other: the sweep is tensor-bound, so batching their seed sets as extra channels would save stash reads,
     not time,
  3. under data parallelism sums all k gradients with ONE all-reduce of the [k, #params] buffer, so every rank derives the
     same weights (k independent all-reduces otherwise).
Anything else (rows from another engine, recompute mode, tensors without that record) takes the reference's k-pass rule.

`Causal_weighting` mirrors Loss_weighting.py:53-67: weights exp(-eps * cumsum of the time-ordered mean residual), first one 1.
"""
import torch
import torch.distributed as dist


def _fused_gradient_norms(net, losses):
    """[k] gradient norms through range-restricted sweeps, or None when the rows do not all come from one recorded forward."""
    srcs = [getattr(l, '_tde_row', None) for l in losses]
    if not srcs or any(s is None for s in srcs):
        return None
    block = srcs[0][0]
    rec = getattr(block, '_tde_forward', None)
    if rec is None or rec.net is not net or any(s[0] is not block for s in srcs):
        return None
    k = len(losses)
    G = torch.zeros(k, rec.runner.n_params, dtype=torch.float32, device=block.device)
    for i, (loss, (_, rng)) in enumerate(zip(losses, srcs)):
        (seeds,) = torch.autograd.grad(loss, block, torch.ones_like(loss), retain_graph=True, allow_unused=True)
        if seeds is not None:
            rec.gradient(seeds.contiguous(), rng, flat_grad=G[i])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(G)                      # one collective for all k rows
        G.mul_(1.0 / dist.get_world_size())
    leaves = list(net.parameters())
    eff = [t for t in rec.params]
    if all(any(t is p for p in leaves) for t in eff):
        return G.norm(dim=1)                    # plain Linear layers: the kernel's gradient IS the parameter gradient
    # re-parameterised layers (RWF, Fourier features): chain d/d(effective W, b) to the leaves with autograd
    mags = []
    for i in range(k):
        outs = [(t, g) for t, g in zip(eff, rec.runner.split_flat(G[i])) if t.requires_grad]
        gs = torch.autograd.grad([t for t, _ in outs], leaves, [g for _, g in outs], retain_graph=True, allow_unused=True)
        mags.append(torch.sqrt(sum(g.pow(2).sum() for g in gs if g is not None)))
    return torch.stack(mags)


def GradNorm(net: torch.nn.Module, global_weights: torch.Tensor, *losses: torch.Tensor, alpha: float = 0.9,
             max_weight=None, eps: float = 1e-5) -> torch.Tensor:
    assert isinstance(global_weights, torch.Tensor), \
        f'global_weights needs to be of type Torch.Tensor!!! Got {type(global_weights)} type instead'
    assert len(global_weights) == len(losses), \
        f'the number of weights {len(global_weights)} given do not match the number of losses ({len(losses)})'
    cap = float('inf') if max_weight is None else max_weight
    mags = _fused_gradient_norms(net, losses)
    GradNorm.last_path = 'fused' if mags is not None else 'k-pass'
    if mags is None:
        params = list(net.parameters())
        grads = [torch.autograd.grad(loss, params, torch.ones_like(loss), retain_graph=True, allow_unused=True) for loss in losses]
        with torch.no_grad():
            mags = torch.stack([torch.sqrt(sum(g.pow(2).sum() for g in gs if g is not None)) for gs in grads])
    with torch.no_grad():
        total = mags.sum()
        new = torch.where(mags > eps, (total / mags.clamp_min(1e-30)).clamp(max=cap), global_weights.to(mags.dtype))
        return alpha * global_weights + (1 - alpha) * new.to(global_weights.dtype)


GradNorm.last_path = None


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
