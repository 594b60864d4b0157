"""Network zoo subset on the hot path + introspection of a user's nn.Module into Linear layers.

`MLP` mirrors the reference constructor and layer placement (src/torch_DE/continuous/Networks/Networks.py:123-152):
`MLP(in, out, hidden, num_hidden_layers, activation)` has num_hidden_layers+1 activated Linear layers then
`linear_out`.  The jet kernels read the nn.Linear parameters in place on every call.
"""
import math
from typing import List, Tuple

import torch
import torch.nn as nn

from . import _lib


def get_activation_function(activation):
    return getattr(torch, activation) if isinstance(activation, str) else activation


class RWF_Linear(nn.Module):
    """Random-weight-factorised Linear, W = v * s with a per-output-row scale (reference Networks.py:25-50).
    Parameter registration order (bias, s, v) matches the reference so state dicts are interchangeable."""

    def __init__(self, in_features: int, out_features: int, mean: float = 1, std_dev: float = 0.1, bias: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        w = torch.empty(out_features, in_features)
        nn.init.xavier_normal_(w)
        if bias:
            bound = 1 / math.sqrt(in_features) if in_features > 0 else 0
            self.bias = nn.Parameter(torch.empty(out_features).uniform_(-bound, bound))
        else:
            self.register_parameter('bias', None)
        self.s = nn.Parameter(torch.exp(torch.normal(mean, std=torch.ones(out_features) * std_dev)).unsqueeze(-1))
        self.v = nn.Parameter(w / self.s.detach())

    def forward(self, x):
        return nn.functional.linear(x, self.v * self.s, self.bias)


class Fourier_Embedding(nn.Module):
    """x -> cat[sin(Wx+b), cos(Wx+b)] with W ~ N(0, std) (reference Networks.py:53-75)."""

    def __init__(self, in_features: int, out_features: int, std_dev: float = 3., bias: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features).normal_(0, std_dev))
        if bias:
            bound = 1 / math.sqrt(in_features) if in_features > 0 else 0
            self.bias = nn.Parameter(torch.empty(out_features).uniform_(-bound, bound))
        else:
            self.register_parameter('bias', None)

    def forward(self, x):
        x = nn.functional.linear(x, self.weight, self.bias)
        return torch.cat([torch.sin(x), torch.cos(x)], dim=-1)


class Fourier_Net(nn.Module):
    """Fourier embedding + (RWF) Linear stack, every layer followed by the activation (reference Networks.py:104-120)."""

    def __init__(self, in_features: int, out_features: int, hidden_features: int, num_hidden_layers: int,
                 activation='tanh', RWF: bool = False, fourier_std=3):
        super().__init__()
        linear = RWF_Linear if RWF else nn.Linear
        self.linear_in = Fourier_Embedding(in_features, hidden_features // 2, std_dev=fourier_std, bias=False)
        self.linear_out = linear(hidden_features, out_features)
        self.activation = get_activation_function(activation)
        self.layers = nn.ModuleList([self.linear_in] + [linear(hidden_features, hidden_features) for _ in range(num_hidden_layers)])

    def forward(self, x):
        for layer in self.layers:
            x = self.activation(layer(x))
        return self.linear_out(x)


class MLP(nn.Module):
    def __init__(self, in_features: int, out_features: int, hidden_features: int, num_hidden_layers: int,
                 activation='tanh', RWF: bool = False):
        super().__init__()
        linear = RWF_Linear if RWF else nn.Linear
        # construction order matters for seeded parity with the reference: linear_in, linear_out, hidden
        self.linear_in = linear(in_features, hidden_features)
        self.linear_out = linear(hidden_features, out_features)
        self.activation = get_activation_function(activation)
        self.layers = nn.ModuleList([self.linear_in] +
                                    [linear(hidden_features, hidden_features) for _ in range(num_hidden_layers)])

    def forward(self, x):
        for layer in self.layers:
            x = self.activation(layer(x))
        return self.linear_out(x)


class Sin(nn.Module):
    def forward(self, x):
        return torch.sin(x)


def _act_code(fn) -> int:
    if fn in (torch.tanh, torch.nn.functional.tanh) or isinstance(fn, nn.Tanh) or getattr(fn, '__name__', '') == 'tanh':
        return _lib.ACT_TANH
    if fn is torch.sin or isinstance(fn, Sin) or getattr(fn, '__name__', '') == 'sin':
        return _lib.ACT_SIN
    raise NotImplementedError(f"activation {fn!r} has no jet kernel (tanh and sin are implemented); no fallback exists")


class LayerSpec:
    """One Linear-like layer as the kernels see it: shapes plus a recipe for its (weight[out,in], bias[out]) tensors.
    The tensors are rebuilt on every call from the module's leaf parameters with ordinary torch ops, so autograd
    carries the kernel's d/dW back to RWF's (v, s) or to the shared Fourier frequencies without any kernel support."""

    def __init__(self, module: nn.Module):
        self.module = module
        self.fourier = False
        if type(module) is nn.Linear:
            self.kind = 'linear'
            self.in_features, self.out_features = module.in_features, module.out_features
        elif all(hasattr(module, a) for a in ('v', 's', 'in_features', 'out_features')):
            self.kind = 'rwf'                                   # RWF_Linear (this package's or the reference's)
            self.in_features, self.out_features = module.in_features, module.out_features
        elif type(module).__name__ == 'Fourier_Embedding':
            self.kind = 'fourier'
            self.fourier = True
            self.in_features, self.out_features = module.in_features, 2 * module.out_features
        else:
            raise NotImplementedError(f"layer {type(module).__name__} has no jet kernel (Linear, RWF_Linear and "
                                      "Fourier_Embedding are implemented); no fallback exists")
        if self.kind != 'fourier' and getattr(module, 'bias', None) is None:
            raise NotImplementedError("Linear layers without bias are not supported")

    @property
    def weight(self):      # plain nn.Linear view (tests / tools that poke the leaf parameters directly)
        return self.module.weight

    @property
    def bias(self):
        return self.module.bias

    def tensors(self) -> Tuple[torch.Tensor, torch.Tensor]:
        m = self.module
        if self.kind == 'linear':
            return m.weight, m.bias
        if self.kind == 'rwf':
            return m.v * m.s, m.bias
        w = torch.cat([m.weight, m.weight])                    # sin half | cos half share the frequencies
        b = torch.cat([m.bias, m.bias]) if m.bias is not None else torch.zeros(self.out_features, device=w.device, dtype=w.dtype)
        return w, b


def extract_layers(net: nn.Module) -> Tuple[List[LayerSpec], int]:
    """Layer specs in forward order (last one un-activated) and the activation code.

    Accepts `MLP` / `Fourier_Net` of this package or of the reference (attributes layers, linear_out, activation;
    Linear, RWF_Linear, and a Fourier_Embedding first layer) and an nn.Sequential alternating Linear and Tanh/Sin
    ending in Linear (Tutorials/Spring.py:36)."""
    if all(hasattr(net, a) for a in ('layers', 'linear_out', 'activation')):
        mods = list(net.layers) + [net.linear_out]
        act = _act_code(net.activation)
    elif isinstance(net, nn.Sequential):
        seq = list(net)
        if len(seq) < 3 or len(seq) % 2 == 0:
            raise NotImplementedError("nn.Sequential must alternate Linear, activation, ..., Linear")
        mods = seq[0::2]
        acts = {_act_code(m) for m in seq[1::2]}
        if len(acts) != 1:
            raise NotImplementedError("all activations of the Sequential must be the same")
        act = acts.pop()
    else:
        raise NotImplementedError(f"{type(net).__name__} is not supported by the jet engine "
                                  "(MLP, Fourier_Net or nn.Sequential(Linear, Tanh, ..., Linear)); no fallback exists")
    specs = [LayerSpec(m) for m in mods]
    if any(sp.fourier for sp in specs[1:]):
        raise NotImplementedError("a Fourier embedding is only supported as the first layer")
    return specs, act
