"""Network zoo subset on the hot path + introspection of a user's nn.Module into Linear layers.

`MLP` mirrors the reference constructor and layer placement (src/torch_DE/continuous/Networks/Networks.py:123-152):
`MLP(in, out, hidden, num_hidden_layers, activation)` has num_hidden_layers+1 activated Linear layers then
`linear_out`.  The jet kernels read the nn.Linear parameters in place on every call.
"""
from typing import List, Tuple

import torch
import torch.nn as nn

from . import _lib


def get_activation_function(activation):
    return getattr(torch, activation) if isinstance(activation, str) else activation


class MLP(nn.Module):
    def __init__(self, in_features: int, out_features: int, hidden_features: int, num_hidden_layers: int,
                 activation='tanh'):
        super().__init__()
        # construction order matters for seeded parity with the reference: linear_in, linear_out, hidden
        self.linear_in = nn.Linear(in_features, hidden_features)
        self.linear_out = nn.Linear(hidden_features, out_features)
        self.activation = get_activation_function(activation)
        self.layers = nn.ModuleList([self.linear_in] +
                                    [nn.Linear(hidden_features, hidden_features) for _ in range(num_hidden_layers)])

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


def extract_layers(net: nn.Module) -> Tuple[List[nn.Linear], int]:
    """Linear layers in forward order (last one un-activated) and the activation code.

    Accepts the reference's / this package's `MLP` (attributes layers, linear_out, activation) and an
    nn.Sequential alternating Linear and Tanh/Sin ending in Linear (Tutorials/Spring.py:36)."""
    if all(hasattr(net, a) for a in ('layers', 'linear_out', 'activation')):
        lins = list(net.layers) + [net.linear_out]
        act = _act_code(net.activation)
    elif isinstance(net, nn.Sequential):
        mods = list(net)
        if len(mods) < 3 or len(mods) % 2 == 0:
            raise NotImplementedError("nn.Sequential must alternate Linear, activation, ..., Linear")
        lins = mods[0::2]
        acts = {_act_code(m) for m in mods[1::2]}
        if len(acts) != 1:
            raise NotImplementedError("all activations of the Sequential must be the same")
        act = acts.pop()
    else:
        raise NotImplementedError(f"{type(net).__name__} is not supported by the jet engine "
                                  "(MLP or nn.Sequential(Linear, Tanh, ..., Linear)); no fallback exists")
    for m in lins:
        if type(m) is not nn.Linear:
            raise NotImplementedError(f"layer {type(m).__name__} is not a plain nn.Linear (RWF/Fourier layers are not built yet)")
        if m.bias is None:
            raise NotImplementedError("Linear layers without bias are not supported")
    return lins, act
