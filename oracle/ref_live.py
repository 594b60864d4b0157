"""ORACLE (test infrastructure, NOT product code): the UNMODIFIED reference, driven live through its own public API.

Only `tests/`, `tests/golden/make_golden.py`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `reference_gpu` /
`--impl reference` legs may import this file.  The product (`torch_de_b200`) never does.

`tools/install_ref.py` materialises `/root/reference/src/torch_DE` byte for byte under `baseline/_ref/` (git-ignored, shipped
to the GPU box with the working tree).  This module puts that copy on `sys.path` and wraps one training step exactly as the
tutorials write it (Tutorials/Lid_Driven_Cavity.py:70-100, Spring.py:36-59, AllenCahn.py:65-92, Stokes_Cylinder.py:84-119):

    PINN_dataset.add_group(...)                      torch_DE/utils/data.py:126-139
    DE_Getter(net, input_vars, output_vars, derivs)  torch_DE/continuous/DE_module.py:25-50   (AD_engine by default)
    out  = PINN.calculate(batch)                     DE_module.py:156-173 -> Engines/AD.py:42-68 (vmap(jacrev(jacrev(net))))
    loss = Loss_handler(dataset)(batch, out)         torch_DE/utils/loss.py:172-206
    total = combine(loss.grouped_losses('loss_type')); total.backward(); Adam.step()

`oracle/ref_autograd.py` (the port) only supplies the seeded point samplers and the weight initialiser so that both sides
see identical inputs.
"""
import math
import os
import sys
from typing import Dict, Sequence

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_autograd as O  # noqa: E402

_NS = None


def available() -> bool:
    from tools import install_ref
    return install_ref.available()


def activate():
    """Import the unmodified reference; returns a namespace with the symbols the hot path uses."""
    global _NS
    if _NS is not None:
        return _NS
    from tools import install_ref
    install_ref.activate()
    from types import SimpleNamespace
    from torch_DE.continuous import DE_Getter
    from torch_DE.continuous.Engines import AD_engine, FD_engine, engine
    from torch_DE.continuous.Networks import MLP, Fourier_Net
    from torch_DE.equations import DE_func, get_NavierStokes
    from torch_DE.utils import Loss_handler
    from torch_DE.utils.data import PINN_dataset, PINN_dict, PINN_group
    from torch_DE.utils.loss_weighting import GradNorm
    from torch_DE.utils.sampling import R3_sampler
    _NS = SimpleNamespace(DE_Getter=DE_Getter, AD_engine=AD_engine, FD_engine=FD_engine, engine=engine, MLP=MLP,
                          Fourier_Net=Fourier_Net, DE_func=DE_func, get_NavierStokes=get_NavierStokes,
                          Loss_handler=Loss_handler, PINN_dataset=PINN_dataset, PINN_dict=PINN_dict, PINN_group=PINN_group,
                          GradNorm=GradNorm, R3_sampler=R3_sampler)
    return _NS


def build_net(wl, params: Sequence[torch.Tensor], dtype, device='cpu'):
    """The reference's own network class holding `params` ([W0, b0, ..., Wout, bout] in forward order)."""
    R = activate()
    n = wl.net
    if wl.name == 'spring':                                   # Tutorials/Spring.py:36
        net = torch.nn.Sequential(torch.nn.Linear(1, n['width']), torch.nn.Tanh(), torch.nn.Linear(n['width'], 1))
        lins = [net[0], net[2]]
    else:
        net = R.MLP(n['d_in'], n['n_out'], n['width'], n['n_act'] - 1, activation=wl.activation)
        lins = list(net.layers) + [net.linear_out]
    with torch.no_grad():
        for i, lin in enumerate(lins):
            lin.weight.copy_(params[2 * i])
            lin.bias.copy_(params[2 * i + 1])
    return net.to(dtype).to(device), lins


def reference_losses(wl, dataset):
    """Loss table built with the reference's own Loss_handler, mirroring each tutorial."""
    R = activate()
    DE_func = R.DE_func
    L = R.Loss_handler(dataset)
    pi = torch.pi
    if wl.name == 'spring':
        L.add_initial_condition('initial condition', {'u': 0., 'u_t': 1.})
        L.add_residual('collocation_points', {'spring_eq': DE_func(lambda u, u_tt, **kw: u_tt + u)})
    elif wl.name == 'burgers':
        L.add_initial_condition('initial condition', {'u': DE_func(lambda x, **kw: -torch.sin(pi * x))})
        L.add_boundary('x0', {'u': 0})
        L.add_boundary('x1', {'u': 0})
        L.add_residual('collocation points',
                       {'burgers1D': DE_func(lambda u, u_t, u_x, u_xx, **kw: u_t + u * u_x - (0.01 / pi) * u_xx)})
    elif wl.name == 'allen_cahn':
        L.add_initial_condition('t0', {'u': DE_func(lambda x, **kw: x ** 2 * torch.cos(pi * x))})
        L.add_residual('collocation points',
                       {'AllenCahn': DE_func(lambda u_t, u_xx, u, **kw: u_t - 0.0001 * u_xx + 5 * u ** 3 - 5 * u)})
    elif wl.name in ('ldc', 'ldc_small'):
        sx = DE_func(lambda u, v, u_x, u_y, u_xx, u_yy, p_x, Re=100, **kw: u * u_x + v * u_y + p_x - 1 / Re * (u_xx + u_yy))
        sy = DE_func(lambda u, v, v_x, v_y, v_xx, v_yy, p_y, Re=100, **kw: u * v_x + v * v_y + p_y - 1 / Re * (v_xx + v_yy))
        inc = DE_func(lambda u_x, v_y, **kw: u_x + v_y)
        sdf = DE_func(lambda x, y, **kw: torch.minimum(0.1 - x.abs(), 0.1 - y.abs()))
        L.add_boundary('no slip', {'u': 0, 'v': 0})
        L.add_boundary('inlet', {'u': 1, 'v': 0}, weighting={'u': DE_func(lambda x, **kw: 1 - 10 * torch.abs(x)), 'v': 1})
        L.add_residual('collocation points', {'stokes_x': sx, 'stokes_y': sy, 'incomp': inc}, weighting=sdf)
    elif wl.name.startswith('cylinder'):
        steady = 'unsteady' not in wl.name
        _, _, _, eqs = R.get_NavierStokes(2, steady_state=steady, Re=100)     # Equations/NavierStokes.py:25-80
        L.add_boundary('no slip', {'u': 0, 'v': 0})
        L.add_boundary('inlet', {'u': 1, 'v': 0})
        if not steady:
            L.add_initial_condition('initial condition', {'u': DE_func(lambda y, **kw: 4 * y * (4.1 - y) / 4.1 ** 2), 'v': 0})
        L.add_residual('collocation points', eqs)
    else:
        raise KeyError(wl.name)
    return L


def combine(wl, type_losses):
    return wl.combine({k: type_losses[k] for k in type_losses.index})


class LiveStep:
    """One reference training step on fixed point groups, written as the tutorials write it."""

    def __init__(self, wl, params, groups: Dict[str, torch.Tensor], dtype=torch.float32, device='cpu', lr=1e-3, net=None):
        R = activate()
        self.wl = wl
        if net is None:
            self.net, self.lins = build_net(wl, params, dtype, device)
        else:
            self.net, self.lins = net, None
        self.dataset = R.PINN_dataset(wl.input_vars)
        for gname, xg in groups.items():
            self.dataset.add_group(gname, xg.to(dtype), batch_size=xg.shape[0])
        self.batch = self.dataset[{g: torch.arange(x.shape[0]) for g, x in groups.items()}].to(device)
        self.pinn = R.DE_Getter(self.net, wl.input_vars, wl.output_vars, list(wl.derivatives))
        self.handler = reference_losses(wl, self.dataset)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=lr)
        self.n_points = sum(x.shape[0] for x in groups.values())

    def forward_backward(self):
        out = self.pinn.calculate(self.batch)
        loss = self.handler(self.batch, out)
        total = combine(self.wl, loss.grouped_losses('loss_type'))
        total.backward()
        return out, loss, total

    def train_step(self):
        _, _, total = self.forward_backward()
        self.opt.step()
        self.opt.zero_grad()
        return total

    def evaluate(self):
        """streams of 'all', index map, row names, rows, total, parameter gradients (forward order W0,b0,...)."""
        self.opt.zero_grad()
        out, loss, total = self.forward_backward()
        df = loss.get_DataFrame()
        row_names = [f"{r['loss_type']}/{r['group']}/{r['variable']}" for _, r in df.iterrows()]
        rows = torch.stack(list(df['aggregated_loss'])).detach()
        grads = []
        if self.lins is not None:
            for lin in self.lins:
                grads += [lin.weight.grad.detach().clone(), lin.bias.grad.detach().clone()]
        else:
            grads = [p.grad.detach().clone() for p in self.net.parameters()]
        self.opt.zero_grad()
        return ({k: v.detach() for k, v in out['all'].items()}, dict(self.pinn.derivatives), row_names, rows, total.detach(), grads)


def run_reference(wl, params, groups, dtype, device='cpu'):
    return LiveStep(wl, params, groups, dtype, device).evaluate()
