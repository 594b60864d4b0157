"""The benchmark workloads of BASELINE.json expressed through this package's public
API -- network, derivative names, synthetic point groups, loss table, tutorial-specific recombination.

Each follows the corresponding tutorial of the reference: Tutorials/Spring.py:24-37, Burgers1D.py:40-77,
AllenCahn.py:35-65, Lid_Driven_Cavity.py:32-94, Stokes_Cylinder.py (steady Navier-Stokes set).
"""
import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List

import torch

from .equations import DE_func, get_NavierStokes
from .loss import Loss_handler
from .networks import MLP


@dataclass
class Workload:
    name: str
    input_vars: List[str]
    output_vars: List[str]
    derivatives: List[str]
    width: int
    n_act: int                                   # activated Linear layers ("L x W" in BASELINE.json)
    groups: Dict[str, Callable]                  # name -> sampler(n, generator, device) -> [n, d]; first = collocation
    build_losses: Callable[[List[str]], Loss_handler] = None
    combine: Callable = None                     # grouped_losses('loss_type') -> total
    activation: str = 'tanh'
    default_points: int = 0
    default_boundary: int = 0
    extras: dict = field(default_factory=dict)

    def make_net(self) -> torch.nn.Module:
        d, m = len(self.input_vars), len(self.output_vars)
        if self.name == 'spring':                # Tutorials/Spring.py:36
            return torch.nn.Sequential(torch.nn.Linear(d, self.width), torch.nn.Tanh(), torch.nn.Linear(self.width, m))
        return MLP(d, m, self.width, self.n_act - 1, activation=self.activation)

    def sample(self, n_col: int, n_bnd: int, seed: int, device='cpu') -> Dict[str, torch.Tensor]:
        gen = torch.Generator().manual_seed(seed)
        out = {}
        for i, (g, sampler) in enumerate(self.groups.items()):
            out[g] = sampler(n_col if i == 0 else n_bnd, gen).to(device)
        return out

    def flops_per_point(self, n_channels: int) -> float:
        """F = 6 * J * sum_layers(W_in * W_out) (2 fwd + 2 dX + 2 dW per MAC per channel)."""
        d, m, w = len(self.input_vars), len(self.output_vars), self.width
        macs = d * w + (self.n_act - 1) * w * w + w * m
        return 6.0 * n_channels * macs


def _uniform(lo, hi):
    lo_t, hi_t = torch.tensor(lo, dtype=torch.float32), torch.tensor(hi, dtype=torch.float32)
    return lambda n, gen: lo_t + (hi_t - lo_t) * torch.rand(n, len(lo), generator=gen)


def _edge(dim, val, lo, hi):
    base = _uniform(lo, hi)

    def f(n, gen):
        x = base(n, gen)
        x[:, dim] = val
        return x
    return f


def spring():
    def losses(names):
        L = Loss_handler(names)
        L.add_initial_condition('initial condition', {'u': 0., 'u_t': 1.})
        L.add_residual('collocation_points', {'spring_eq': DE_func(lambda u, u_tt, **kw: u_tt + u)})
        return L
    return Workload('spring', ['t'], ['u'], ['u_t', 'u_tt'], 200, 1,
                    {'collocation_points': _uniform([0.0], [2 * math.pi]),
                     'initial condition': lambda n, gen: torch.zeros(n, 1)},
                    losses, lambda s: sum(s.values()), default_points=1000, default_boundary=1)


def burgers():
    nu = 0.01 / math.pi

    def losses(names):
        L = Loss_handler(names)
        L.add_initial_condition('initial condition', {'u': DE_func(lambda x, **kw: -torch.sin(math.pi * x))})
        L.add_boundary('x0', {'u': 0})
        L.add_boundary('x1', {'u': 0})
        L.add_residual('collocation points', {'burgers1D': DE_func(lambda u, u_t, u_x, u_xx, **kw: u_t + u * u_x - nu * u_xx)})
        return L
    box = ([-1.0, 0.0], [1.0, 1.0])
    return Workload('burgers', ['x', 't'], ['u'], ['u_x', 'u_t', 'u_xx'], 64, 8,
                    {'collocation points': _uniform(*box), 'initial condition': _edge(1, 0.0, *box),
                     'x0': _edge(0, -1.0, *box), 'x1': _edge(0, 1.0, *box)},
                    losses, lambda s: 100 * s['initial condition'] + s['residual'] + 100 * s['boundary'],
                    default_points=100_000, default_boundary=100)


def allen_cahn():
    def losses(names):
        L = Loss_handler(names)
        L.add_initial_condition('t0', {'u': DE_func(lambda x, **kw: x ** 2 * torch.cos(math.pi * x))})
        L.add_residual('collocation points',
                       {'AllenCahn': DE_func(lambda u_t, u_xx, u, **kw: u_t - 0.0001 * u_xx + 5 * u ** 3 - 5 * u)})
        return L
    box = ([-1.0, 0.0], [1.0, 1.0])
    return Workload('allen_cahn', ['x', 't'], ['u'], ['u_t', 'u_xx'], 128, 6,
                    {'collocation points': _uniform(*box), 't0': _edge(1, 0.0, *box)},
                    losses, lambda s: sum(s.values()), default_points=1_000_000, default_boundary=1000)


_NS_DERIVS = ['u_x', 'u_y', 'u_xx', 'u_yy', 'p_x', 'v_y', 'v_x', 'v_xx', 'v_yy', 'p_y']


def _ns_losses(names, sdf=None, inlet_w=None, res_names=('stokes_x', 'stokes_y', 'incomp')):
    sx = DE_func(lambda u, v, u_x, u_y, u_xx, u_yy, p_x, Re=100, **kw: u * u_x + v * u_y + p_x - 1 / Re * (u_xx + u_yy))
    sy = DE_func(lambda u, v, v_x, v_y, v_xx, v_yy, p_y, Re=100, **kw: u * v_x + v * v_y + p_y - 1 / Re * (v_xx + v_yy))
    inc = DE_func(lambda u_x, v_y, **kw: u_x + v_y)
    L = Loss_handler(names)
    L.add_boundary('no slip', {'u': 0, 'v': 0})
    L.add_boundary('inlet', {'u': 1, 'v': 0}, weighting={'u': inlet_w, 'v': 1} if inlet_w is not None else 1)
    L.add_residual('collocation points', dict(zip(res_names, (sx, sy, inc))), weighting=sdf if sdf is not None else 1)
    return L


def ldc(width=256, n_act=8, name='ldc'):
    lo, hi = [-0.1, -0.1], [0.1, 0.1]
    sdf = DE_func(lambda x, y, **kw: torch.minimum(0.1 - x.abs(), 0.1 - y.abs()))      # wall-distance stand-in
    inlet_w = DE_func(lambda x, **kw: 1 - 10 * torch.abs(x))                           # Lid_Driven_Cavity.py:51

    def no_slip(n, gen):
        x = _uniform(lo, hi)(n, gen)
        k = n // 3
        x[:k, 0] = -0.1
        x[k:2 * k, 0] = 0.1
        x[2 * k:, 1] = -0.1
        return x
    return Workload(name, ['x', 'y'], ['u', 'v', 'p'], list(_NS_DERIVS), width, n_act,
                    {'collocation points': _uniform(lo, hi), 'inlet': _edge(1, 0.1, lo, hi), 'no slip': no_slip},
                    lambda names: _ns_losses(names, sdf, inlet_w),
                    lambda s: s['residual'] + 100 * s['boundary'], default_points=4_000_000, default_boundary=40_000)


_NS_DERIVS_T = ['u_x', 'u_y', 'u_t', 'u_xx', 'u_yy', 'p_x', 'v_x', 'v_y', 'v_t', 'v_xx', 'v_yy', 'p_y']


def cylinder(width=512, n_act=10, name='cylinder', steady=True, activation='tanh'):
    """Cylinder-flow box with the residuals of get_NavierStokes(2, steady_state, Re=100) (Equations/NavierStokes.py:25-80,
    Tutorials/Stokes_Cylinder.py:65-79); the unsteady variant (Unsteady_Cylinder.py:146) adds t in [0, 1] and an
    initial-condition group with the parabolic inlet profile of Stokes_Cylinder.py:64."""
    lo, hi = [0.0, 0.0] + ([] if steady else [0.0]), [22.0, 4.1] + ([] if steady else [1.0])

    def losses(names):
        _, _, _, eqs = get_NavierStokes(2, steady_state=steady, Re=100)
        L = Loss_handler(names)
        L.add_boundary('no slip', {'u': 0, 'v': 0})
        L.add_boundary('inlet', {'u': 1, 'v': 0})
        if not steady:
            L.add_initial_condition('initial condition', {'u': DE_func(lambda y, **kw: 4 * y * (4.1 - y) / 4.1 ** 2), 'v': 0})
        L.add_residual('collocation points', eqs)
        return L
    groups = {'collocation points': _uniform(lo, hi), 'inlet': _edge(0, 0.0, lo, hi), 'no slip': _edge(1, 0.0, lo, hi)}
    if not steady:
        groups['initial condition'] = _edge(2, 0.0, lo, hi)
    return Workload(name, ['x', 'y'] + ([] if steady else ['t']), ['u', 'v', 'p'], list(_NS_DERIVS if steady else _NS_DERIVS_T),
                    width, n_act, groups, losses,
                    lambda s: s['residual'] + 100 * s['boundary'] + (0 if steady else 100 * s['initial condition']),
                    activation=activation, default_points=16_000_000, default_boundary=40_000)


def cylinder_unsteady(width=512, n_act=10, name='cylinder_unsteady', activation='tanh'):
    return cylinder(width, n_act, name, steady=False, activation=activation)


WORKLOADS = {'spring': spring, 'burgers': burgers, 'allen_cahn': allen_cahn, 'ldc': ldc, 'cylinder': cylinder,
             'cylinder_unsteady': cylinder_unsteady}
