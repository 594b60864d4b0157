"""ORACLE (test infrastructure, NOT product code): CPU restatement of torch_DE's PINN residual hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file.  The product (`torch_de_b200`) never does.

The reference's arithmetic for this path lives in a third-party dependency, PyTorch
(`torch.func.vmap/jacrev` + the autograd engine; pinned `torch==2.5.1+cu124` in
`/root/reference/requirements.txt:9`, this image has 2.11.0+cu128).  This file restates the
reference's *use* of it, function by function, with the file:line each piece follows:

  derivative_index_map   <- src/torch_DE/continuous/DE_module.py:84-122  (+ symbols/DE_Symbols.py:109-114)
  mlp_apply              <- src/torch_DE/continuous/Networks/Networks.py:141-152, Tutorials/Spring.py:36
  autodiff_all           <- src/torch_DE/continuous/Engines/AD.py:21-40, 70-90   (vmap(jacrev(jacrev(net))))
  assign_derivs          <- src/torch_DE/continuous/Engines/AD.py:115-131
  row_loss / total_loss  <- src/torch_DE/utils/loss.py:25-64, 244-249          (mean(w*|r|^2), sum rows)
  WORKLOADS residuals    <- Tutorials/Spring.py:24-33, Burgers1D.py:59-66, AllenCahn.py:35-42,
                            Lid_Driven_Cavity.py:32-67, Equations/NavierStokes.py:8-23

PINNED: `tests/golden/*.npz` hold outputs of the UNMODIFIED reference (DE_Getter + AD_engine +
Loss_handler + backward) generated in the build container by `tests/golden/make_golden.py`;
`tests/test_oracle_golden.py` checks this restatement against every one of them.
"""
import math
from typing import Callable, Dict, List, Sequence, Tuple

import torch
from torch.func import jacrev, vmap

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# spec: 'u_xx' -> (out_idx, in_idx, in_idx)      DE_module.py:53-122
# --------------------------------------------------------------------------------------
def derivative_index_map(input_vars: Sequence[str], output_vars: Sequence[str],
                         derivatives: Sequence[str]) -> Dict[str, Tuple[int, ...]]:
    """Ordered map name -> index tuple. Outputs first (set_vars, :66), then each requested
    derivative followed by its lower-order prefixes (get_deriv_index walks previous_Deriv, :115-122)."""
    in_idx = {v: i for i, v in enumerate(input_vars)}
    out_idx = {v: i for i, v in enumerate(output_vars)}
    table: Dict[str, Tuple[int, ...]] = {o: (i,) for i, o in enumerate(output_vars)}
    for name in derivatives:
        out, ins = name.split('_')
        assert out in out_idx, f'Output Variable {out} does not exist'
        for v in ins:
            assert v in in_idx, f'Variable {v} is not an input Variable'
        ins = list(ins)
        while ins:
            key = out + '_' + ''.join(ins)
            if key not in table:
                table[key] = (out_idx[out],) + tuple(in_idx[v] for v in ins)
            ins = ins[:-1]          # previous_Deriv drops the last variable (DE_Symbols.py:109-114)
    return table


# --------------------------------------------------------------------------------------
# network: params = [W0,b0, W1,b1, ..., Wout,bout]      Networks.py:148-152
# --------------------------------------------------------------------------------------
def mlp_apply(params: Sequence[Tensor], x: Tensor, activation: str = 'tanh') -> Tensor:
    act = getattr(torch, activation)            # Networks.py:4-7
    n_lin = len(params) // 2
    for l in range(n_lin - 1):
        x = act(torch.nn.functional.linear(x, params[2 * l], params[2 * l + 1]))
    return torch.nn.functional.linear(x, params[-2], params[-1])


def fourier_net_apply(sd: Dict[str, Tensor], x: Tensor, activation: str, n_hidden: int, rwf: bool) -> Tensor:
    """Fourier_Net.forward (Networks.py:104-120) from its state-dict tensors: Fourier_Embedding (:73-75) then
    activation(Linear) layers; RWF_Linear (:49-50) multiplies v by the per-row scale s."""
    act = getattr(torch, activation)

    def lin(prefix, z):
        if rwf:
            return torch.nn.functional.linear(z, sd[prefix + '.v'] * sd[prefix + '.s'], sd[prefix + '.bias'])
        return torch.nn.functional.linear(z, sd[prefix + '.weight'], sd[prefix + '.bias'])

    h = torch.nn.functional.linear(x, sd['linear_in.weight'])
    z = act(torch.cat([torch.sin(h), torch.cos(h)], dim=-1))
    for i in range(1, n_hidden + 1):
        z = act(lin(f'layers.{i}', z))
    return lin('linear_out', z)


def step_net(wl: 'Workload', net_fn: Callable[[Sequence[Tensor], Tensor], Tensor], leaves: Sequence[Tensor], groups: Dict[str, Tensor]):
    """`step` for an arbitrary network function net_fn(leaves, x): streams, loss rows, total, d(total)/d(leaves)."""
    leaves = [p.detach().clone().requires_grad_(True) for p in leaves]
    imap = wl.index_map()
    x_all = torch.cat(list(groups.values()))
    order = max(len(i) for i in imap.values()) - 1
    s_all = assign_derivs(autodiff_all(lambda z: net_fn(leaves, z), x_all, order), imap)
    row_names, row_vals, type_sums = [], [], {}
    start = 0
    for gname, xg in groups.items():
        sl = slice(start, start + xg.shape[0])
        start += xg.shape[0]
        o = {k: v[sl] for k, v in s_all.items()}
        for loss_type, rname, r, w in wl.rows(gname, xg, o):
            val = row_loss(r, w)
            row_names.append(f'{loss_type}/{gname}/{rname}')
            row_vals.append(val)
            type_sums[loss_type] = type_sums.get(loss_type, 0.0) + val
    total = wl.combine(type_sums)
    grads = torch.autograd.grad(total, leaves)
    return dict(streams={k: v.detach() for k, v in s_all.items()}, row_names=row_names,
                rows=torch.stack(row_vals).detach(), total=total.detach(), grads=[g.detach() for g in grads])


def init_mlp_params(d_in: int, n_out: int, width: int, n_act_layers: int, seed: int,
                    dtype=torch.float32) -> List[Tensor]:
    """Default nn.Linear init under torch.manual_seed(seed), in MLP construction order
    (linear_in, linear_out, then the hidden layers: Networks.py:141-145), returned in forward order."""
    torch.manual_seed(seed)
    lin_in = torch.nn.Linear(d_in, width)
    lin_out = torch.nn.Linear(width, n_out)
    hidden = [torch.nn.Linear(width, width) for _ in range(n_act_layers - 1)]
    out = []
    for lin in [lin_in] + hidden + [lin_out]:
        out += [lin.weight.detach().to(dtype).clone(), lin.bias.detach().to(dtype).clone()]
    return out


# --------------------------------------------------------------------------------------
# engine: vmap(jacrev(jacrev(net)))      AD.py:21-40, 70-90
# --------------------------------------------------------------------------------------
def autodiff_all(net_fn: Callable[[Tensor], Tensor], x: Tensor, highest_order: int) -> List[Tensor]:
    """[y[B,m], J[B,m,d], H[B,m,d,d], ...] exactly as AD_engine.autodiff un-nests them."""
    def with_aux(f, is_aux):
        # aux_function (base.py:76-93): first level returns (out, out), deeper levels (out[0], out)
        def g(z):
            out = f(z)
            return (out[0], out) if is_aux else (out, out)
        return g

    f = net_fn
    for level in range(highest_order):
        f = jacrev(with_aux(f, level > 0), has_aux=True)
    if highest_order == 0:
        return [net_fn(x)]
    nested = vmap(f)(x)
    derivs = []
    for _ in range(highest_order):
        dy, nested = nested
        derivs.append(dy)
    derivs.append(nested)
    return derivs[::-1]


def assign_derivs(derivs: List[Tensor], index_map: Dict[str, Tuple[int, ...]]) -> Dict[str, Tensor]:
    return {name: derivs[len(idx) - 1][(slice(None),) + idx] for name, idx in index_map.items()}


def streams(params, x, index_map, activation='tanh') -> Dict[str, Tensor]:
    order = max(len(i) for i in index_map.values()) - 1
    return assign_derivs(autodiff_all(lambda z: mlp_apply(params, z, activation), x, order), index_map)


def streams_double_backward(params, x, index_map, activation='tanh') -> Dict[str, Tensor]:
    """Independent second opinion: plain torch.autograd.grad(create_graph=True)."""
    x = x.detach().clone().requires_grad_(True)
    y = mlp_apply(params, x, activation)
    out = {}
    cache = {}

    def d(t, key):
        if key not in cache:
            cache[key] = torch.autograd.grad(t.sum(), x, create_graph=True)[0]
        return cache[key]

    for name, idx in index_map.items():
        t = y[:, idx[0]]
        key = (idx[0],)
        for i in idx[1:]:
            t = d(t, key)[:, i]
            key = key + (i,)
        out[name] = t
    return out


# --------------------------------------------------------------------------------------
# loss      utils/loss.py:25-64
# --------------------------------------------------------------------------------------
def row_loss(residual: Tensor, weight=1.0, power: int = 2) -> Tensor:
    return torch.mean(weight * torch.abs(residual).pow(power))


# --------------------------------------------------------------------------------------
# workloads: each returns the rows of the loss table for a batch
# --------------------------------------------------------------------------------------
class Workload:
    """name, input/output vars, requested derivative names, net shape, point groups and loss rows.

    groups: list of (group_name, sampler(n, gen) -> x[n,d]); the first group is the collocation set.
    rows(group_name, x, out) -> list of (loss_type, row_name, residual[B_g], weight)
    combine(type_sums: dict) -> total   (tutorial-specific recombination)
    """
    def __init__(self, name, input_vars, output_vars, derivatives, net, groups, rows, combine, activation='tanh'):
        self.name, self.input_vars, self.output_vars = name, input_vars, output_vars
        self.derivatives, self.net, self.groups = derivatives, net, groups
        self.rows, self.combine, self.activation = rows, combine, activation

    def index_map(self):
        return derivative_index_map(self.input_vars, self.output_vars, self.derivatives)


def _uniform(lo, hi):
    lo_t, hi_t = torch.tensor(lo, dtype=torch.float32), torch.tensor(hi, dtype=torch.float32)
    return lambda n, gen: lo_t + (hi_t - lo_t) * torch.rand(n, len(lo), generator=gen)


def _edge(fixed_dim, fixed_val, lo, hi):
    base = _uniform(lo, hi)

    def f(n, gen):
        x = base(n, gen)
        x[:, fixed_dim] = fixed_val
        return x
    return f


def _spring():
    def rows(g, x, o):
        if g == 'collocation_points':
            return [('residual', 'spring_eq', o['u_tt'] + o['u'], 1.0)]          # Spring.py:30
        return [('initial condition', 'u', o['u'] - 0.0, 1.0),                   # Spring.py:32
                ('initial condition', 'u_t', o['u_t'] - 1.0, 1.0)]
    return Workload('spring', ['t'], ['u'], ['u_t', 'u_tt'], dict(d_in=1, n_out=1, width=200, n_act=1),
                    [('collocation_points', _uniform([0.0], [2 * math.pi])),
                     ('initial condition', lambda n, gen: torch.zeros(n, 1))],
                    rows, lambda s: sum(s.values()))


def _burgers():
    nu = 0.01 / math.pi

    def rows(g, x, o):
        if g == 'collocation points':
            return [('residual', 'burgers1D', o['u_t'] + o['u'] * o['u_x'] - nu * o['u_xx'], 1.0)]   # Burgers1D.py:59-60
        if g == 'initial condition':
            return [('initial condition', 'u', o['u'] - (-torch.sin(math.pi * x[:, 0])), 1.0)]       # :63
        return [('boundary', 'u', o['u'] - 0.0, 1.0)]                                               # :64-65
    return Workload('burgers', ['x', 't'], ['u'], ['u_x', 'u_t', 'u_xx'], dict(d_in=2, n_out=1, width=64, n_act=8),
                    [('collocation points', _uniform([-1.0, 0.0], [1.0, 1.0])),
                     ('initial condition', _edge(1, 0.0, [-1.0, 0.0], [1.0, 1.0])),
                     ('x0', _edge(0, -1.0, [-1.0, 0.0], [1.0, 1.0])),
                     ('x1', _edge(0, 1.0, [-1.0, 0.0], [1.0, 1.0]))],
                    rows,
                    lambda s: 100 * s['initial condition'] + s['residual'] + 100 * s['boundary'])    # :77


def _allen_cahn():
    def rows(g, x, o):
        if g == 'collocation points':
            u = o['u']
            return [('residual', 'AllenCahn', o['u_t'] - 0.0001 * o['u_xx'] + 5 * u ** 3 - 5 * u, 1.0)]   # AllenCahn.py:40-42
        xx = x[:, 0]
        return [('initial condition', 'u', o['u'] - xx ** 2 * torch.cos(math.pi * xx), 1.0)]             # :36-38,60
    return Workload('allen_cahn', ['x', 't'], ['u'], ['u_t', 'u_xx'], dict(d_in=2, n_out=1, width=128, n_act=6),
                    [('collocation points', _uniform([-1.0, 0.0], [1.0, 1.0])),
                     ('t0', _edge(1, 0.0, [-1.0, 0.0], [1.0, 1.0]))],
                    rows, lambda s: sum(s.values()))


_NS_DERIVS = ['u_x', 'u_y', 'u_xx', 'u_yy', 'p_x', 'v_y', 'v_x', 'v_xx', 'v_yy', 'p_y']


def _ns_rows(o, Re):
    ns_x = o['u'] * o['u_x'] + o['v'] * o['u_y'] + o['p_x'] - 1 / Re * (o['u_xx'] + o['u_yy'])   # Lid_Driven_Cavity.py:33
    ns_y = o['u'] * o['v_x'] + o['v'] * o['v_y'] + o['p_y'] - 1 / Re * (o['v_xx'] + o['v_yy'])   # :38
    inc = o['u_x'] + o['v_y']                                                                    # :42
    return ns_x, ns_y, inc


def _ldc(width=256, n_act=8, name='ldc'):
    lo, hi = [-0.1, -0.1], [0.1, 0.1]

    def sdf(x):     # stand-in for domain.create_sdf: distance to the nearest wall
        return torch.minimum(0.1 - x[:, 0].abs(), 0.1 - x[:, 1].abs())

    def rows(g, x, o):
        if g == 'collocation points':
            w = sdf(x)
            ns_x, ns_y, inc = _ns_rows(o, 100.0)
            return [('residual', 'stokes_x', ns_x, w), ('residual', 'stokes_y', ns_y, w), ('residual', 'incomp', inc, w)]
        if g == 'inlet':
            return [('boundary', 'u', o['u'] - 1.0, 1 - 10 * torch.abs(x[:, 0])),       # :51,57-58
                    ('boundary', 'v', o['v'] - 0.0, 1.0)]
        return [('boundary', 'u', o['u'] - 0.0, 1.0), ('boundary', 'v', o['v'] - 0.0, 1.0)]   # :55-56

    def no_slip(n, gen):
        # left, right, bottom walls in equal thirds (Lid_Driven_Cavity.py:28)
        x = _uniform(lo, hi)(n, gen)
        k = n // 3
        x[:k, 0] = -0.1
        x[k:2 * k, 0] = 0.1
        x[2 * k:, 1] = -0.1
        return x
    return Workload(name, ['x', 'y'], ['u', 'v', 'p'], _NS_DERIVS, dict(d_in=2, n_out=3, width=width, n_act=n_act),
                    [('collocation points', _uniform(lo, hi)),
                     ('inlet', _edge(1, 0.1, lo, hi)),
                     ('no slip', no_slip)],
                    rows, lambda s: s['residual'] + 100 * s['boundary'])                # :92-94


_NS_DERIVS_T = ['u_x', 'u_y', 'u_t', 'u_xx', 'u_yy', 'p_x', 'v_x', 'v_y', 'v_t', 'v_xx', 'v_yy', 'p_y']


def _cylinder(width=512, n_act=10, name='cylinder', steady=True, activation='tanh'):
    """Cylinder-flow box with get_NavierStokes(2, steady_state, Re=100) residuals (Equations/NavierStokes.py:8-80,
    Tutorials/Stokes_Cylinder.py:21,65-79); the unsteady variant adds t in [0,1], the u_t / v_t terms (NavierStokes.py:9,14)
    and an initial-condition group with the tutorial's parabolic inlet profile (Stokes_Cylinder.py:64)."""
    d = 2 if steady else 3
    lo, hi = [0.0, 0.0] + ([] if steady else [0.0]), [22.0, 4.1] + ([] if steady else [1.0])

    def rows(g, x, o):
        if g == 'collocation points':
            ns_x, ns_y, inc = _ns_rows(o, 100.0)
            if not steady:
                ns_x, ns_y = o['u_t'] + ns_x, o['v_t'] + ns_y
            return [('residual', 'NavierStokes_x', ns_x, 1.0), ('residual', 'NavierStokes_y', ns_y, 1.0),
                    ('residual', 'incompressible', inc, 1.0)]
        if g == 'inlet':
            return [('boundary', 'u', o['u'] - 1.0, 1.0), ('boundary', 'v', o['v'] - 0.0, 1.0)]
        if g == 'initial condition':
            y = x[:, 1]
            return [('initial condition', 'u', o['u'] - 4 * y * (4.1 - y) / 4.1 ** 2, 1.0), ('initial condition', 'v', o['v'] - 0.0, 1.0)]
        return [('boundary', 'u', o['u'] - 0.0, 1.0), ('boundary', 'v', o['v'] - 0.0, 1.0)]
    groups = [('collocation points', _uniform(lo, hi)), ('inlet', _edge(0, 0.0, lo, hi)), ('no slip', _edge(1, 0.0, lo, hi))]
    if not steady:
        groups.append(('initial condition', _edge(2, 0.0, lo, hi)))
    return Workload(name, ['x', 'y'] + ([] if steady else ['t']), ['u', 'v', 'p'], _NS_DERIVS if steady else _NS_DERIVS_T,
                    dict(d_in=d, n_out=3, width=width, n_act=n_act), groups, rows,
                    lambda s: s['residual'] + 100 * s['boundary'] + (0 if steady else 100 * s['initial condition']), activation)


def _cylinder_unsteady(width=512, n_act=10, name='cylinder_unsteady', activation='tanh'):
    return _cylinder(width, n_act, name, steady=False, activation=activation)


WORKLOADS = {'spring': _spring, 'burgers': _burgers, 'allen_cahn': _allen_cahn, 'ldc': _ldc, 'cylinder': _cylinder,
             'cylinder_unsteady': _cylinder_unsteady}


def sample_groups(wl: Workload, n_col: int, n_bnd: int, seed: int) -> Dict[str, Tensor]:
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for i, (gname, sampler) in enumerate(wl.groups):
        out[gname] = sampler(n_col if i == 0 else n_bnd, gen)
    return out


def step(wl: Workload, params: Sequence[Tensor], groups: Dict[str, Tensor]):
    """One reference training-step evaluation: streams for every group (AD.py:62-64 differentiates all
    groups), loss rows, total, and d(total)/d(params) by autograd.  Returns a dict of detached tensors."""
    params = [p.detach().clone().requires_grad_(True) for p in params]
    imap = wl.index_map()
    x_all = torch.cat(list(groups.values()))                                   # base.py:24-26
    s_all = streams(params, x_all, imap, wl.activation)
    row_names, row_vals, type_sums = [], [], {}
    start = 0
    for gname, xg in groups.items():
        sl = slice(start, start + xg.shape[0])
        start += xg.shape[0]
        o = {k: v[sl] for k, v in s_all.items()}
        for loss_type, rname, r, w in wl.rows(gname, xg, o):
            val = row_loss(r, w)
            row_names.append(f'{loss_type}/{gname}/{rname}')
            row_vals.append(val)
            type_sums[loss_type] = type_sums.get(loss_type, 0.0) + val
    total = wl.combine(type_sums)
    grads = torch.autograd.grad(total, params)
    return dict(streams={k: v.detach() for k, v in s_all.items()},
                row_names=row_names, rows=torch.stack(row_vals).detach(),
                total=total.detach(), grads=[g.detach() for g in grads])


def step_chunked(wl: Workload, params: Sequence[Tensor], groups: Dict[str, Tensor], chunk: int = 8192,
                 keep_streams: bool = True):
    """`step` for batches too large for one vmap(jacrev(jacrev)) call: the same arithmetic evaluated chunk by chunk.

    Every loss row is a mean over its group (loss.py:26,57) and every tutorial's recombination is linear in the rows, so
    total = sum_rows c_row * (1/N_g) * sum_chunks sum_i w_i r_i^2 and d(total)/d(params) accumulates over chunks.  The
    coefficients c_row are read off `wl.combine` by autograd.  Runs in the dtype / on the device of `params` (the large-batch
    GPU parity tests run it in float64 on the B200 -- as the checker only)."""
    params = [p.detach().clone().requires_grad_(True) for p in params]
    imap = wl.index_map()
    dev, dtype = params[0].device, params[0].dtype
    # pass 0: row layout and the combination coefficients (from one point per group)
    layout = []
    with torch.no_grad():
        for gname, xg in groups.items():
            o = {k: torch.zeros(1, dtype=dtype, device=dev) for k in imap}
            for loss_type, rname, _, _ in wl.rows(gname, xg[:1].to(dev, dtype), o):
                layout.append((loss_type, gname, rname))
    leaves = torch.zeros(len(layout), dtype=dtype, device=dev, requires_grad=True)
    type_sums = {}
    for i, (lt, _, _) in enumerate(layout):
        type_sums[lt] = type_sums.get(lt, 0.0) + leaves[i]
    coef = torch.autograd.grad(wl.combine(type_sums), leaves)[0]
    rows = torch.zeros(len(layout), dtype=dtype, device=dev)
    grads = [torch.zeros_like(p) for p in params]
    streams_out = {k: [] for k in imap} if keep_streams else None
    r0 = 0
    for gname, xg in groups.items():
        n_g = xg.shape[0]
        n_rows_g = sum(1 for _, g, _ in layout if g == gname)
        for lo in range(0, n_g, chunk):
            xc = xg[lo:lo + chunk].to(dev, dtype)
            s = streams(params, xc, imap, wl.activation)
            part = torch.stack([(w * torch.abs(r).pow(2)).sum() / n_g if torch.is_tensor(w) else w * torch.abs(r).pow(2).sum() / n_g
                                for _, _, r, w in wl.rows(gname, xc, s)])
            g = torch.autograd.grad((coef[r0:r0 + n_rows_g] * part).sum(), params, allow_unused=True)
            for acc, gi in zip(grads, g):
                if gi is not None:
                    acc += gi
            rows[r0:r0 + n_rows_g] += part.detach()
            if keep_streams:
                for k, v in s.items():
                    streams_out[k].append(v.detach())
            del s, part, g
        r0 += n_rows_g
    total = (coef * rows).sum()
    return dict(streams={k: torch.cat(v) for k, v in streams_out.items()} if keep_streams else None,
                row_names=[f'{lt}/{g}/{r}' for lt, g, r in layout], rows=rows, total=total, grads=grads)
