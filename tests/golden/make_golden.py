"""Generate tests/golden/*.npz by running the UNMODIFIED reference (`/root/reference`) in the build container.

    python tests/golden/make_golden.py

Each fixture holds, for one workload at a small size: the network parameters (or the seed that
regenerates them), the seeded points of every group, and what the reference produced through its own
public API -- DE_Getter(...).calculate(PINN_dict) -> Loss_handler(...) -> grouped/summed loss ->
backward() -- i.e. every named stream of the 'all' group, every loss-table row, the total and the
parameter gradients, in float32 and again in float64 ("truth").  The oracle (`oracle/ref_autograd.py`)
is only used here for its seeded point samplers and its weight initialiser, so that tests can rebuild
the identical inputs; all outputs come from torch_DE itself.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

ref_shim.install()
from torch_DE.continuous import DE_Getter  # noqa: E402
from torch_DE.continuous.Networks import MLP  # noqa: E402
from torch_DE.equations import DE_func  # noqa: E402
from torch_DE.utils import Loss_handler  # noqa: E402
from torch_DE.utils.data import PINN_dataset  # noqa: E402

from oracle import ref_autograd as O  # noqa: E402

torch.set_num_threads(8)


def build_net(wl, params, dtype):
    n = wl.net
    if wl.name == 'spring':                                   # Tutorials/Spring.py:36
        net = torch.nn.Sequential(torch.nn.Linear(1, n['width']), torch.nn.Tanh(), torch.nn.Linear(n['width'], 1))
        lins = [net[0], net[2]]
    else:
        net = MLP(n['d_in'], n['n_out'], n['width'], n['n_act'] - 1, activation=wl.activation)
        lins = list(net.layers) + [net.linear_out]
    with torch.no_grad():
        for i, lin in enumerate(lins):
            lin.weight.copy_(params[2 * i])
            lin.bias.copy_(params[2 * i + 1])
    return net.to(dtype), lins


def reference_losses(wl, dataset):
    """Loss table built with the reference's own Loss_handler, mirroring each tutorial."""
    L = Loss_handler(dataset)
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
    else:
        raise KeyError(wl.name)
    return L


def combine(wl, type_losses):
    d = {k: type_losses[k] for k in type_losses.index}
    return wl.combine(d)


def run_reference(wl, params, groups, dtype):
    net, lins = build_net(wl, params, dtype)
    dataset = PINN_dataset(wl.input_vars)
    for gname, xg in groups.items():
        dataset.add_group(gname, xg.to(dtype), batch_size=xg.shape[0])
    batch = dataset[{g: torch.arange(x.shape[0]) for g, x in groups.items()}]
    pinn = DE_Getter(net, wl.input_vars, wl.output_vars, list(wl.derivatives))
    out = pinn.calculate(batch)
    loss = reference_losses(wl, dataset)(batch, out)
    total = combine(wl, loss.grouped_losses('loss_type'))
    total.backward()
    df = loss.get_DataFrame()
    row_names = [f"{r['loss_type']}/{r['group']}/{r['variable']}" for _, r in df.iterrows()]
    rows = torch.stack(list(df['aggregated_loss'])).detach()
    grads = []
    for lin in lins:
        grads += [lin.weight.grad.detach().clone(), lin.bias.grad.detach().clone()]
    return ({k: v.detach() for k, v in out['all'].items()}, dict(pinn.derivatives), row_names, rows, total.detach(), grads)


CASES = [
    # (fixture, workload factory, n_col, n_bnd, seed, store_params)
    ('spring', O.WORKLOADS['spring'], 96, 1, 11, True),
    ('burgers', O.WORKLOADS['burgers'], 96, 24, 12, True),
    ('allen_cahn', O.WORKLOADS['allen_cahn'], 64, 16, 13, False),
    ('ldc_small', lambda: O._ldc(width=64, n_act=4, name='ldc_small'), 80, 24, 14, True),
    ('ldc', O.WORKLOADS['ldc'], 48, 12, 15, False),
]


def fourier_case():
    """AllenCahn tutorial network: Fourier_Net(2,1,100,4,RWF=True) (Tutorials/AllenCahn.py:65) on the Allen-Cahn loss table."""
    from torch_DE.continuous.Networks import Fourier_Net
    wl = O.WORKLOADS['allen_cahn']()
    groups = O.sample_groups(wl, 64, 16, 16)
    blob = {}
    for tag, dtype in (('32', torch.float32), ('64', torch.float64)):
        torch.manual_seed(1234)
        net = Fourier_Net(2, 1, 100, 4, RWF=True).to(dtype)
        names = [n for n, _ in net.named_parameters()]
        if tag == '32':
            for n, p_ in net.named_parameters():
                blob['p_' + n] = p_.detach().numpy().copy()
        dataset = PINN_dataset(wl.input_vars)
        for gname, xg in groups.items():
            dataset.add_group(gname, xg.to(dtype), batch_size=xg.shape[0])
        batch = dataset[{g: torch.arange(x.shape[0]) for g, x in groups.items()}]
        pinn = DE_Getter(net, wl.input_vars, wl.output_vars, list(wl.derivatives))
        out = pinn.calculate(batch)
        loss = reference_losses(wl, dataset)(batch, out)
        total = combine(wl, loss.grouped_losses('loss_type'))
        total.backward()
        df = loss.get_DataFrame()
        blob['row_names'] = np.array([f"{r['loss_type']}/{r['group']}/{r['variable']}" for _, r in df.iterrows()])
        blob['rows' + tag] = torch.stack(list(df['aggregated_loss'])).detach().numpy()
        blob['total' + tag] = total.detach().numpy()
        for k, v in out['all'].items():
            blob[f's{tag}_{k}'] = v.detach().numpy()
        for n, p_ in net.named_parameters():
            blob[f'g{tag}_{n}'] = p_.grad.detach().numpy().astype(np.float64 if tag == '64' else np.float32)
        blob['names'] = np.array(list(pinn.derivatives.keys()))
        blob['param_names'] = np.array(names)
    blob['group_names'] = np.array(list(groups.keys()))
    for g, x in groups.items():
        blob['x_' + g] = x.numpy()
    path = os.path.join(HERE, 'allen_cahn_fourier.npz')
    np.savez_compressed(path, **blob)
    print('allen_cahn_fourier total32', float(blob['total32']), 'total64', float(blob['total64']), os.path.getsize(path), 'bytes')


def main():
    fourier_case()
    for fixture, factory, n_col, n_bnd, seed, store_params in CASES:
        wl = factory()
        n = wl.net
        params = O.init_mlp_params(n['d_in'], n['n_out'], n['width'], n['n_act'], seed=1234)
        groups = O.sample_groups(wl, n_col, n_bnd, seed)
        s32, imap, row_names, rows32, tot32, g32 = run_reference(wl, params, groups, torch.float32)
        s64, _, _, rows64, tot64, g64 = run_reference(wl, [p.double() for p in params], groups, torch.float64)
        blob = dict(seed=np.int64(seed), n_col=np.int64(n_col), n_bnd=np.int64(n_bnd), param_seed=np.int64(1234),
                    names=np.array(list(imap.keys())), index_tuples=np.array([str(v) for v in imap.values()]),
                    row_names=np.array(row_names), rows32=rows32.numpy(), rows64=rows64.numpy(),
                    total32=tot32.numpy(), total64=tot64.numpy(), group_names=np.array(list(groups.keys())))
        for g, x in groups.items():
            blob['x_' + g] = x.numpy()
        for k in s32:
            blob['s32_' + k] = s32[k].numpy()
            blob['s64_' + k] = s64[k].numpy()
        for i in range(len(g32)):
            blob[f'g64_{i}'] = g64[i].numpy().astype(np.float64)
            if store_params:
                blob[f'p_{i}'] = params[i].numpy()
                blob[f'g32_{i}'] = g32[i].numpy()
        if not store_params:       # big nets: keep the fixture small -- gradient digests only
            flat32 = torch.cat([g.flatten() for g in g32]).double()
            flat64 = torch.cat([g.flatten() for g in g64])
            proj = torch.Generator().manual_seed(99)
            R = torch.randn(64, flat64.numel(), generator=proj, dtype=torch.float64) / flat64.numel() ** 0.5
            blob['gflat32_norm'] = flat32.norm().numpy()
            blob['gflat64_norm'] = flat64.norm().numpy()
            blob['gproj32'] = (R @ flat32).numpy()
            blob['gproj64'] = (R @ flat64).numpy()
            for i in range(len(g64)):
                del blob[f'g64_{i}']
            # last layer + first layer gradients in full (small)
            for i in (0, 1, len(g32) - 2, len(g32) - 1):
                blob[f'g32_{i}'] = g32[i].numpy()
                blob[f'g64_{i}'] = g64[i].numpy()
        path = os.path.join(HERE, fixture + '.npz')
        np.savez_compressed(path, **blob)
        print(fixture, 'total32', float(tot32), 'total64', float(tot64), os.path.getsize(path), 'bytes', list(imap.items())[:4])


if __name__ == '__main__':
    main()
