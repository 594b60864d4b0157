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
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_autograd as O  # noqa: E402
from oracle import ref_live as RL  # noqa: E402

R = RL.activate()
DE_Getter, PINN_dataset = R.DE_Getter, R.PINN_dataset
reference_losses, combine, run_reference = RL.reference_losses, RL.combine, RL.run_reference

torch.set_num_threads(8)


CASES = [
    # (fixture, workload factory, n_col, n_bnd, seed, store_params)
    ('spring', O.WORKLOADS['spring'], 96, 1, 11, True),
    ('burgers', O.WORKLOADS['burgers'], 96, 24, 12, True),
    ('allen_cahn', O.WORKLOADS['allen_cahn'], 64, 16, 13, False),
    ('ldc_small', lambda: O._ldc(width=64, n_act=4, name='ldc_small'), 80, 24, 14, True),
    ('ldc', O.WORKLOADS['ldc'], 48, 12, 15, False),
    # Tutorials/Stokes_Cylinder.py + Equations/NavierStokes.py:25-80: get_NavierStokes(2, steady_state, Re=100) on MLP 10x512
    ('cylinder', O.WORKLOADS['cylinder'], 40, 8, 17, False),                      # steady, J = 5
    ('cylinder_unsteady', O.WORKLOADS['cylinder_unsteady'], 40, 8, 18, False),    # x, y, t: J = 6
]


def fourier_case():
    """AllenCahn tutorial network: Fourier_Net(2,1,100,4,RWF=True) (Tutorials/AllenCahn.py:65) on the Allen-Cahn loss table."""
    Fourier_Net = R.Fourier_Net
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


def fd_engine_case():
    """The reference's FD_engine (Engines/FD.py:4-133) on MLP(2,2,64,2): central differences with an sdf-clipped step, plain
    tensor input and a grouped (PINN_dict) input."""
    blob = {}
    params = O.init_mlp_params(2, 2, 64, 3, seed=4)
    x = torch.rand(200, 2, generator=torch.Generator().manual_seed(6)) * 0.8 + 0.1
    sdf = lambda p_: torch.minimum(p_[:, 0], p_[:, 1])            # noqa: E731
    ders = ['u_x', 'u_yy', 'v_y']
    for tag, dtype in (('32', torch.float32), ('64', torch.float64)):
        net = R.MLP(2, 2, 64, 2).to(dtype)
        lins = list(net.layers) + [net.linear_out]
        with torch.no_grad():
            for i, lin in enumerate(lins):
                lin.weight.copy_(params[2 * i]); lin.bias.copy_(params[2 * i + 1])
        pinn = DE_Getter(net, ['x', 'y'], ['u', 'v'], ders)
        pinn.set_deriv_method('FD', dxs=[1e-2, 2e-2], sdf=sdf)
        out = pinn.calculate(x.to(dtype))['all']
        blob['names'] = np.array(list(out.keys()))
        for k, v in out.items():
            blob[f's{tag}_{k}'] = v.detach().numpy()
        out['u_yy'].mean().backward()
        for i, lin in enumerate(lins):
            blob[f'g{tag}_{2 * i}'] = lin.weight.grad.numpy().copy()
            blob[f'g{tag}_{2 * i + 1}'] = lin.bias.grad.numpy().copy()
        dataset = PINN_dataset(['x', 'y'])
        dataset.add_group('a', x[:50].to(dtype), batch_size=50)
        dataset.add_group('b', x[50:].to(dtype), batch_size=150)
        batch = dataset[{'a': torch.arange(50), 'b': torch.arange(150)}]
        # (target_groups raises inside the reference: FD.py:61 builds a plain dict that base.py:16-30 rejects)
        grouped = pinn.calculate(batch)
        blob['group_a_names'] = np.array(list(grouped['a'].keys()))
        blob['group_b_names'] = np.array(list(grouped['b'].keys()))
        for g in ('a', 'b'):
            for k, v in grouped[g].items():
                blob[f'grp{tag}_{g}_{k}'] = v.detach().numpy()
    blob['x'] = x.numpy()
    for i, p_ in enumerate(params):
        blob[f'p_{i}'] = p_.numpy()
    path = os.path.join(HERE, 'fd_engine.npz')
    np.savez_compressed(path, **blob)
    print('fd_engine', list(blob['names']), os.path.getsize(path), 'bytes')


def gradnorm_case():
    """The reference's GradNorm (utils/Loss_weighting.py:30-50; call site Tutorials/Stokes_Cylinder.py:107-110) on the
    ldc_small loss table: k x autograd.grad(retain_graph=True, allow_unused=True) -> norm-balanced weights."""
    wl = O._ldc(width=64, n_act=4, name='ldc_small')
    params = O.init_mlp_params(2, 3, 64, 4, seed=1234)
    groups = O.sample_groups(wl, 80, 24, 14)
    blob = {}
    for tag, dtype in (('32', torch.float32), ('64', torch.float64)):
        live = RL.LiveStep(wl, [p_.to(dtype) for p_ in params], groups, dtype)
        out = live.pinn.calculate(live.batch)
        rows = live.handler(live.batch, out).individual_losses()
        w0 = torch.linspace(0.5, 2.0, len(rows), dtype=dtype)
        w1 = R.GradNorm(live.net, w0, *rows)
        w2 = R.GradNorm(live.net, w0, *rows, alpha=0.5, max_weight=50.0)
        ps = list(live.net.parameters())
        mags = [torch.sqrt(sum(g.pow(2).sum() for g in torch.autograd.grad(r, ps, retain_graph=True, allow_unused=True) if g is not None))
                for r in rows]
        blob['w0'] = w0.double().numpy()
        blob['w1_' + tag] = w1.detach().double().numpy()
        blob['w2_' + tag] = w2.detach().double().numpy()
        blob['mags_' + tag] = torch.stack(mags).detach().double().numpy()
        blob['rows_' + tag] = rows.detach().double().numpy()
    path = os.path.join(HERE, 'gradnorm.npz')
    np.savez_compressed(path, **blob)
    print('gradnorm', blob['w1_64'], os.path.getsize(path), 'bytes')


def main():
    fourier_case()
    fd_engine_case()
    gradnorm_case()
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
