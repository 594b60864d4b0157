"""GPU tests against the UNMODIFIED reference running live on the same B200 (`baseline/_ref`, materialised by
tools/install_ref.py; oracle/ref_live.py drives it through its own public API).

* `Jet_engine` is installed through the reference's own plugin point, `DE_Getter.set_deriv_method(engine_instance)`
  (src/torch_DE/continuous/DE_module.py:124-151), fed a batch cut by the reference's `PINN_dataset`, and its output goes
  through the reference's `Loss_handler` -- and must match `AD_engine` run on the same tensors, same network class.
* The reference's `FD_engine` and `GradNorm` outputs are pinned by fixtures generated from it (tests/golden/fd_engine.npz,
  gradnorm.npz).
"""
import os

import numpy as np
import pytest
import torch

import torch_de_b200 as T
from oracle import ref_autograd as O
from oracle import ref_live as RL
from tests.golden_util import GOLDEN, max_rel, rel_l2

pytestmark = pytest.mark.gpu

needs_reference = pytest.mark.skipif(not RL.available(), reason='baseline/_ref (tools/install_ref.py) is not in this tree')


@needs_reference
@pytest.mark.parametrize('name,n_col,n_bnd', [('ldc', 3000, 200), ('burgers', 5000, 100), ('cylinder_unsteady_small', 1500, 100)])
def test_jet_engine_plugs_into_the_reference_de_getter(name, n_col, n_bnd):
    R = RL.activate()
    owl = O._cylinder_unsteady(width=128, n_act=4, name='cylinder_unsteady_small') if name.endswith('small') else O.WORKLOADS[name]()
    n = owl.net
    params = O.init_mlp_params(n['d_in'], n['n_out'], n['width'], n['n_act'], seed=1234)
    groups = O.sample_groups(owl, n_col, n_bnd, seed=41)
    results = {}
    for which in ('AD', 'jet'):
        live = RL.LiveStep(owl, params, groups, torch.float32, 'cuda')
        assert type(live.pinn.deriv_method).__name__ == 'AD_engine'
        if which == 'jet':
            eng = T.Jet_engine(live.net, live.pinn.derivatives)
            assert isinstance(eng, R.engine)
            live.pinn.set_deriv_method(eng)                       # the reference's own plugin point
            assert live.pinn.deriv_method is eng
        streams, imap, row_names, rows, total, grads = live.evaluate()
        out = live.pinn.calculate(live.batch)
        assert set(out.keys()) == {'all'} | set(groups.keys())     # AD.py:98-112 grouping
        for g, x in groups.items():
            assert out[g]['u'].shape == (x.shape[0],)
        results[which] = (streams, rows, total, grads)
    # float64 truth from the reference itself
    live64 = RL.LiveStep(owl, [p.double() for p in params], groups, torch.float64, 'cuda')
    s64, _, _, rows64, total64, g64 = live64.evaluate()
    (s_ad, rows_ad, tot_ad, g_ad), (s_j, rows_j, tot_j, g_j) = results['AD'], results['jet']
    for k in s64:
        assert rel_l2(s_j[k], s64[k]) < 1e-5 and max_rel(s_j[k], s64[k]) < 1e-5, (k, rel_l2(s_j[k], s64[k]))
        assert rel_l2(s_j[k], s_ad[k]) < 1e-5, k
    assert torch.allclose(rows_j.double().cpu(), rows64.cpu(), rtol=1e-4, atol=1e-12)
    assert abs(float(tot_j) - float(total64)) <= 1e-4 * abs(float(total64))
    fj, f64, fa = (torch.cat([g.flatten().double().cpu() for g in gs]) for gs in (g_j, g64, g_ad))
    assert rel_l2(fj, f64) < 1e-4 and max_rel(fj, f64) < 1e-4, rel_l2(fj, f64)
    assert rel_l2(fj, f64) <= max(2 * rel_l2(fa, f64), 2e-6)       # no further from the truth than the reference's own FP32 run


def _fd_setup(z, device):
    net = T.MLP(2, 2, 64, 2)
    lins, _ = T.extract_layers(net)
    with torch.no_grad():
        for i, lin in enumerate(lins):
            lin.weight.copy_(torch.from_numpy(z[f'p_{2 * i}']))
            lin.bias.copy_(torch.from_numpy(z[f'p_{2 * i + 1}']))
    net = net.to(device)
    pinn = T.DE_Getter(net, ['x', 'y'], ['u', 'v'], ['u_x', 'u_yy', 'v_y'])
    pinn.set_deriv_method('FD', dxs=[1e-2, 2e-2], sdf=lambda p_: torch.minimum(p_[:, 0], p_[:, 1]))
    return net, lins, pinn


def test_fd_engine_matches_the_reference_fd_engine_golden():
    """SURVEY 8f row 3, pinned: outputs of the reference's FD_engine (Engines/FD.py:24-133) on MLP(2,2,64,2)."""
    z = np.load(os.path.join(GOLDEN, 'fd_engine.npz'))
    net, lins, pinn = _fd_setup(z, 'cuda')
    x = torch.from_numpy(z['x']).cuda()
    out = pinn(x)['all']
    assert list(out.keys()) == [str(k) for k in z['names']]
    # values: FP32 network evaluation against the reference's float64 run
    for k in ('u', 'v'):
        assert rel_l2(out[k].detach().cpu(), z['s64_' + k]) < 5e-6
    # differences amplify FP32 value noise by 1/(2h) and 1/h^2: the reference's OWN float32 FD output sets the scale (its
    # distance to its float64 run)
    for k in ('u_x', 'v_y', 'u_y', 'u_yy'):
        ours, r32, r64 = out[k].detach().cpu(), z['s32_' + k], z['s64_' + k]
        assert rel_l2(ours, r64) <= 3 * rel_l2(r32, r64) + 1e-6, (k, rel_l2(ours, r64), rel_l2(r32, r64))
    out['u_yy'].mean().backward()
    got = torch.cat([p.grad.flatten().cpu() for lin in lins for p in (lin.weight, lin.bias)])
    ref64 = torch.cat([torch.from_numpy(z[f'g64_{i}']).flatten() for i in range(2 * len(lins))])
    ref32 = torch.cat([torch.from_numpy(z[f'g32_{i}']).flatten() for i in range(2 * len(lins))])
    assert rel_l2(got, ref64) <= 3 * rel_l2(ref32, ref64) + 1e-5, (rel_l2(got, ref64), rel_l2(ref32, ref64))
    # grouped input: same group keys and per-group names as the reference
    batch = T.make_batch(['x', 'y'], {'a': x[:50], 'b': x[50:]})
    grouped = pinn.calculate(batch)
    assert list(grouped['a'].keys()) == [str(k) for k in z['group_a_names']]
    for g in ('a', 'b'):
        for k in ('u', 'u_x', 'u_yy'):
            r32, r64 = z[f'grp32_{g}_{k}'], z[f'grp64_{g}_{k}']
            assert rel_l2(grouped[g][k].detach().cpu(), r64) <= 3 * rel_l2(r32, r64) + 5e-6, (g, k)
    part = pinn.calculate(batch, target_groups='b')                 # base.py:41-54: values only for the other groups
    assert set(part['a']) == {'u', 'v'} and set(part['b']) == set(out)
    with pytest.raises(ValueError):
        T.FD_engine(net, {'u': (0,), 'u_xy': (0, 0, 1)}, dxs=[1e-3, 1e-3])


def test_gradnorm_matches_the_reference_gradnorm_golden():
    """SURVEY 8f row 2, pinned: weights the reference's GradNorm (utils/Loss_weighting.py:30-50) returns on the ldc_small table."""
    from torch_de_b200 import workloads as WL
    z = np.load(os.path.join(GOLDEN, 'gradnorm.npz'))
    owl = O._ldc(width=64, n_act=4, name='ldc_small')
    wl = WL.ldc(width=64, n_act=4, name='ldc_small')
    params = O.init_mlp_params(2, 3, 64, 4, seed=1234)
    groups = O.sample_groups(owl, 80, 24, 14)
    net = wl.make_net()
    lins, _ = T.extract_layers(net)
    with torch.no_grad():
        for i, lin in enumerate(lins):
            lin.weight.copy_(params[2 * i])
            lin.bias.copy_(params[2 * i + 1])
    net = net.cuda()
    pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
    batch = T.make_batch(wl.input_vars, {g: x.cuda() for g, x in groups.items()})
    loss = wl.build_losses(list(groups.keys()))(batch, pinn.calculate(batch))
    rows = loss.individual_losses()
    assert torch.allclose(rows.double().cpu(), torch.from_numpy(z['rows_64']), rtol=1e-4)
    w0 = torch.from_numpy(z['w0']).float().cuda()
    w1 = T.GradNorm(net, w0, *rows)
    assert T.GradNorm.last_path == 'fused'          # range-restricted sweeps on the shared stash, not k autograd passes
    w2 = T.GradNorm(net, w0, *rows, alpha=0.5, max_weight=50.0)
    assert torch.allclose(w1.double().cpu(), torch.from_numpy(z['w1_64']), rtol=2e-4)
    assert torch.allclose(w2.double().cpu(), torch.from_numpy(z['w2_64']), rtol=2e-4)


def test_adam_step_kernel_matches_torch_adam():
    """tde_adam_step (include/tde_jet.h) against torch.optim.Adam on the same flat buffers, 5 steps."""
    from torch_de_b200 import _lib
    torch.manual_seed(0)
    n = 100_003
    p0 = torch.randn(n, device='cuda')
    grads = [torch.randn(n, device='cuda') * (0.1 + i) for i in range(5)]
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=3e-3, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.clone(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    for i, g in enumerate(grads):
        ref.grad = g.clone()
        opt.step()
        rc = _lib.lib().tde_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 3e-3, 0.9, 0.999, 1e-8, i + 1,
                                      torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, 'tde_adam_step')
    st = opt.state[ref]
    assert torch.allclose(m, st['exp_avg'], rtol=2e-6, atol=1e-7)
    assert torch.allclose(v, st['exp_avg_sq'], rtol=2e-6, atol=1e-9)
    assert torch.allclose(p, ref.detach(), rtol=0, atol=1e-6)
    print('adam: max |dp| vs torch', float((p - ref.detach()).abs().max()), 'bit-equal fraction', float((p == ref.detach()).float().mean()))
