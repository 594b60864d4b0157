"""CPU tests of the host-side mirror (Loss_handler / Loss / DE_func / get_derivatives / workloads): fed with the
reference's own golden streams, the loss table must reproduce the reference's rows and totals."""
import numpy as np
import pytest
import torch

import torch_de_b200 as T
from torch_de_b200 import workloads as WL
from tests.golden_util import FIXTURES, load


def product_workload(fixture):
    return WL.ldc(width=64, n_act=4, name='ldc_small') if fixture == 'ldc_small' else WL.WORKLOADS[fixture]()


@pytest.mark.parametrize('fixture', FIXTURES)
def test_loss_table_reproduces_reference_rows(fixture):
    z, owl, params, groups = load(fixture)
    wl = product_workload(fixture)
    names = [str(n) for n in z['names']]
    batch = T.make_batch(wl.input_vars, groups)
    out, lo = {}, 0
    for g, x in groups.items():
        n = x.shape[0]
        out[g] = {k: torch.from_numpy(z['s32_' + k][lo:lo + n]) for k in names}
        lo += n
    loss = wl.build_losses(list(groups.keys()))(batch, out)
    rows = {f"{r['loss_type']}/{r['group']}/{r['variable']}": float(r['aggregated_loss']) for r in
            (loss.aggregated_loss() and loss.losses)}
    ref = dict(zip([str(r) for r in z['row_names']], z['rows32']))
    assert set(rows) == set(ref)
    for k in rows:
        assert rows[k] == pytest.approx(float(ref[k]), rel=2e-6, abs=1e-12), k
    total = wl.combine(loss.grouped_losses('loss_type'))
    assert float(total) == pytest.approx(float(z['total32']), rel=2e-6)
    assert float(loss.sum()) == pytest.approx(float(sum(loss.individual_losses())), rel=1e-6)
    assert len(loss) == len(ref)


def test_get_derivatives_and_navier_stokes():
    ins, outs, ders, eqs = T.get_NavierStokes(2, steady_state=True, Re=100)
    assert ins == ('x', 'y') and outs == ('u', 'v', 'p')
    assert set(ders) == {'u_x', 'u_y', 'u_xx', 'u_yy', 'p_x', 'v_x', 'v_y', 'v_xx', 'v_yy', 'p_y'}
    assert list(eqs) == ['NavierStokes_x', 'NavierStokes_y', 'incompressible']
    ins, outs, ders, eqs = T.get_NavierStokes(2, steady_state=False, Re=10)
    assert ins == ('x', 'y', 't') and {'u_t', 'v_t'} <= set(ders)
    f = T.DE_func(lambda u, u_xx, x, **kw: u_xx + u * x)
    assert set(T.get_derivatives(['x'], ['u'], f.base_func)) == {'u_xx'}
    with pytest.raises(AssertionError):
        T.DE_func(lambda u: u)


def test_de_func_calling_convention_and_groups():
    g = T.PINN_group('g', torch.tensor([[1.0, 2.0], [3.0, 4.0]]), ['x', 't'], {'target': torch.tensor([5.0, 6.0])}, {'nu': 0.5})
    f = T.DE_func(lambda x, t, target, nu, u, **kw: u + x * t - target * nu)
    r = f(g, {'u': torch.tensor([1.0, 1.0]), 'u_x': torch.zeros(2)})
    assert torch.allclose(r, torch.tensor([1 + 2 - 2.5, 1 + 12 - 3.0]))
    sub = g.subgroup(torch.tensor([1]))
    assert sub.inputs['input'].shape == (1, 2) and float(sub.batchables['target']) == 6.0
    with pytest.raises(ValueError):
        T.PINN_group('bad', torch.zeros(2, 2), ['x', 't'], {'x': torch.zeros(2)})


def test_loss_handler_variants():
    L = T.Loss_handler(['a', 'b'])
    L.add_boundary('a', {'u': 1.0}, weighting=2.0)
    L.add_data_constraint('data', 'b', ['u'])
    L.add_periodic('a', 'b', 'u')
    with pytest.raises(ValueError):
        L.add_boundary('a', {'u': 0.0})
    with pytest.raises(AssertionError):
        L.add_boundary('zzz', {'u': 0.0})
    ga = T.PINN_group('a', torch.zeros(3, 1), ['x'])
    gb = T.PINN_group('b', torch.zeros(3, 1), ['x'], {'u': torch.tensor([1.0, 2.0, 3.0])})
    out = {'a': {'u': torch.tensor([2.0, 2.0, 2.0])}, 'b': {'u': torch.tensor([1.0, 1.0, 1.0])}}
    loss = L({'a': ga, 'b': gb}, out)
    g = loss.grouped_losses('loss_type')
    assert float(g['boundary']) == pytest.approx(2.0)            # mean(2 * (2-1)^2)
    assert float(g['data']) == pytest.approx((0 + 1 + 4) / 3)
    assert float(g['periodic']) == pytest.approx(1.0)
    assert float(L({'a': ga, 'b': gb}, out, power=1, aggregation_method='sum').grouped_losses('group')['a']) == pytest.approx(6.0 + 3.0)
    with pytest.raises(ValueError):
        loss.grouped_losses('nope')


def test_loss_weighting_rules_match_reference_formulas():
    """GradNorm / Causal_weighting (reference utils/Loss_weighting.py:30-67) on plain torch graphs: the update rules are host
    logic, independent of the kernels."""
    import math
    import torch
    import torch_de_b200 as T
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(2, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    x = torch.rand(7, 2)
    u = net(x)
    losses = [u.pow(2).mean(), 10.0 * (u - 1).pow(2).mean(), u.sum() * 0.0]      # the last one has zero gradient
    w0 = torch.tensor([1.0, 2.0, 3.0])
    w = T.GradNorm(net, w0, *losses, alpha=0.5, max_weight=100.0)
    mags = []
    for l in losses:
        g = torch.autograd.grad(l, list(net.parameters()), retain_graph=True, allow_unused=True)
        mags.append(math.sqrt(sum(float(t.pow(2).sum()) for t in g if t is not None)))
    total = sum(mags)
    want = [0.5 * float(w0[k]) + 0.5 * (min(total / mags[k], 100.0) if mags[k] > 1e-5 else float(w0[k])) for k in range(3)]
    assert torch.allclose(w, torch.tensor(want), rtol=1e-5)
    # causal weights: exp(-eps * running sum of the mean residual), first weight 1
    r1, r2 = torch.rand(6), torch.rand(6)
    cw = T.Causal_weighting({'residual': {'g': {'a': r1, 'b': r2}}}, eps=0.7)
    mean = (r1 + r2) / 2
    ref = torch.exp(-0.7 * torch.cumsum(mean, 0))
    ref[0] = 1
    assert torch.allclose(cw, ref)


def test_graphed_step_needs_a_device():
    import pytest
    import torch
    import torch_de_b200 as T
    if torch.cuda.is_available():
        pytest.skip('covered by the GPU test')
    with pytest.raises(RuntimeError):
        T.GraphedStep(lambda: torch.zeros(()))
