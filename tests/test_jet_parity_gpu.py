"""GPU parity: the sm_100a jet path through the public API (DE_Getter -> Jet_engine -> Loss_handler ->
backward) against (a) the golden vectors produced by the unmodified reference and (b) the oracle on the
same seeded inputs.  Tolerances are the north star's: <=1e-5 (relL2 and max|d|/max|ref|) on u and every
derivative, <=1e-4 on each loss row, the total and the parameter gradient."""
import numpy as np
import pytest
import torch

import torch_de_b200 as T
from torch_de_b200 import workloads as WL
from tests.golden_util import FIXTURES, load, max_rel, rel_l2

pytestmark = pytest.mark.gpu

TOL_STREAM, TOL_LOSS = 1e-5, 1e-4


def product_workload(fixture):
    if fixture == 'ldc_small':
        return WL.ldc(width=64, n_act=4, name='ldc_small')
    return WL.WORKLOADS[fixture]()


def load_net(wl, params, device):
    net = wl.make_net()
    lins, _ = T.extract_layers(net)
    with torch.no_grad():
        for i, lin in enumerate(lins):
            lin.weight.copy_(params[2 * i])
            lin.bias.copy_(params[2 * i + 1])
    return net.to(device), lins


def run_product(fixture, device='cuda'):
    z, owl, params, groups = load(fixture)
    wl = product_workload(fixture)
    net, lins = load_net(wl, params, device)
    pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
    batch = T.make_batch(wl.input_vars, {g: x.to(device) for g, x in groups.items()})
    out = pinn.calculate(batch)
    loss = wl.build_losses(list(groups.keys()))(batch, out)
    total = wl.combine(loss.grouped_losses('loss_type'))
    total.backward()
    grads = []
    for lin in lins:
        grads += [lin.weight.grad.detach().cpu(), lin.bias.grad.detach().cpu()]
    rows = {f"{r['loss_type']}/{r['group']}/{r['variable']}": float(r['aggregated_loss']) for r in loss.losses}
    return z, owl, params, groups, pinn, out, rows, float(total), grads


@pytest.mark.parametrize('fixture', FIXTURES)
def test_streams_loss_grads_match_reference_golden(fixture):
    z, owl, params, groups, pinn, out, rows, total, grads = run_product(fixture)
    assert list(pinn.derivatives.keys()) == [str(n) for n in z['names']]
    for name in pinn.derivatives:
        got = out['all'][name].detach().cpu()
        assert got.shape == (sum(x.shape[0] for x in groups.values()),)
        # the FP64 run of the reference is the truth; FP32 reference must not be closer by more than the tolerance
        assert rel_l2(got, z['s64_' + name]) < TOL_STREAM, (name, rel_l2(got, z['s64_' + name]))
        assert max_rel(got, z['s64_' + name]) < TOL_STREAM, name
        # against the reference's own float32 run: two float32 results can differ by the sum of their distances to the truth
        assert rel_l2(got, z['s32_' + name]) < TOL_STREAM + rel_l2(z['s32_' + name], z['s64_' + name]), name
    ref_rows = dict(zip([str(r) for r in z['row_names']], z['rows64']))
    assert set(ref_rows) == set(rows)
    for k, v in rows.items():
        assert abs(v - ref_rows[k]) <= TOL_LOSS * abs(ref_rows[k]) + 1e-12, (k, v, ref_rows[k])
    assert abs(total - float(z['total64'])) <= TOL_LOSS * abs(float(z['total64']))
    for i, g in enumerate(grads):
        if f'g64_{i}' in z:
            assert rel_l2(g, z[f'g64_{i}']) < TOL_LOSS, (i, rel_l2(g, z[f'g64_{i}']))
    if 'gproj64' in z:
        flat = torch.cat([g.flatten() for g in grads]).double()
        gen = torch.Generator().manual_seed(99)
        R = torch.randn(64, flat.numel(), generator=gen, dtype=torch.float64) / flat.numel() ** 0.5
        assert rel_l2(R @ flat, z['gproj64']) < TOL_LOSS
        assert abs(float(flat.norm()) - float(z['gflat64_norm'])) < TOL_LOSS * float(z['gflat64_norm'])


@pytest.mark.parametrize('fixture', ['burgers', 'ldc'])
def test_full_gradient_matches_oracle(fixture):
    """Every parameter gradient against the oracle's autograd (float64) on the fixture inputs."""
    from oracle import ref_autograd as O
    z, owl, params, groups, pinn, out, rows, total, grads = run_product(fixture)
    res = O.step(owl, [p.double() for p in params], {g: x.double() for g, x in groups.items()})
    for i, (g, r) in enumerate(zip(grads, res['grads'])):
        assert rel_l2(g, r) < TOL_LOSS, (i, rel_l2(g, r))
    flat = torch.cat([g.flatten() for g in grads])
    flat_ref = torch.cat([g.flatten() for g in res['grads']])
    assert rel_l2(flat, flat_ref) < TOL_LOSS
    assert max_rel(flat, flat_ref) < TOL_LOSS


def test_group_slices_and_all_are_consistent():
    z, owl, params, groups, pinn, out, rows, total, grads = run_product('ldc_small')
    lo = 0
    for g, x in groups.items():
        n = x.shape[0]
        for name in pinn.derivatives:
            assert torch.equal(out[g][name], out['all'][name][lo:lo + n])
        lo += n


def test_repeatable_bitwise():
    """Fixed tile->CTA map and ordered second-stage reduction: gradients are bitwise reproducible."""
    a = run_product('ldc_small')[-1]
    b = run_product('ldc_small')[-1]
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.parametrize('fixture', ['spring', 'burgers', 'allen_cahn', 'ldc_small'])
def test_fused_loss_kernel_matches_torch_path(fixture):
    """tde_residual_loss (rows + seeds) against the torch elementwise evaluation of the same loss table."""
    z, owl, params, groups = load(fixture)
    wl = product_workload(fixture)
    results = {}
    for fused in (True, False):
        net, lins = load_net(wl, params, 'cuda')
        pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
        batch = T.make_batch(wl.input_vars, {g: x.cuda() for g, x in groups.items()})
        out = pinn.calculate(batch)
        handler = wl.build_losses(list(groups.keys()))
        handler.fused = fused
        loss = handler(batch, out)
        total = wl.combine(loss.grouped_losses('loss_type'))
        total.backward()
        if fused:
            assert set(handler.fusion_report.values()) == {'fused'}, handler.fusion_report
            assert all(r['residual'] is None for r in loss.losses)
        results[fused] = (loss.individual_losses().detach().cpu(), float(total),
                          torch.cat([p.grad.flatten().cpu() for lin in lins for p in (lin.weight, lin.bias)]))
    rows_f, tot_f, g_f = results[True]
    rows_t, tot_t, g_t = results[False]
    assert torch.allclose(rows_f, rows_t, rtol=2e-5, atol=1e-12)
    assert abs(tot_f - tot_t) <= 2e-5 * abs(tot_t)
    assert rel_l2(g_f, g_t) < 2e-5


def test_unfusable_row_falls_back_to_torch_ops_for_its_group_only():
    z, owl, params, groups = load('burgers')
    wl = product_workload('burgers')
    net, lins = load_net(wl, params, 'cuda')
    pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
    batch = T.make_batch(wl.input_vars, {g: x.cuda() for g, x in groups.items()})
    out = pinn.calculate(batch)
    handler = wl.build_losses(list(groups.keys()))
    handler.add_residual('collocation points', {'exp_term': T.DE_func(lambda u, **kw: torch.exp(u) - 1)})
    loss = handler(batch, out)
    assert handler.fusion_report['collocation points'].startswith('torch path')
    assert handler.fusion_report['x0'] == 'fused'
    ref = wl.build_losses(list(groups.keys()))
    ref.add_residual('collocation points', {'exp_term': T.DE_func(lambda u, **kw: torch.exp(u) - 1)})
    ref.fused = False
    assert torch.allclose(loss.individual_losses(), ref(batch, out).individual_losses(), rtol=2e-5)
    loss.sum().backward()


def test_target_groups_value_only_for_other_groups():
    z, owl, params, groups = load('ldc_small')
    wl = product_workload('ldc_small')
    net, lins = load_net(wl, params, 'cuda')
    pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
    batch = T.make_batch(wl.input_vars, {g: x.cuda() for g, x in groups.items()})
    full = pinn.calculate(batch)
    part = pinn.calculate(batch, target_groups='collocation points')
    assert set(part['inlet'].keys()) == {'u', 'v', 'p'}                       # base.py:41-54 semantics
    assert set(part['collocation points'].keys()) == set(pinn.derivatives)
    for k in ('u', 'v', 'p'):
        assert torch.allclose(part['inlet'][k], full['inlet'][k], rtol=1e-6, atol=1e-7)
    for k in pinn.derivatives:
        assert torch.allclose(part['collocation points'][k], full['collocation points'][k], rtol=1e-6, atol=1e-6)
    (part['inlet']['u'].sum() + part['collocation points']['u_xx'].sum()).backward()
    assert all(p.grad is not None for p in net.parameters())


def test_sin_activation_and_plain_tensor_input():
    from oracle import ref_autograd as O
    params = O.init_mlp_params(2, 1, 64, 3, seed=5)
    net = T.MLP(2, 1, 64, 2, activation='sin')
    lins, _ = T.extract_layers(net)
    with torch.no_grad():
        for i, lin in enumerate(lins):
            lin.weight.copy_(params[2 * i]); lin.bias.copy_(params[2 * i + 1])
    net = net.cuda()
    pinn = T.DE_Getter(net, ['x', 't'], ['u'], ['u_xt', 'u_tt'])
    x = torch.rand(77, 2, generator=torch.Generator().manual_seed(3))
    out = pinn(x.cuda())['all']
    imap = O.derivative_index_map(['x', 't'], ['u'], ['u_xt', 'u_tt'])
    ref = O.streams([p.double() for p in params], x.double(), imap, 'sin')
    for k in imap:
        assert rel_l2(out[k].detach().cpu(), ref[k]) < TOL_STREAM, k
    loss = sum(v.pow(2).mean() for v in out.values())
    loss.backward()
    pr = [p.double().requires_grad_(True) for p in params]
    rref = O.streams(pr, x.double(), imap, 'sin')
    gref = torch.autograd.grad(sum(v.pow(2).mean() for v in rref.values()), pr)
    got = torch.cat([p.grad.flatten().cpu() for lin in lins for p in (lin.weight, lin.bias)])
    assert rel_l2(got, torch.cat([g.flatten() for g in gref])) < TOL_LOSS


def test_retain_graph_multiple_backwards_like_gradnorm():
    """utils/Loss_weighting.py:39: k x autograd.grad(loss_k, params, retain_graph=True, allow_unused=True) on one forward."""
    z, owl, params, groups = load('ldc_small')
    wl = product_workload('ldc_small')
    net, lins = load_net(wl, params, 'cuda')
    pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
    batch = T.make_batch(wl.input_vars, {g: x.cuda() for g, x in groups.items()})
    out = pinn.calculate(batch)
    loss = wl.build_losses(list(groups.keys()))(batch, out)
    rows = loss.individual_losses()
    ps = list(net.parameters())
    per_row = [torch.autograd.grad(r, ps, retain_graph=True, allow_unused=True) for r in rows]
    total = torch.autograd.grad(rows.sum(), ps)
    for i, g in enumerate(total):
        acc = sum(pr[i] for pr in per_row)
        assert rel_l2(acc.cpu(), g.cpu()) < 1e-5


def test_empty_and_ragged_batches():
    net = T.MLP(2, 1, 64, 2).cuda()
    pinn = T.DE_Getter(net, ['x', 't'], ['u'], ['u_x', 'u_xx'])
    out = pinn(torch.zeros(0, 2, device='cuda'))['all']
    assert out['u_xx'].shape == (0,)
    for n in (1, 31, 97, 129):                  # not multiples of the tile size
        x = torch.rand(n, 2, device='cuda')
        o = pinn(x)['all']
        assert o['u'].shape == (n,)
        assert torch.allclose(o['u'], net(x)[:, 0], rtol=1e-5, atol=1e-6)
        o['u_xx'].sum().backward()


def test_recompute_mode_matches_stash_mode(monkeypatch):
    """Batches whose stash would not fit switch to chunked recompute in backward; gradients must agree."""
    from torch_de_b200 import engine as E
    z, owl, params, groups = load('ldc_small')
    wl = product_workload('ldc_small')
    grads = {}
    for frac in (0.55, 1e-9):
        monkeypatch.setattr(E, 'STASH_FRACTION_OF_FREE', frac)
        net, lins = load_net(wl, params, 'cuda')
        pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
        x = torch.rand(5000, 2, generator=torch.Generator().manual_seed(1)).cuda() * 0.2 - 0.1
        out = pinn(x)['all']
        (out['u_xx'] * out['v_y'] + out['p_x'].pow(2)).mean().backward()
        grads[frac] = torch.cat([p.grad.flatten().cpu() for lin in lins for p in (lin.weight, lin.bias)])
    assert rel_l2(grads[1e-9], grads[0.55]) < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
# tensor-core (tcgen05, 3xTF32) path: every shape class it accepts, against the oracle in float64
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_case(d_in, n_out, width, n_act, ins, outs, ders, act, n_pts, seed):
    from oracle import ref_autograd as O
    params = O.init_mlp_params(d_in, n_out, width, n_act, seed=seed)
    net = T.MLP(d_in, n_out, width, n_act - 1, activation=act)
    lins, _ = T.extract_layers(net)
    with torch.no_grad():
        for i, lin in enumerate(lins):
            lin.weight.copy_(params[2 * i]); lin.bias.copy_(params[2 * i + 1])
    net = net.cuda()
    x = torch.rand(n_pts, d_in, generator=torch.Generator().manual_seed(seed + 1)) * 2 - 1
    imap = O.derivative_index_map(ins, outs, ders)
    pinn = T.DE_Getter(net, ins, outs, ders)
    out = pinn(x.cuda())['all']
    pr = [p.double().requires_grad_(True) for p in params]
    ref = O.streams(pr, x.double(), imap, act)
    weights = {k: 0.3 + 0.1 * i for i, k in enumerate(imap)}
    loss = sum(weights[k] * out[k].pow(2).mean() for k in imap)
    loss.backward()
    gref = torch.autograd.grad(sum(weights[k] * ref[k].pow(2).mean() for k in imap), pr)
    got = torch.cat([p.grad.flatten().cpu() for lin in lins for p in (lin.weight, lin.bias)])
    return pinn, out, ref, got, torch.cat([g.flatten() for g in gref])


@pytest.mark.parametrize('case', [
    dict(d_in=2, n_out=1, width=256, n_act=4, ins=['x', 't'], outs=['u'], ders=['u_x', 'u_t', 'u_xx'], act='tanh'),          # J=4
    dict(d_in=3, n_out=3, width=128, n_act=3, ins=['x', 'y', 't'], outs=['u', 'v', 'p'], ders=['u_t', 'u_xx', 'v_yy', 'p_x'], act='tanh'),   # J=6, 2 accumulator sets
    dict(d_in=1, n_out=1, width=256, n_act=3, ins=['t'], outs=['u'], ders=['u_t', 'u_tt'], act='sin'),                      # J=3, sin
    dict(d_in=2, n_out=2, width=128, n_act=5, ins=['x', 'y'], outs=['u', 'v'], ders=[], act='tanh'),                        # J=1 (value only)
    dict(d_in=2, n_out=1, width=200, n_act=3, ins=['x', 't'], outs=['u'], ders=['u_xt'], act='tanh'),                       # padded width, mixed -> J=6
    dict(d_in=2, n_out=3, width=256, n_act=3, ins=['x', 'y'], outs=['u', 'v', 'p'], ders=['u_xx', 'v_yy', 'p_x'], act='tanh'),
    # W = 512: 8-point tiles, four neuron blocks; odd J uses the four-product MMA scheme, even J the merged + single one
    dict(d_in=2, n_out=3, width=512, n_act=3, ins=['x', 'y'], outs=['u', 'v', 'p'], ders=['u_xx', 'v_yy', 'p_x'], act='tanh'),             # J=5
    dict(d_in=3, n_out=3, width=512, n_act=3, ins=['x', 'y', 't'], outs=['u', 'v', 'p'], ders=['u_t', 'u_xx', 'v_yy', 'p_x'], act='tanh'),  # J=6, 2 sets
    dict(d_in=2, n_out=1, width=512, n_act=4, ins=['x', 't'], outs=['u'], ders=['u_x', 'u_t', 'u_xx'], act='tanh'),                         # J=4, 3 sets
    dict(d_in=1, n_out=1, width=512, n_act=3, ins=['t'], outs=['u'], ders=['u_t', 'u_tt'], act='sin'),                                      # J=3, sin
    dict(d_in=2, n_out=2, width=512, n_act=3, ins=['x', 'y'], outs=['u', 'v'], ders=[], act='tanh'),                                        # J=1
    dict(d_in=2, n_out=3, width=400, n_act=3, ins=['x', 'y'], outs=['u', 'v', 'p'], ders=['u_xx', 'v_yy', 'p_x'], act='tanh'),             # padded to 512
])
@pytest.mark.parametrize('n_pts', [1, 17, 333, 5000])
def test_tensor_core_path_shapes(case, n_pts):
    pinn, out, ref, got, gref = _oracle_case(n_pts=n_pts, seed=21, **case)
    runner = pinn.deriv_method._full
    import ctypes as C
    from torch_de_b200 import _lib
    paths = _lib.lib().tde_plan_paths(C.byref(runner._net), C.byref(runner._spec), n_pts)
    assert paths == 3, f'expected the tcgen05 forward+backward path, got {paths}'
    # a single point makes the norm-relative metric pointwise-relative (meaningless near a zero crossing): looser bar there
    tol = TOL_STREAM if n_pts > 1 else 5 * TOL_STREAM
    for k in ref:
        assert rel_l2(out[k].detach().cpu(), ref[k].detach()) < tol, k
        assert max_rel(out[k].detach().cpu(), ref[k].detach()) < tol, k
    assert rel_l2(got, gref) < TOL_LOSS
    assert max_rel(got, gref) < TOL_LOSS


def test_tensor_core_and_ffma_paths_agree(monkeypatch):
    z, owl, params, groups = load('ldc')
    res = {}
    from torch_de_b200 import _lib
    for no_tc in ('1', '0'):
        _lib.set_option('no_tc', int(no_tc))
        try:
            res[no_tc] = run_product('ldc')
        finally:
            _lib.set_option('no_tc', 0)
    for k in res['0'][4].derivatives:
        assert rel_l2(res['0'][5]['all'][k].detach().cpu(), res['1'][5]['all'][k].detach().cpu()) < TOL_STREAM
    g0 = torch.cat([g.flatten() for g in res['0'][8]])
    g1 = torch.cat([g.flatten() for g in res['1'][8]])
    assert rel_l2(g0, g1) < 1e-5


def test_cta_pair_kernels_match_reference_and_single_cta(monkeypatch):
    """TDE_TC_PAIR=1 routes W = 256 through the cta_group::2 kernels (csrc/jet_pair_tc.cuh): same parity bar against the
    reference's float64 run, and agreement with the default single-CTA kernels."""
    z, owl, params, groups = load('ldc')
    res = {}
    from torch_de_b200 import _lib
    for pair in ('0', '1'):
        _lib.set_option('tc_pair', int(pair))
        try:
            res[pair] = run_product('ldc')
        finally:
            _lib.set_option('tc_pair', 0)
    for k in res['1'][4].derivatives:
        a, b = res['1'][5]['all'][k].detach().cpu(), res['0'][5]['all'][k].detach().cpu()
        assert rel_l2(a, b) < TOL_STREAM
        assert rel_l2(a, z['s64_' + k]) < TOL_STREAM, (k, rel_l2(a, z['s64_' + k]))
    g1 = torch.cat([g.flatten() for g in res['1'][8]])
    g0 = torch.cat([g.flatten() for g in res['0'][8]])
    assert rel_l2(g1, g0) < 1e-5
    # and it is bitwise repeatable (slab copies per thread, fixed-order reduction)
    _lib.set_option('tc_pair', 1)
    try:
        again = run_product('ldc')
    finally:
        _lib.set_option('tc_pair', 0)
    for x, y in zip(res['1'][8], again[8]):
        assert torch.equal(x, y)


def test_tensor_core_recompute_mode(monkeypatch):
    from torch_de_b200 import engine as E
    grads = {}
    for frac in (0.7, 1e-9):
        monkeypatch.setattr(E, 'STASH_FRACTION_OF_FREE', frac)
        torch.manual_seed(3)
        net = T.MLP(2, 1, 128, 3).cuda()
        pinn = T.DE_Getter(net, ['x', 't'], ['u'], ['u_x', 'u_xx'])
        x = torch.rand(3000, 2, generator=torch.Generator().manual_seed(1)).cuda()
        out = pinn(x)['all']
        (out['u_xx'] * out['u_x'] + out['u'].pow(2)).mean().backward()
        grads[frac] = torch.cat([p.grad.flatten().cpu() for p in net.parameters()])
    assert rel_l2(grads[1e-9], grads[0.7]) < 1e-5


def test_tensor_core_gradients_bitwise_repeatable():
    a = run_product('ldc')[-1]
    b = run_product('ldc')[-1]
    for x, y in zip(a, b):
        assert torch.equal(x, y)


# ---------------------------------------------------------------------------------------------------------------------
# RWF_Linear + Fourier_Embedding / Fourier_Net (Networks.py:25-75,104-120; AllenCahn.py:65, Stokes_Cylinder.py:84)
# ---------------------------------------------------------------------------------------------------------------------
def test_fourier_net_rwf_matches_reference_golden(golden_dir):
    import os
    z = np.load(os.path.join(golden_dir, 'allen_cahn_fourier.npz'))
    wl = WL.allen_cahn()
    names = [str(n) for n in z['param_names']]
    groups = {str(g): torch.from_numpy(z['x_' + str(g)]) for g in z['group_names']}
    net = T.Fourier_Net(2, 1, 100, 4, RWF=True)
    assert [n for n, _ in net.named_parameters()] == names          # same registration order as the reference
    with torch.no_grad():
        for n, p_ in net.named_parameters():
            p_.copy_(torch.from_numpy(z['p_' + n]))
    net = net.cuda()
    pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
    batch = T.make_batch(wl.input_vars, {g: x.cuda() for g, x in groups.items()})
    out = pinn.calculate(batch)
    loss = wl.build_losses(list(groups.keys()))(batch, out)
    total = wl.combine(loss.grouped_losses('loss_type'))
    total.backward()
    for k in pinn.derivatives:
        got = out['all'][k].detach().cpu()
        assert rel_l2(got, z['s64_' + k]) < TOL_STREAM, (k, rel_l2(got, z['s64_' + k]))
        assert max_rel(got, z['s64_' + k]) < TOL_STREAM, k
    assert abs(float(total) - float(z['total64'])) <= TOL_LOSS * abs(float(z['total64']))
    for n, p_ in net.named_parameters():
        assert rel_l2(p_.grad.cpu(), z['g64_' + n]) < TOL_LOSS, (n, rel_l2(p_.grad.cpu(), z['g64_' + n]))


@pytest.mark.parametrize('cfg', [dict(d=2, m=3, w=128, h=3, act='sin', rwf=True), dict(d=3, m=3, w=64, h=2, act='tanh', rwf=True),
                                 dict(d=2, m=1, w=256, h=2, act='tanh', rwf=False)])
def test_fourier_net_variants_match_oracle(cfg):
    from oracle import ref_autograd as O
    torch.manual_seed(5)
    net = T.Fourier_Net(cfg['d'], cfg['m'], cfg['w'], cfg['h'], activation=cfg['act'], RWF=cfg['rwf'], fourier_std=1.0)
    ins, outs = ['x', 'y', 't'][:cfg['d']], ['u', 'v', 'p'][:cfg['m']]
    ders = ['u_x', 'u_xx'] + (['u_y', 'u_yy'] if cfg['d'] >= 2 else [])
    names = [n for n, _ in net.named_parameters()]
    leaves = [p_.detach().clone().double() for _, p_ in net.named_parameters()]
    x = torch.rand(97, cfg['d'], generator=torch.Generator().manual_seed(2))
    imap = O.derivative_index_map(ins, outs, ders)
    lv = [t.requires_grad_(True) for t in leaves]
    ref = O.assign_derivs(O.autodiff_all(lambda zz: O.fourier_net_apply(dict(zip(names, lv)), zz, cfg['act'], cfg['h'], cfg['rwf']),
                                         x.double(), 2), imap)
    net = net.cuda()
    pinn = T.DE_Getter(net, ins, outs, ders)
    out = pinn(x.cuda())['all']
    for k in imap:
        assert rel_l2(out[k].detach().cpu(), ref[k].detach()) < TOL_STREAM, k
    sum(v.pow(2).mean() for v in out.values()).backward()
    gref = torch.autograd.grad(sum(v.pow(2).mean() for v in ref.values()), lv)
    for (n, p_), g in zip(net.named_parameters(), gref):
        assert rel_l2(p_.grad.cpu(), g) < TOL_LOSS, n


def test_rwf_mlp_matches_plain_mlp_with_same_effective_weights():
    torch.manual_seed(9)
    rwf = T.MLP(2, 1, 64, 2, RWF=True).cuda()
    plain = T.MLP(2, 1, 64, 2).cuda()
    with torch.no_grad():
        for a, b in zip(list(rwf.layers) + [rwf.linear_out], list(plain.layers) + [plain.linear_out]):
            b.weight.copy_(a.v * a.s); b.bias.copy_(a.bias)
    x = torch.rand(50, 2, device='cuda')
    o1 = T.DE_Getter(rwf, ['x', 't'], ['u'], ['u_xx'])(x)['all']
    o2 = T.DE_Getter(plain, ['x', 't'], ['u'], ['u_xx'])(x)['all']
    for k in o1:
        assert torch.allclose(o1[k], o2[k], rtol=1e-6, atol=1e-7)
    o1['u_xx'].sum().backward(); o2['u_xx'].sum().backward()
    for a, b in zip(list(rwf.layers) + [rwf.linear_out], list(plain.layers) + [plain.linear_out]):
        assert rel_l2((a.v.grad * a.s).cpu() , (b.weight.grad * a.s * a.s).cpu()) < 1e-5     # dL/dv = dL/dW * s
        assert rel_l2(a.s.grad.cpu(), (b.weight.grad * a.v).sum(1, keepdim=True).cpu()) < 1e-5


@pytest.mark.parametrize('fixture', ['ldc_small', 'ldc', 'burgers'])
def test_gradnorm_weights_follow_reference_rule(fixture):
    """Fused GradNorm (per-row sweeps over the row's own group, tde_jet_backward_range) against the reference's rule evaluated
    with k full autograd passes on the same forward.  ldc: tensor-core path; the groups are ragged, so every range starts and
    ends inside a tile."""
    z, owl, params, groups = load(fixture)
    wl = product_workload(fixture)
    net, lins = load_net(wl, params, 'cuda')
    pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
    batch = T.make_batch(wl.input_vars, {g: x.cuda() for g, x in groups.items()})
    rows = wl.build_losses(list(groups.keys()))(batch, pinn.calculate(batch)).individual_losses()
    w0 = torch.ones(len(rows), device='cuda')
    w1 = T.GradNorm(net, w0, *rows, alpha=0.9)
    assert T.GradNorm.last_path == 'fused'
    ps = list(net.parameters())
    mags = torch.stack([torch.sqrt(sum(g.pow(2).sum() for g in torch.autograd.grad(r, ps, retain_graph=True, allow_unused=True) if g is not None))
                        for r in rows])
    expect = 0.9 * w0 + 0.1 * (mags.sum() / mags)
    assert torch.allclose(w1, expect, rtol=1e-4)
    # a plain tensor that does not come from the loss table takes the k-pass rule
    T.GradNorm(net, w0[:1], rows.sum())
    assert T.GradNorm.last_path == 'k-pass'
    # like autograd's saved-tensor check: a forward whose weights were changed in place afterwards cannot be swept again
    with torch.no_grad():
        next(net.parameters()).add_(0.0)
    with pytest.raises(RuntimeError, match='modified in place'):
        T.GradNorm(net, w0, *rows)


def test_gradnorm_fused_path_at_scale_and_through_rwf_layers():
    """(a) 20 011 + 3 x 1 003 ragged points on the 8x256 tensor path; (b) a Fourier_Net with RWF layers: the kernel's d/dW is
    chained to the (v, s) leaves."""
    from oracle import ref_autograd as O
    owl = O.WORKLOADS['ldc']()
    wl = WL.ldc()
    groups = O.sample_groups(owl, 20_011, 1_003, seed=5)
    torch.manual_seed(3)
    for net in (wl.make_net().cuda(), T.Fourier_Net(2, 3, 128, 3, RWF=True).cuda()):
        pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
        batch = T.make_batch(wl.input_vars, {g: x.cuda() for g, x in groups.items()})
        rows = wl.build_losses(list(groups.keys()))(batch, pinn.calculate(batch)).individual_losses()
        w0 = torch.ones(len(rows), device='cuda')
        w1 = T.GradNorm(net, w0, *rows, alpha=0.5)
        assert T.GradNorm.last_path == 'fused'
        ps = list(net.parameters())
        mags = torch.stack([torch.sqrt(sum(g.pow(2).sum() for g in torch.autograd.grad(r, ps, retain_graph=True, allow_unused=True) if g is not None))
                            for r in rows])
        expect = 0.5 * w0 + 0.5 * (mags.sum() / mags)
        assert torch.allclose(w1, expect, rtol=2e-4), (w1, expect)


def test_graphed_step_replays_the_eager_step_bitwise():
    """torch_de_b200.GraphedStep: the whole training step (calculate -> fused losses -> backward -> Adam) captured in one CUDA
    graph gives bit-identical losses and parameters to the eager step (Burgers 8x64: launch-bound)."""
    z, owl, params, groups = load('burgers')
    wl = product_workload('burgers')
    results = {}
    for mode in ('eager', 'graph'):
        net, lins = load_net(wl, params, 'cuda')
        pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
        batch = T.make_batch(wl.input_vars, {g: x.to('cuda') for g, x in groups.items()})
        handler = wl.build_losses(list(groups.keys()))
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)

        def step():
            opt.zero_grad(set_to_none=True)
            out = pinn.calculate(batch)
            total = wl.combine(handler(batch, out).grouped_losses('loss_type'))
            total.backward()
            opt.step()
            return total

        if mode == 'graph':
            step = T.GraphedStep(step, warmup=2)           # 2 warm-up steps + 1 captured (not run) ...
            losses = [float(step()) for _ in range(3)]
        else:
            for _ in range(2):
                step()                                      # ... = the same 2 steps before the 3 compared ones
            losses = [float(step()) for _ in range(3)]
        results[mode] = (losses, torch.cat([p.detach().flatten().cpu() for p in net.parameters()]))
    assert results['eager'][0] == results['graph'][0]
    assert torch.equal(results['eager'][1], results['graph'][1])


def test_no_reference_cycle_keeps_the_stash_alive():
    """A training step must free its stream block and activation stash by reference counting alone (the cycle collector runs
    whenever it likes: at LDC scale a leaked step is 21 GB)."""
    import gc
    z, owl, params, groups = load('ldc_small')
    wl = product_workload('ldc_small')
    net, lins = load_net(wl, params, 'cuda')
    pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
    batch = T.make_batch(wl.input_vars, {g: x.cuda() for g, x in groups.items()})
    handler = wl.build_losses(list(groups.keys()))

    def step():
        out = pinn.calculate(batch)
        loss = handler(batch, out)
        loss.point_error()                      # materialises the per-point pass too
        total = wl.combine(loss.grouped_losses('loss_type'))
        total.backward()
        return float(total)

    gc.collect()
    gc.disable()
    try:
        step()
        torch.cuda.synchronize()
        base = torch.cuda.memory_allocated()
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        assert torch.cuda.memory_allocated() <= base + (1 << 16), (torch.cuda.memory_allocated(), base)
    finally:
        gc.enable()
