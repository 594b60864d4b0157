"""Pin the oracle (oracle/ref_autograd.py) against outputs of the unmodified reference (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from oracle import ref_autograd as O
from tests.golden_util import FIXTURES, load, max_rel, rel_l2


@pytest.mark.parametrize('fixture', FIXTURES)
def test_index_map_matches_reference(fixture):
    z, wl, _, _ = load(fixture)
    imap = wl.index_map()
    assert list(imap.keys()) == [str(n) for n in z['names']]
    assert [str(v) for v in imap.values()] == [str(v) for v in z['index_tuples']]


@pytest.mark.parametrize('fixture', FIXTURES)
def test_oracle_step_matches_reference(fixture):
    z, wl, params, groups = load(fixture)
    torch.set_num_threads(8)
    res = O.step(wl, params, groups)
    # streams: same torch.func calls as the reference; allow last-bit differences between CPUs
    for name, s in res['streams'].items():
        assert rel_l2(s, z['s32_' + name]) < 2e-6, name
        assert max_rel(s, z['s64_' + name]) < 2e-5, name
    ref_rows = dict(zip([str(r) for r in z['row_names']], z['rows32']))
    assert set(ref_rows) == set(res['row_names'])
    for n, v in zip(res['row_names'], res['rows']):
        assert abs(float(v) - ref_rows[n]) <= 2e-6 * abs(ref_rows[n]) + 1e-12, n
    assert abs(float(res['total']) - float(z['total32'])) <= 2e-6 * abs(float(z['total32']))
    for i, g in enumerate(res['grads']):
        if f'g32_{i}' in z:
            assert rel_l2(g, z[f'g32_{i}']) < 5e-6, i
    if 'gproj32' in z:
        flat = torch.cat([g.flatten() for g in res['grads']]).double()
        gen = torch.Generator().manual_seed(99)
        R = torch.randn(64, flat.numel(), generator=gen, dtype=torch.float64) / flat.numel() ** 0.5
        assert rel_l2(R @ flat, z['gproj64']) < 1e-5
        assert abs(float(flat.norm()) - float(z['gflat64_norm'])) < 1e-5 * float(z['gflat64_norm'])


def test_double_backward_agrees_with_vmap_jacrev():
    """autograd.grad(create_graph=True) is an independent route to the same numbers."""
    wl = O.WORKLOADS['burgers']()
    params = O.init_mlp_params(2, 1, 64, 8, seed=3)
    x = torch.rand(50, 2, generator=torch.Generator().manual_seed(5))
    a = O.streams(params, x, wl.index_map())
    b = O.streams_double_backward(params, x, wl.index_map())
    for k in a:
        assert rel_l2(a[k], b[k]) < 1e-6, k


def test_closed_form_single_neuron():
    """u = tanh(w x + b): u_x = w(1-u^2), u_xx = -2 w^2 u (1-u^2)   (SURVEY.md section 4 (i))."""
    w, b, v, c = 0.7, -0.2, 1.3, 0.1
    params = [torch.tensor([[w]]), torch.tensor([b]), torch.tensor([[v]]), torch.tensor([c])]
    x = torch.linspace(-2, 2, 41).unsqueeze(-1)
    s = O.streams(params, x, O.derivative_index_map(['x'], ['u'], ['u_xx']))
    t = torch.tanh(w * x[:, 0] + b)
    assert torch.allclose(s['u'], v * t + c, atol=1e-6)
    assert torch.allclose(s['u_x'], v * w * (1 - t * t), atol=1e-6)
    assert torch.allclose(s['u_xx'], v * (-2 * w * w * t * (1 - t * t)), atol=1e-6)


def test_oracle_fourier_net_matches_reference(golden_dir):
    """Fourier_Net(2,1,100,4,RWF=True) on the Allen-Cahn loss table (Tutorials/AllenCahn.py:65)."""
    import os
    z = np.load(os.path.join(golden_dir, 'allen_cahn_fourier.npz'))
    wl = O.WORKLOADS['allen_cahn']()
    names = [str(n) for n in z['param_names']]
    groups = {str(g): torch.from_numpy(z['x_' + str(g)]) for g in z['group_names']}
    leaves = [torch.from_numpy(z['p_' + n]) for n in names]
    res = O.step_net(wl, lambda lv, x: O.fourier_net_apply(dict(zip(names, lv)), x, 'tanh', 4, True), leaves, groups)
    for k, s in res['streams'].items():
        assert rel_l2(s, z['s32_' + k]) < 5e-6, k
    assert abs(float(res['total']) - float(z['total32'])) <= 5e-6 * abs(float(z['total32']))
    for n, g in zip(names, res['grads']):
        assert rel_l2(g, z['g32_' + n]) < 2e-5, n
