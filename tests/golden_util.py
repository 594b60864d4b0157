"""Helpers shared by the parity tests: load a golden fixture and rebuild its exact inputs."""
import os

import numpy as np
import torch

from oracle import ref_autograd as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ['spring', 'burgers', 'allen_cahn', 'ldc_small', 'ldc', 'cylinder', 'cylinder_unsteady']


def factory(fixture):
    if fixture == 'ldc_small':
        return O._ldc(width=64, n_act=4, name='ldc_small')
    return O.WORKLOADS[fixture]()


def load(fixture):
    z = np.load(os.path.join(GOLDEN, fixture + '.npz'))
    wl = factory(fixture)
    n = wl.net
    params = O.init_mlp_params(n['d_in'], n['n_out'], n['width'], n['n_act'], seed=int(z['param_seed']))
    if 'p_0' in z:
        for i, p in enumerate(params):      # the seeded initialiser must reproduce the stored weights bit for bit
            assert np.array_equal(p.numpy(), z[f'p_{i}']), 'nn.Linear init under the stored seed changed'
    groups = {str(g): torch.from_numpy(z['x_' + str(g)]) for g in z['group_names']}
    return z, wl, params, groups


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).flatten()
    b = torch.as_tensor(b, dtype=torch.float64).flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def max_rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).flatten()
    b = torch.as_tensor(b, dtype=torch.float64).flatten()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
