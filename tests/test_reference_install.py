"""The reference shipped to the GPU box (`baseline/_ref`, tools/install_ref.py) is the UNMODIFIED source tree, and the oracle
port (oracle/ref_autograd.py) agrees with it when both run live on the CPU."""
import hashlib
import json
import os

import pytest
import torch

from oracle import ref_autograd as O
from oracle import ref_live as RL
from tests.golden_util import rel_l2
from tools import install_ref

needs_reference = pytest.mark.skipif(not RL.available(), reason='baseline/_ref (tools/install_ref.py) is not in this tree')


@needs_reference
def test_materialised_reference_is_byte_identical_to_its_manifest_and_source():
    dest = install_ref.install()
    manifest = json.load(open(os.path.join(dest, 'MANIFEST.json')))
    assert len(manifest['files']) > 30
    for rel, sha in manifest['files'].items():
        with open(os.path.join(dest, 'torch_DE', rel), 'rb') as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == sha, rel
        src = os.path.join(install_ref.REFERENCE_SRC, rel)
        if os.path.isdir(install_ref.REFERENCE_SRC):            # build container only
            with open(src, 'rb') as fh:
                assert hashlib.sha256(fh.read()).hexdigest() == sha, rel


@needs_reference
@pytest.mark.parametrize('name', ['burgers', 'cylinder_unsteady'])
def test_live_reference_step_matches_the_port(name):
    owl = O.WORKLOADS[name]() if name == 'burgers' else O._cylinder_unsteady(width=32, n_act=3)
    n = owl.net
    params = O.init_mlp_params(n['d_in'], n['n_out'], n['width'], n['n_act'], seed=3)
    groups = O.sample_groups(owl, 64, 8, seed=5)
    torch.set_num_threads(4)
    streams, imap, row_names, rows, total, grads = RL.run_reference(owl, params, groups, torch.float32)
    port = O.step(owl, params, groups)
    assert list(imap.keys()) == list(owl.index_map().keys())
    assert [tuple(v) for v in imap.values()] == list(owl.index_map().values())
    for k, v in port['streams'].items():
        assert rel_l2(v, streams[k]) < 2e-6, k
    assert sorted(row_names) == sorted(port['row_names'])
    got = dict(zip(port['row_names'], port['rows']))
    for nme, v in zip(row_names, rows):
        assert abs(float(got[nme]) - float(v)) <= 2e-6 * abs(float(v)) + 1e-12
    assert abs(float(port['total']) - float(total)) <= 2e-6 * abs(float(total))
    for a, b in zip(port['grads'], grads):
        assert rel_l2(a, b) < 1e-5


@needs_reference
def test_jet_engine_is_an_instance_of_the_reference_engine_base():
    import torch_de_b200 as T
    R = RL.activate()
    net = R.MLP(2, 1, 64, 2)
    pinn = R.DE_Getter(net, ['x', 't'], ['u'], ['u_x', 'u_xx'])
    eng = T.Jet_engine(net, pinn.derivatives)
    assert isinstance(eng, R.engine) and isinstance(eng, T.Jet_engine)
    pinn.set_deriv_method(eng)                                   # DE_module.py:148-149 accepts it
    assert pinn.deriv_method is eng
    fd = T.FD_engine(net, pinn.derivatives, dxs=[1e-3, 1e-3])
    assert isinstance(fd, R.engine)
    with pytest.raises(RuntimeError):
        pinn.calculate(torch.zeros(4, 2))                        # CPU tensor: no fallback
