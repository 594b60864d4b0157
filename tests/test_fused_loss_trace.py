"""CPU tests of the residual tracer/compiler: the register program produced for each workload, interpreted in
numpy on the reference's golden streams, must reproduce the reference's loss rows."""
import math

import numpy as np
import pytest
import torch

import torch_de_b200 as T
from torch_de_b200 import _lib
from torch_de_b200 import fused_loss as FL
from torch_de_b200 import workloads as WL
from tests.golden_util import FIXTURES, load


def interpret(gp, streams, aux):
    """numpy model of residual_loss_kernel's forward sweep. streams: [rows, n], aux: [n_aux, n]."""
    prog = gp.prog
    v = {}
    for k in range(prog.n_ops):
        op = prog.ops[k]
        a, b = op.a, op.b
        if op.op == _lib.OP_CONST: r = np.full(streams.shape[1], op.imm, dtype=np.float64)
        elif op.op == _lib.OP_LOAD: r = streams[a].astype(np.float64)
        elif op.op == _lib.OP_AUX: r = aux[a].astype(np.float64)
        elif op.op == _lib.OP_ADD: r = v[a] + v[b]
        elif op.op == _lib.OP_SUB: r = v[a] - v[b]
        elif op.op == _lib.OP_MUL: r = v[a] * v[b]
        elif op.op == _lib.OP_NEG: r = -v[a]
        elif op.op == _lib.OP_POWI: r = v[a] ** int(op.imm)
        elif op.op == _lib.OP_ABS: r = np.abs(v[a])
        elif op.op == _lib.OP_SIN: r = np.sin(v[a])
        elif op.op == _lib.OP_COS: r = np.cos(v[a])
        elif op.op == _lib.OP_DIV: r = v[a] / v[b]
        else: raise AssertionError(op.op)
        assert op.dst == k, 'program must be in SSA form (dst == op index)'
        v[op.dst] = r
    out = []
    for i in range(prog.n_rows):
        w = aux[prog.w_aux[i]].astype(np.float64) if prog.w_aux[i] >= 0 else prog.w_const[i]
        out.append(np.mean(w * v[prog.res_reg[i]] ** 2))
    return out


def product_workload(fixture):
    return WL.ldc(width=64, n_act=4, name='ldc_small') if fixture == 'ldc_small' else WL.WORKLOADS[fixture]()


@pytest.mark.parametrize('fixture', FIXTURES)
def test_traced_programs_reproduce_reference_rows(fixture):
    z, owl, params, groups = load(fixture)
    wl = product_workload(fixture)
    names = [str(n) for n in z['names']]
    stream_row = {n: i for i, n in enumerate(names)}
    streams_all = np.stack([z['s64_' + n] for n in names])
    losses = wl.build_losses(list(groups.keys()))
    batch = T.make_batch(wl.input_vars, groups)
    ref = dict(zip([str(r) for r in z['row_names']], z['rows64']))
    lo = 0
    for g, x in groups.items():
        n = x.shape[0]
        specs = [r for r in losses.rows if r['group'] == g]
        gp = FL.trace_group(batch[g], specs, stream_row, len(names))
        weights = [FL.eval_weight(sp, batch[g], names) for sp in specs]
        for i, w in enumerate(weights):
            if isinstance(w, torch.Tensor):
                gp.weight_slot(i)
            else:
                gp.prog.w_const[i] = w
        aux = []
        for a in gp.aux_names:
            aux.append(weights[int(a[3:])].numpy() if a.startswith('__w') else batch[g].batchables[a].numpy())
        aux = np.stack(aux) if aux else np.zeros((0, n))
        got = interpret(gp, streams_all[:, lo:lo + n], aux)
        for sp, val in zip(specs, got):
            key = f"{sp['loss_type']}/{sp['group']}/{sp['variable']}"
            assert val == pytest.approx(float(ref[key]), rel=5e-6, abs=1e-14), key
        lo += n


def test_common_subexpressions_and_loads_are_shared():
    u, ux = FL.Sym('load', imm='u'), FL.Sym('load', imm='u_x')
    e1 = u * ux + 2.0 * u
    e2 = u * ux - 1.0
    gp = FL.GroupProgram('g', [0, 1], [e1, e2], {'u': 0, 'u_x': 1}, 2)
    ops = [gp.prog.ops[k].op for k in range(gp.prog.n_ops)]
    assert ops.count(_lib.OP_LOAD) == 2 and ops.count(_lib.OP_MUL) == 2      # u*ux once, 2*u once


def test_untraceable_pieces_raise_trace_error():
    u = FL.Sym('load', imm='u')
    with pytest.raises(FL.TraceError):
        torch.exp(u)
    with pytest.raises(FL.TraceError):
        u ** 0.5
    with pytest.raises(FL.TraceError):
        bool(u)
    with pytest.raises(FL.TraceError):
        FL.GroupProgram('g', [0], [FL.Sym('load', imm='nope')], {'u': 0}, 1)
    # a weighting that reads network outputs cannot be treated as a constant of the step
    L = T.Loss_handler(['g'])
    L.add_residual('g', {'r': T.DE_func(lambda u, **kw: u)}, weighting=T.DE_func(lambda u, **kw: u * 2))
    grp = T.PINN_group('g', torch.zeros(4, 1), ['x'])
    with pytest.raises(FL.TraceError):
        FL.eval_weight(L.rows[0], grp, ['u'])


def test_tensor_scalars_are_reread_and_learnable_ones_refuse_to_fuse():
    """A tensor scalar in a residual (annealed Re, inverse-problem coefficient) is never baked into the program."""
    grp = T.PINN_group('g', torch.zeros(4, 1), ['x'])
    nu = torch.tensor(0.25)
    L = T.Loss_handler(['g'])
    L.add_residual('g', {'r': T.DE_func(lambda u, u_x, **kw: u_x - nu * u)})
    gp = FL.trace_group(grp, L.rows, {'u': 0, 'u_x': 1}, 2)
    assert gp.aux_names == ['__s0'] and gp.scalars[0] is nu                 # an aux leaf, not an OP_CONST
    assert all(gp.prog.ops[k].op != _lib.OP_CONST for k in range(gp.prog.n_ops))
    streams = np.array([[1.0, 2.0, 3.0, 4.0], [0.5, 0.5, 0.5, 0.5]])
    for val in (0.25, 0.125):                                                 # changed in place between "steps"
        nu.fill_(val)
        aux = np.stack([gp.scalars[0].reshape(1).expand(4).numpy()])
        got = interpret(gp, streams, aux)[0]
        assert got == pytest.approx(np.mean((0.5 - val * streams[0]) ** 2))
    learn = torch.nn.Parameter(torch.tensor(0.1))
    L2 = T.Loss_handler(['g'])
    L2.add_residual('g', {'r': T.DE_func(lambda u, **kw: learn * u)})
    with pytest.raises(FL.TraceError):
        FL.trace_group(grp, L2.rows, {'u': 0}, 1)


def test_signature_sees_a_changed_python_constant():
    grp = T.PINN_group('g', torch.zeros(4, 1), ['x'])
    box = {'re': 100.0}
    L = T.Loss_handler(['g'])
    L.add_residual('g', {'r': T.DE_func(lambda u, u_x, **kw: u_x - u / box['re'])})
    gp = FL.trace_group(grp, L.rows, {'u': 0, 'u_x': 1}, 2)
    assert FL.signature_of(grp, L.rows, {'u': 0, 'u_x': 1}) == gp.signature
    box['re'] = 200.0
    assert FL.signature_of(grp, L.rows, {'u': 0, 'u_x': 1}) != gp.signature


def test_fixed_per_point_weight_tensor_is_a_weight_row():
    grp = T.PINN_group('g', torch.zeros(4, 1), ['x'])
    w = torch.tensor([1.0, 2.0, 3.0, 4.0])
    L = T.Loss_handler(['g'])
    L.add_residual('g', {'r': T.DE_func(lambda u, **kw: u)}, weighting=w)
    got = FL.eval_weight(L.rows[0], grp, ['u'])
    assert isinstance(got, torch.Tensor) and torch.equal(got, w)
