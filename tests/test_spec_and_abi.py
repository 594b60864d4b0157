"""CPU tests: derivative-spec handling, channel canonicalisation, and the C-ABI library (load, exports, planning,
error behaviour) -- no kernel is launched."""
import ctypes as C
import os
import re

import pytest
import torch

import torch_de_b200 as T
from torch_de_b200 import _lib
from torch_de_b200.spec import COMPILED_SPECS, JetChannels, derivative_index_map
from oracle import ref_autograd as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('ins,outs,ders', [
    (['t'], ['u'], ['u_t', 'u_tt']),
    (['x', 't'], ['u'], ['u_x', 'u_t', 'u_xx']),
    (['x', 't'], ['u'], ['u_t', 'u_xx']),
    (['x', 't'], ['u'], ['u_xt', 'u_tt']),
    (['x', 'y'], ['u', 'v', 'p'], ['u_x', 'u_y', 'u_xx', 'u_yy', 'p_x', 'v_y', 'v_x', 'v_xx', 'v_yy', 'p_y']),
    (['x', 'y', 't'], ['u', 'v', 'p'], ['u_t', 'u_xx', 'v_yy', 'p_x']),
])
def test_index_map_matches_oracle(ins, outs, ders):
    assert derivative_index_map(ins, outs, ders) == O.derivative_index_map(ins, outs, ders)
    assert list(derivative_index_map(ins, outs, ders)) == list(O.derivative_index_map(ins, outs, ders))


def test_index_map_errors_like_reference():
    with pytest.raises(AssertionError, match='Output Variable w does not exist'):
        derivative_index_map(['x'], ['u'], ['w_x'])
    with pytest.raises(AssertionError, match='is not an input Variable'):
        derivative_index_map(['x'], ['u'], ['u_y'])


def test_channel_layout_ldc():
    imap = derivative_index_map(['x', 'y'], ['u', 'v', 'p'], ['u_x', 'u_y', 'u_xx', 'u_yy', 'p_x', 'v_y', 'v_x', 'v_xx', 'v_yy', 'p_y'])
    ch = JetChannels(imap, 2, 3)
    assert (ch.n_first, ch.pairs, ch.n_channels, ch.n_rows) == (2, [(0, 0), (1, 1)], 5, 15)
    assert ch.row('u') == 0 and ch.row('p') == 2
    assert ch.row('v_x') == (1 + ch.first_in.index(0)) * 3 + 1
    assert ch.row('u_yy') == (3 + ch.pairs.index((ch.first_in.index(1),) * 2)) * 3 + 0


def test_channel_layout_permutes_to_a_compiled_spec():
    # only u_tt on (x,t): first-order channel set {t}; (1,{00}) is compiled
    ch = JetChannels(derivative_index_map(['x', 't'], ['u'], ['u_tt']), 2, 1)
    assert ch.first_in == [1] and ch.pairs == [(0, 0)] and ch.mask in COMPILED_SPECS[1]
    # u_x, u_tt: needs pair on t only -> t must become channel 0 of the (2,{00}) spec
    ch = JetChannels(derivative_index_map(['x', 't'], ['u'], ['u_x', 'u_tt']), 2, 1)
    assert ch.first_in == [1, 0] and ch.pairs == [(0, 0)]
    # mixed derivative widens to the full d=2 Hessian spec
    ch = JetChannels(derivative_index_map(['x', 't'], ['u'], ['u_xt']), 2, 1)
    assert ch.mask == 7 and ch.n_channels == 6
    assert ch.row('u_xt') == (1 + 2 + ch.pairs.index((0, 1))) * 1
    # value only
    ch = JetChannels({'u': (0,), 'v': (1,)}, 3, 2)
    assert ch.n_channels == 1 and ch.row('v') == 1


def test_order_three_is_refused():
    with pytest.raises(NotImplementedError):
        JetChannels(derivative_index_map(['x'], ['u'], ['u_xxx']), 1, 1)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'tde_jet.h')).read()
    declared = set(re.findall(r'\b(tde_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.tde_version() >= 100


def _runner(d, m, w, h, ders, ins, outs):
    net = T.MLP(d, m, w, h) if h > 0 else torch.nn.Sequential(torch.nn.Linear(d, w), torch.nn.Tanh(), torch.nn.Linear(w, m))
    return T.DE_Getter(net, ins, outs, ders).deriv_method._full


def test_plan_geometry_without_a_device():
    r = _runner(2, 3, 256, 7, ['u_xx', 'u_yy', 'v_xx', 'v_yy', 'p_x', 'p_y'], ['x', 'y'], ['u', 'v', 'p'])
    p = r.plan(500_000)
    assert p.n_channels == 5 and p.width_padded == 256 and p.n_params == 462_083
    assert p.n_tiles == -(-500_000 // p.tile_points)
    assert p.smem_forward <= 232448 and p.smem_backward <= 232448
    exact = 8 * p.n_tiles * 256 * p.tile_points * 5 * 4
    pad = 8 * 3 * 16 * 5 * 256 * 4     # the tensor-core kernels may pad the last cluster's tiles (at most 3 tiles of 16 points per layer)
    # tensor-core backward: the forward also keeps the post-activation jets (second half of the stash buffer)
    assert exact <= p.stash_bytes <= exact + pad or 2 * exact <= p.stash_bytes <= 2 * (exact + pad)
    r = _runner(1, 1, 200, 0, ['u_t', 'u_tt'], ['t'], ['u'])
    p = r.plan(1001)
    assert p.width_padded == 256 and p.n_params == 601 and p.n_channels == 3
    assert r.plan(0).n_tiles == 0


def test_abi_rejects_bad_specs_with_a_message():
    lib = _lib.lib()
    r = _runner(2, 1, 64, 3, ['u_x'], ['x', 't'], ['u'])
    r.plan(10)
    spec = _lib.TdeJetSpec()
    C.memmove(C.byref(spec), C.byref(r._spec), C.sizeof(spec))
    plan = _lib.TdePlan()
    spec.first_in[0] = 5
    assert lib.tde_plan(C.byref(r._net), C.byref(spec), 10, C.byref(plan)) != 0
    assert b'not an input column' in lib.tde_last_error()
    spec.first_in[0] = 0
    spec.n_first, spec.n_second = 2, 1
    spec.first_in[1] = 1
    spec.second_a[0], spec.second_b[0] = 0, 1            # only the mixed pair: not a compiled mask
    assert lib.tde_plan(C.byref(r._net), C.byref(spec), 10, C.byref(plan)) != 0
    assert b'not compiled in' in lib.tde_last_error()
    r._net.layers[1].out_features = 32                   # ragged hidden widths
    assert lib.tde_plan(C.byref(r._net), C.byref(r._spec), 10, C.byref(plan)) != 0
    assert b'hidden widths' in lib.tde_last_error()


def test_unsupported_networks_raise_not_fallback():
    with pytest.raises(NotImplementedError):
        T.extract_layers(torch.nn.Sequential(torch.nn.Linear(1, 8), torch.nn.ReLU(), torch.nn.Linear(8, 1)))
    with pytest.raises(NotImplementedError):
        T.extract_layers(torch.nn.LSTM(2, 2))


def test_engine_has_no_cpu_path():
    pinn = T.DE_Getter(T.MLP(2, 1, 64, 2), ['x', 't'], ['u'], ['u_x'])
    with pytest.raises(RuntimeError, match='CUDA'):
        pinn.calculate(torch.rand(4, 2))


def test_abi_argument_checks_of_the_round_two_entry_points():
    """tde_jet_backward_range / tde_residual_loss_points / GRID2D programs: bad arguments come back as error codes with a
    message, before anything is launched (so these run without a device)."""
    lib = _lib.lib()
    r = _runner(2, 1, 64, 3, ['u_x'], ['x', 't'], ['u'])
    r.plan(100)
    one = C.c_void_p(16)          # any non-null pointer: the checks below fail before it is touched
    for lo, hi in ((-1, 10), (50, 40), (0, 101)):
        rc = lib.tde_jet_backward_range(C.byref(r._net), C.byref(r._spec), one, 100, one, one, 100, lo, hi, 1.0, one, one, None)
        assert rc != 0 and b'range' in lib.tde_last_error()
    assert lib.tde_jet_backward_range(C.byref(r._net), C.byref(r._spec), one, 100, one, one, 100, 32, 32, 1.0, one, one, None) == 0   # empty range
    prog = _lib.TdeLossProgram()
    prog.n_ops, prog.n_rows = 1, 1
    prog.ops[0].op, prog.ops[0].dst, prog.ops[0].a = _lib.OP_LOAD, 0, 0
    prog.res_reg[0], prog.w_aux[0], prog.w_reg[0] = 0, -1, -1
    assert lib.tde_residual_loss_points(C.byref(prog), one, 8, 1, None, 0, 0, 8, None, None, None) != 0
    assert b'nothing to compute' in lib.tde_last_error()
    prog.n_ops = 2
    prog.ops[1].op, prog.ops[1].dst, prog.ops[1].a, prog.ops[1].b, prog.ops[1].imm = _lib.OP_GRID2D, 1, 0, 0, 0.0
    assert lib.tde_residual_loss_points(C.byref(prog), one, 8, 1, None, 0, 0, 8, one, None, None) != 0
    assert b'grid 0 out of range' in lib.tde_last_error()         # the program names a grid it does not carry
    prog.n_grids = 1
    assert lib.tde_residual_loss_points(C.byref(prog), one, 8, 1, None, 0, 0, 8, one, None, None) != 0
    assert b'null pointer or empty axis' in lib.tde_last_error()
    prog.n_grids = 0
    prog.n_ops = 1
    prog.w_reg[0] = _lib.MAX_REGS
    assert lib.tde_residual_loss_points(C.byref(prog), one, 8, 1, None, 0, 0, 8, one, None, None) != 0
    assert b'weight register' in lib.tde_last_error()
