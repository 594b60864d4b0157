"""ctypes binding of csrc/libtde_jet.so (C ABI in include/tde_jet.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (TDE_LIB_PATH: load another build of the same library -- A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("TDE_LIB_PATH") or os.path.join(_HERE, "csrc", "libtde_jet.so")

ACT_NONE, ACT_TANH, ACT_SIN = 0, 1, 2
ACT_FOURIER = 0x100
MAX_LAYERS, MAX_IN, MAX_OUT, MAX_FIRST, MAX_SECOND = 24, 8, 4, 3, 6
MAX_REGS, MAX_OPS, MAX_ROWS = 64, 128, 8
(OP_CONST, OP_LOAD, OP_AUX, OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_POWI, OP_ABS, OP_SIN, OP_COS, OP_DIV, OP_GRID2D) = range(13)
MAX_GRIDS = 2


class TdeLayer(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("in_features", C.c_int32),
                ("out_features", C.c_int32), ("act", C.c_int32), ("reserved", C.c_int32)]


class TdeNet(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("reserved", C.c_int32), ("layers", TdeLayer * MAX_LAYERS)]


class TdeJetSpec(C.Structure):
    _fields_ = [("d_in", C.c_int32), ("n_out", C.c_int32), ("n_first", C.c_int32), ("n_second", C.c_int32),
                ("first_in", C.c_int32 * MAX_FIRST), ("second_a", C.c_int32 * MAX_SECOND),
                ("second_b", C.c_int32 * MAX_SECOND)]


class TdePlan(C.Structure):
    _fields_ = [("n_channels", C.c_int32), ("width_padded", C.c_int32), ("tile_points", C.c_int32),
                ("n_ctas", C.c_int32), ("n_tiles", C.c_int64), ("stash_bytes", C.c_int64),
                ("workspace_bytes", C.c_int64), ("n_params", C.c_int64), ("smem_forward", C.c_int32),
                ("smem_backward", C.c_int32), ("workspace_forward_bytes", C.c_int64)]


class TdeOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("dst", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("imm", C.c_float)]


class TdeGrid2D(C.Structure):
    _fields_ = [("gx", C.c_void_p), ("gy", C.c_void_p), ("values", C.c_void_p), ("nx", C.c_int32), ("ny", C.c_int32)]


class TdeLossProgram(C.Structure):
    _fields_ = [("n_ops", C.c_int32), ("n_rows", C.c_int32), ("ops", TdeOp * MAX_OPS),
                ("res_reg", C.c_int32 * MAX_ROWS), ("w_aux", C.c_int32 * MAX_ROWS),
                ("w_const", C.c_float * MAX_ROWS), ("coef", C.c_float * MAX_ROWS),
                ("w_reg", C.c_int32 * MAX_ROWS), ("n_grids", C.c_int32), ("reserved", C.c_int32),
                ("grids", TdeGrid2D * MAX_GRIDS)]


EXPORTS = ["tde_last_error", "tde_version", "tde_abi_struct_bytes", "tde_plan", "tde_jet_forward", "tde_jet_backward",
           "tde_residual_loss", "tde_adam_step", "tde_fp32_peak_probe", "tde_launch_count", "tde_profile_begin",
           "tde_profile_end", "tde_profile_kind", "tde_plan_paths", "tde_grid_interp2d", "tde_set_option",
           "tde_jet_backward_range", "tde_residual_loss_points"]

_LIB = None


def _declare(lib):
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float
    lib.tde_last_error.restype = C.c_char_p
    lib.tde_last_error.argtypes = []
    lib.tde_version.restype = C.c_int
    lib.tde_plan.argtypes = [C.POINTER(TdeNet), C.POINTER(TdeJetSpec), i64, C.POINTER(TdePlan)]
    lib.tde_jet_forward.argtypes = [C.POINTER(TdeNet), C.POINTER(TdeJetSpec), vp, i64, vp, i64, vp, vp, vp]
    lib.tde_jet_backward.argtypes = [C.POINTER(TdeNet), C.POINTER(TdeJetSpec), vp, i64, vp, vp, i64, f32, vp, vp, vp]
    lib.tde_jet_backward_range.argtypes = [C.POINTER(TdeNet), C.POINTER(TdeJetSpec), vp, i64, vp, vp, i64, i64, i64, f32, vp, vp, vp]
    lib.tde_residual_loss.argtypes = [C.POINTER(TdeLossProgram), vp, i64, i32, vp, i64, i64, i64, f32, vp, vp, vp, vp, vp]
    lib.tde_residual_loss_points.argtypes = [C.POINTER(TdeLossProgram), vp, i64, i32, vp, i64, i64, i64, vp, vp, vp]
    lib.tde_adam_step.argtypes = [vp, vp, vp, vp, i64, C.c_double, C.c_double, C.c_double, C.c_double, i32, vp]
    lib.tde_fp32_peak_probe.argtypes = [C.POINTER(C.c_double), vp]
    lib.tde_profile_begin.argtypes = []
    lib.tde_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    for name in EXPORTS[3:]:
        getattr(lib, name).restype = C.c_int
    lib.tde_profile_kind.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    lib.tde_plan_paths.argtypes = [C.POINTER(TdeNet), C.POINTER(TdeJetSpec), i64]
    lib.tde_grid_interp2d.argtypes = [vp, i32, vp, i32, vp, vp, vp, i64, vp, vp]
    lib.tde_set_option.argtypes = [C.c_char_p, C.c_int]
    lib.tde_launch_count.restype = C.c_longlong
    lib.tde_launch_count.argtypes = []
    return lib


def lib():
    """The loaded CUDA library. Raises if it has not been built (run __graft_entry__.build())."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(torch_de_b200 has no CPU or eager fallback)")
        handle = _declare(C.CDLL(LIB_PATH))
        handle.tde_abi_struct_bytes.argtypes = [C.c_int]
        handle.tde_abi_struct_bytes.restype = C.c_int
        for i, st in enumerate((TdeNet, TdeJetSpec, TdePlan, TdeLossProgram)):
            if handle.tde_abi_struct_bytes(i) != C.sizeof(st):
                raise RuntimeError(f"{LIB_PATH} was built from a different include/tde_jet.h ({st.__name__}: "
                                   f"{handle.tde_abi_struct_bytes(i)} vs {C.sizeof(st)} bytes): rebuild it")
        _LIB = handle
    return _LIB


def set_option(name: str, value: int) -> None:
    """Override a run-time switch of the library (see tde_set_option in include/tde_jet.h)."""
    check(lib().tde_set_option(name.encode(), int(value)), "tde_set_option")


def on_device(device):
    """Context manager: make `device` the process's current CUDA device around a library call.  The kernels, the
    cudaFuncSetAttribute / occupancy queries and the SM-count lookup inside libtde_jet.so all act on the current device,
    which is not necessarily the device the tensors live on (several GPUs in one process, DP ranks that only use
    `.to(f'cuda:{rank}')`)."""
    import contextlib
    import torch
    if getattr(device, 'type', None) == 'cuda':
        return torch.cuda.device(device)
    return contextlib.nullcontext()


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib().tde_last_error().decode()}")
