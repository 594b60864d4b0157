"""Fused residual assembly + loss reduction (north-star subsystem 2).

The reference evaluates every residual as a chain of torch elementwise kernels on `[B]` tensors, then
abs/pow/mul/mean per loss row (utils/loss.py:33-59) and lets autograd walk it back.  Here each group's rows
are traced ONCE into a small register program (`Sym` proxies flow through the user's own `DE_func` /
rhs callables), and one kernel (`tde_residual_loss`) evaluates all rows of a group per point, reduces
`mean(w*|r|^2)` and -- in backward -- emits d(total)/d(stream) seeds for the reverse jet sweep.
Rows that cannot be traced (unsupported torch calls, non-integer powers, power != 2 ...) stay on the torch
elementwise path for their group; nothing ever runs on the CPU.
"""
import ctypes as C
import numbers
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib


class TraceError(Exception):
    pass


class Sym:
    """Expression node. Hash-consed by structure so common sub-expressions (and loads) are shared."""
    __slots__ = ('op', 'args', 'imm', 'key')
    _BIN = {'add': _lib.OP_ADD, 'sub': _lib.OP_SUB, 'mul': _lib.OP_MUL, 'div': _lib.OP_DIV}
    _UN = {'neg': _lib.OP_NEG, 'abs': _lib.OP_ABS, 'sin': _lib.OP_SIN, 'cos': _lib.OP_COS, 'powi': _lib.OP_POWI}
    # 'grid2d': (x, y) -> bilinear lookup in the grid registered under slot `imm` (RegularGridInterpolator called on Syms)

    def __init__(self, op, args=(), imm=None):
        self.op, self.args, self.imm = op, tuple(args), imm
        self.key = (op, tuple(a.key for a in self.args), imm)

    @staticmethod
    def lift(v):
        if isinstance(v, Sym):
            return v
        if isinstance(v, numbers.Real):
            return Sym('const', imm=float(v))
        if isinstance(v, torch.Tensor) and v.numel() == 1:
            # A tensor scalar (an annealed Re, a PDE coefficient of an inverse problem ...) is NOT baked into the program: a
            # learnable one needs autograd (torch path), any other one is re-read from the tensor at every launch through an
            # aux row, so a value changed in place between steps is seen.
            if v.requires_grad:
                raise TraceError('a learnable scalar in a residual needs the torch path (its gradient)')
            if _TRACE is None:
                raise TraceError('tensor scalar outside a trace')
            return Sym('aux', imm=_TRACE.scalar_slot(v))
        raise TraceError(f'cannot mix {type(v).__name__} into a traced residual')

    def _bin(self, other, op, swap=False):
        o = Sym.lift(other)
        return Sym(op, (o, self) if swap else (self, o))

    def __add__(self, o): return self._bin(o, 'add')
    def __radd__(self, o): return self._bin(o, 'add', True)
    def __sub__(self, o): return self._bin(o, 'sub')
    def __rsub__(self, o): return self._bin(o, 'sub', True)
    def __mul__(self, o): return self._bin(o, 'mul')
    def __rmul__(self, o): return self._bin(o, 'mul', True)
    def __truediv__(self, o): return self._bin(o, 'div')
    def __rtruediv__(self, o): return self._bin(o, 'div', True)
    def __neg__(self): return Sym('neg', (self,))
    def __pos__(self): return self
    def __abs__(self): return Sym('abs', (self,))

    def __pow__(self, e):
        if isinstance(e, numbers.Real) and float(e) == int(e) and abs(int(e)) <= 16:
            return Sym('powi', (self,), imm=float(int(e)))
        raise TraceError('only small integer powers can be fused')

    def abs(self): return abs(self)
    def pow(self, e): return self ** e
    def square(self): return self ** 2
    def sin(self): return Sym('sin', (self,))
    def cos(self): return Sym('cos', (self,))

    def __bool__(self):
        raise TraceError('data-dependent control flow cannot be fused')

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        name = getattr(func, '__name__', '')
        table = {'abs': lambda a: abs(Sym.lift(a)), 'sin': lambda a: Sym.lift(a).sin(), 'cos': lambda a: Sym.lift(a).cos(),
                 'neg': lambda a: -Sym.lift(a), 'square': lambda a: Sym.lift(a) ** 2,
                 'add': lambda a, b: Sym.lift(a) + b, 'sub': lambda a, b: Sym.lift(a) - b,
                 'mul': lambda a, b: Sym.lift(a) * b, 'div': lambda a, b: Sym.lift(a) / b,
                 'true_divide': lambda a, b: Sym.lift(a) / b, 'pow': lambda a, b: Sym.lift(a) ** b}
        if name in table and not kwargs:
            return table[name](*args)
        raise TraceError(f'torch.{name} has no fused op')

    def loads(self):
        out, stack, seen = set(), [self], set()
        while stack:
            n = stack.pop()
            if n.key in seen:
                continue
            seen.add(n.key)
            if n.op == 'load':
                out.add(n.imm)
            stack.extend(n.args)
        return out


class _TraceContext:
    """Tensor scalars met while tracing one group: each becomes the aux leaf '__s<k>' that is re-read every launch.
    Grid interpolators called on traced coordinates get a slot in the program's grid table."""
    def __init__(self):
        self.scalars: List[torch.Tensor] = []
        self.grids: list = []

    def grid_slot(self, interp) -> int:
        for k, g in enumerate(self.grids):
            if g is interp:
                return k
        if len(self.grids) >= _lib.MAX_GRIDS:
            raise TraceError('too many distinct grid fields in one group')
        self.grids.append(interp)
        return len(self.grids) - 1

    def scalar_slot(self, t: torch.Tensor) -> str:
        for k, u in enumerate(self.scalars):
            if u is t:
                return f'__s{k}'
        self.scalars.append(t)
        return f'__s{len(self.scalars) - 1}'


_TRACE: Optional[_TraceContext] = None


class _GroupProxy:
    """Stands in for a PINN_group while tracing: batchable columns become aux leaves."""
    def __init__(self, group):
        self.name = getattr(group, 'name', None)
        self.batchables = {k: Sym('aux', imm=k) for k in group.batchables if k != 'input'}
        self.unbatchables = dict(getattr(group, 'unbatchables', {}))
        self.inputs = {k: v for k, v in self.batchables.items() if k in getattr(group, 'input_vars', [])}


class _Guard:
    """Value that must not be used: detects weight functions that read network outputs."""
    def _boom(self, *a, **k):
        raise TraceError('weighting depends on network outputs')
    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = _boom
    __neg__ = __abs__ = __pow__ = __rpow__ = __float__ = __bool__ = _boom

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        raise TraceError('weighting depends on network outputs')


class GroupProgram:
    """Compiled rows of one group."""
    def __init__(self, group_name: str, row_ids: List[int], res_syms: List[Sym], stream_row: Dict[str, int], n_stream_rows: int,
                 weight_syms: Optional[List[Optional[Sym]]] = None):
        self.group = group_name
        self.row_ids = row_ids
        self.aux_names: List[str] = []
        self.n_stream_rows = n_stream_rows
        prog = _lib.TdeLossProgram()
        reg_of: Dict[tuple, int] = {}
        ops: List[Tuple[int, int, int, int, float]] = []

        def emit(node: Sym) -> int:
            if node.key in reg_of:
                return reg_of[node.key]
            if node.op == 'const':
                rec = (_lib.OP_CONST, 0, 0, node.imm)
            elif node.op == 'load':
                if node.imm not in stream_row:
                    raise TraceError(f"'{node.imm}' is not a stream of this engine")
                rec = (_lib.OP_LOAD, stream_row[node.imm], 0, 0.0)
            elif node.op == 'aux':
                rec = (_lib.OP_AUX, self._aux(node.imm), 0, 0.0)
            elif node.op in Sym._BIN:
                a, b = emit(node.args[0]), emit(node.args[1])
                rec = (Sym._BIN[node.op], a, b, 0.0)
            elif node.op == 'grid2d':
                a, b = emit(node.args[0]), emit(node.args[1])
                rec = (_lib.OP_GRID2D, a, b, node.imm)
            else:
                a = emit(node.args[0])
                rec = (Sym._UN[node.op], a, 0, node.imm or 0.0)
            r = len(ops)
            if r >= _lib.MAX_REGS or r >= _lib.MAX_OPS:
                raise TraceError('residual program too large to fuse')
            ops.append((rec[0], r, rec[1], rec[2], rec[3]))
            reg_of[node.key] = r
            return r

        if len(res_syms) > _lib.MAX_ROWS:
            raise TraceError('too many rows in one group')
        for i, s in enumerate(res_syms):
            prog.res_reg[i] = emit(s)
            prog.w_aux[i] = -1
            prog.w_reg[i] = -1
            prog.w_const[i] = 1.0
            prog.coef[i] = 1.0
        self.traced_weight = [False] * len(res_syms)      # rows whose weighting function runs inside the kernel
        for i, w in enumerate(weight_syms or []):
            if w is not None:
                prog.w_reg[i] = emit(w)
                self.traced_weight[i] = True
        prog.n_ops, prog.n_rows = len(ops), len(res_syms)
        for k, (op, dst, a, b, imm) in enumerate(ops):
            prog.ops[k].op, prog.ops[k].dst, prog.ops[k].a, prog.ops[k].b, prog.ops[k].imm = op, dst, a, b, imm
        self.prog = prog
        self.weight_aux: List[Optional[int]] = [None] * len(res_syms)
        self.scalars: List[torch.Tensor] = []          # tensor scalars behind the '__s<k>' aux rows
        self.grids: list = []                          # interpolators behind the GRID2D slots
        self._grid_keep = []                           # device tensors the grid table points at
        self.signature = None                          # set by trace_group: what the rows traced to (see signature_of)

    def bind_grids(self, device) -> None:
        """(Re)fill the program's grid table with device pointers (the interpolator may have been moved / re-valued)."""
        self.prog.n_grids = len(self.grids)
        keep = []
        for k, g in enumerate(self.grids):
            gx, gy = (p.to(device=device, dtype=torch.float32).contiguous() for p in g.points)
            vals = g.values.to(device=device, dtype=torch.float32).contiguous()
            keep += [gx, gy, vals]
            e = self.prog.grids[k]
            e.gx, e.gy, e.values, e.nx, e.ny = gx.data_ptr(), gy.data_ptr(), vals.data_ptr(), gx.shape[0], gy.shape[0]
        self._grid_keep = keep

    def _aux(self, name: str) -> int:
        if name not in self.aux_names:
            self.aux_names.append(name)
        return self.aux_names.index(name)

    def weight_slot(self, i: int) -> int:
        """Reserve an aux row for a per-point weight of row i."""
        slot = self._aux(f'__w{i}')
        self.prog.w_aux[i] = slot
        self.weight_aux[i] = slot
        return slot


def trace_group(group, rows: List[dict], stream_row: Dict[str, int], n_stream_rows: int) -> GroupProgram:
    """rows: Loss_handler records of ONE group (non-custom). Raises TraceError if any row cannot be fused."""
    syms, wsyms, ctx = trace_rows(group, rows, stream_row)
    prog = GroupProgram(getattr(group, 'name', None), list(range(len(rows))), syms, stream_row, n_stream_rows, wsyms)
    prog.scalars = ctx.scalars
    prog.grids = ctx.grids
    prog.signature = _signature(syms, wsyms, ctx)
    return prog


def _signature(syms, wsyms, ctx) -> tuple:
    return (tuple(s.key for s in syms), tuple(w.key if w is not None else None for w in wsyms), tuple(id(g) for g in ctx.grids))


def grid_lookup(interp, x, y) -> Sym:
    """RegularGridInterpolator(x, y) on traced coordinates: one GRID2D op instead of a separate lookup launch + aux row."""
    if _TRACE is None:
        raise TraceError('grid lookup outside a trace')
    return Sym('grid2d', (Sym.lift(x), Sym.lift(y)), imm=float(_TRACE.grid_slot(interp)))


def trace_rows(group, rows: List[dict], stream_row: Dict[str, int]):
    """Symbolic evaluation of the rows' own callables (pure Python, no tensor work): the Sym of every row + the trace context."""
    global _TRACE
    gp = _GroupProxy(group)
    outs = {name: Sym('load', imm=name) for name in stream_row}
    prev, _TRACE = _TRACE, _TraceContext()
    try:
        syms = [Sym.lift(spec['evaluation'](gp, outs)) for spec in rows]
        wsyms = [trace_weight(spec, gp, outs) for spec in rows]
        return syms, wsyms, _TRACE
    finally:
        _TRACE = prev


def trace_weight(spec: dict, gp, outs) -> Optional[Sym]:
    """The row's weighting function as a Sym of the group's input coordinates (an SDF lookup, `1 - 10 |x|`, ...), or None:
    constants, per-point tensors and anything that does not trace keep the aux-row / constant path (eval_weight)."""
    w = spec.get('weighting_spec', 1)
    if not callable(w):
        return None
    guard = {k: _Guard() for k in outs}
    try:
        val = spec['weighting_function'](gp, guard)
    except TraceError:
        return None
    except Exception:       # noqa: BLE001 - a weighting written for tensors only (torch.where, indexing ...): evaluate it with torch
        return None
    return val if isinstance(val, Sym) and val.op != 'const' else None


def signature_of(group, rows: List[dict], stream_row: Dict[str, int]) -> tuple:
    """Signature of what the rows trace to NOW: differs from a cached program's when a Python number captured by a residual
    closure (or a group's unbatchable) changed since that program was compiled."""
    return _signature(*trace_rows(group, rows, stream_row))


def eval_weight(spec: dict, group, out_names) -> object:
    """Per-point weight of a row, evaluated with torch on the real inputs (weights are constants of the step)."""
    w = spec.get('weighting_spec', 1)
    if not callable(w):
        if isinstance(w, torch.Tensor) and w.numel() > 1:      # a fixed per-point weight tensor
            return w.detach().to(torch.float32)
        return float(w)
    guard = {k: _Guard() for k in out_names}
    with torch.no_grad():
        val = spec['weighting_function'](group, guard)
    if isinstance(val, torch.Tensor) and val.numel() > 1:
        return val.to(torch.float32)
    return float(val)


class _FusedRows(torch.autograd.Function):
    """rows[n_rows_total] = per-row mean(w*r^2) of every fused group; backward emits the stream seeds."""

    @staticmethod
    def forward(ctx, block: torch.Tensor, launches):
        B = block.shape[1]
        n_total = sum(gp.prog.n_rows for gp, *_ in launches)
        rows = torch.zeros(n_total, dtype=torch.float32, device=block.device)
        lib = _lib.lib()
        stream = torch.cuda.current_stream(block.device).cuda_stream
        off = 0
        with _lib.on_device(block.device):
            for gp, aux, lo, n in launches:
                if n > 0:
                    rc = lib.tde_residual_loss(C.byref(gp.prog), block.data_ptr(), B, gp.n_stream_rows,
                                               aux.data_ptr() if aux is not None else None, n, lo, n, 1.0 / n, None,
                                               rows.data_ptr() + 4 * off, None, None, stream)
                    _lib.check(rc, 'tde_residual_loss')
                off += gp.prog.n_rows
        ctx.launches = launches
        ctx.save_for_backward(block)
        return rows

    @staticmethod
    def backward(ctx, grad_rows):
        (block,) = ctx.saved_tensors
        B = block.shape[1]
        covered = sum(n for _, _, _, n in ctx.launches)
        seeds = torch.empty_like(block) if covered == B else torch.zeros_like(block)
        grad_rows = grad_rows.contiguous().to(torch.float32)
        lib = _lib.lib()
        stream = torch.cuda.current_stream(block.device).cuda_stream
        off = 0
        with _lib.on_device(block.device):
            for gp, aux, lo, n in ctx.launches:
                if n > 0:
                    rc = lib.tde_residual_loss(C.byref(gp.prog), block.data_ptr(), B, gp.n_stream_rows,
                                               aux.data_ptr() if aux is not None else None, n, lo, n, 1.0 / n,
                                               grad_rows.data_ptr() + 4 * off, None, seeds.data_ptr(), None, stream)
                    _lib.check(rc, 'tde_residual_loss')
                off += gp.prog.n_rows
        return seeds, None


def fused_rows(block: torch.Tensor, launches) -> torch.Tensor:
    return _FusedRows.apply(block, launches)


def point_rows(block: torch.Tensor, gp: GroupProgram, aux, lo: int, n: int):
    """(residuals[n_rows, n], weights[n_rows, n]) of one fused group: the per-point quantities `Loss.point_error()` /
    `weighted_point_error()` and `R3_sampler` read (utils/loss.py:33-52, utils/sampling.py:70-112).  One forward-only launch
    of the loss kernel; nothing is kept for autograd (these feed samplers and diagnostics)."""
    res = torch.empty((gp.prog.n_rows, n), dtype=torch.float32, device=block.device)
    wts = torch.empty_like(res)
    if n > 0:
        with _lib.on_device(block.device):
            rc = _lib.lib().tde_residual_loss_points(C.byref(gp.prog), block.data_ptr(), block.shape[1], gp.n_stream_rows,
                                                     aux.data_ptr() if aux is not None else None, n, lo, n, res.data_ptr(),
                                                     wts.data_ptr(), torch.cuda.current_stream(block.device).cuda_stream)
        _lib.check(rc, 'tde_residual_loss_points')
    return res, wts
