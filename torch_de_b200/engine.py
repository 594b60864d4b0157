"""`Jet_engine`: the drop-in derivative engine behind torch_DE's own plugin point.

The reference accepts any `engine` instance in `DE_Getter.set_deriv_method(obj)`
(src/torch_DE/continuous/DE_module.py:148-149) and forwards `calculate(x, **kw)` to it (:173).
`Jet_engine` implements that contract -- `engine.calculate(x, target_groups=None)` with the grouping
semantics of `AD_engine.calculate/group_output/assign_derivs` (Engines/AD.py:42-68, 92-131) and of
`engine.net_pass_from_dict` (Engines/base.py:41-54) -- on top of the sm_100a jet kernels.

Returned tensors are FP32 `[B_g]` views of one `[J*n_out, B]` block, differentiable w.r.t. the
network parameters through `_JetFunction` (forward = tde_jet_forward, backward = tde_jet_backward).
"""
import ctypes as C
import sys
from typing import Dict, List, Optional, Sequence, Union

import torch

from . import _lib
from .networks import extract_layers
from .spec import JetChannels

class _EngineBase:  # same public surface as the reference's engine base (Engines/base.py:5-11)
    def __init__(self) -> None:
        self.derivatives = None
        self.net = None
        self.output_vars = None

    def __call__(self, *args, **kwargs):
        return self.calculate(*args, **kwargs)

    def __new__(cls, *args, **kwargs):
        """The reference installs a custom engine only if `isinstance(obj, torch_DE...engine)` (DE_module.py:148-149).  When
        the reference package is loaded in this process, instances are created from a subclass that also derives from its
        `engine`, whatever the import order was."""
        mod = sys.modules.get('torch_DE.continuous.Engines')
        ref = getattr(mod, 'engine', None) if mod is not None else None
        if isinstance(ref, type) and not issubclass(cls, ref):
            key = (cls, ref)
            if key not in _MIXED:
                _MIXED[key] = type(cls.__name__, (cls, ref), {'__module__': cls.__module__})
            cls = _MIXED[key]
        return object.__new__(cls)


_MIXED: Dict[tuple, type] = {}

# CPU tests flip this to exercise the host-side checks without a device; the product never does (no CPU path exists)
_REQUIRE_CUDA = True


class JetRunner:
    """Binds one network + one channel layout to the C ABI: builds the TdeNet/TdeJetSpec structs,
    sizes and caches stash/workspace, launches forward / backward."""

    def __init__(self, lins, act: int, channels: JetChannels):
        self.lins = list(lins)                  # LayerSpec per Linear-like layer
        self.act = act
        self.ch = channels
        self.n_params = sum(sp.in_features * sp.out_features + sp.out_features for sp in self.lins)
        self._spec = _lib.TdeJetSpec()
        s = self._spec
        s.d_in, s.n_out, s.n_first, s.n_second = channels.d_in, channels.n_out, channels.n_first, len(channels.pairs)
        for i, v in enumerate(channels.first_in):
            s.first_in[i] = v
        for i, (a, b) in enumerate(channels.pairs):
            s.second_a[i], s.second_b[i] = a, b
        self._net = _lib.TdeNet()
        self._net.n_layers = len(self.lins)
        self._plans: Dict[int, _lib.TdePlan] = {}
        self._workspace: Optional[torch.Tensor] = None
        self._last_forward = None
        self._chunk_stash_owner = None
        self._chunk_stash: Optional[torch.Tensor] = None
        self.chunking: Dict[tuple, int] = {}

    def _bind(self, tensors: Sequence[torch.Tensor]):
        n = len(self.lins)
        for l, lin in enumerate(self.lins):
            w, b = tensors[2 * l], tensors[2 * l + 1]
            if w.dtype != torch.float32 or not w.is_contiguous() or not b.is_contiguous():
                raise TypeError("jet kernels need contiguous float32 parameters")
            ly = self._net.layers[l]
            ly.weight, ly.bias = w.data_ptr(), b.data_ptr()
            ly.in_features, ly.out_features = lin.in_features, lin.out_features
            ly.act = (self.act | (_lib.ACT_FOURIER if lin.fourier else 0)) if l < n - 1 else _lib.ACT_NONE

    def plan(self, n_points: int) -> _lib.TdePlan:
        p = self._plans.get(n_points)
        if p is None:
            with torch.no_grad():
                self._bind(self.effective_tensors())
            p = _lib.TdePlan()
            with _lib.on_device(self.lins[0].module.parameters().__next__().device):
                _lib.check(_lib.lib().tde_plan(C.byref(self._net), C.byref(self._spec), n_points, C.byref(p)), "tde_plan")
            self._plans[n_points] = p
        return p

    def effective_tensors(self) -> List[torch.Tensor]:
        """[W0, b0, W1, b1, ...] as the kernels consume them (autograd-tracked functions of the leaf parameters)."""
        out = []
        for sp in self.lins:
            w, b = sp.tensors()
            out += [w.contiguous(), b.contiguous()]
        return out

    def _ws(self, nbytes: int, device) -> torch.Tensor:
        if self._workspace is None or self._workspace.numel() * 4 < nbytes or self._workspace.device != device:
            self._workspace = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
        return self._workspace

    @staticmethod
    def _stream(device):
        return torch.cuda.current_stream(device).cuda_stream if device.type == 'cuda' else 0

    def _check_x(self, x: torch.Tensor) -> torch.Tensor:
        if _REQUIRE_CUDA and not x.is_cuda:
            raise RuntimeError("Jet_engine runs on CUDA tensors only (sm_100a kernels; there is no CPU fallback)")
        if x.dim() != 2 or x.shape[1] != self.ch.d_in:
            raise ValueError(f"input must be [B, {self.ch.d_in}], got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            raise TypeError("input must be float32")
        for p in self.lins[0].module.parameters():
            if p.device != x.device:
                raise RuntimeError(f"network parameters live on {p.device} but the input is on {x.device}")
        return x.contiguous()

    def forward(self, x: torch.Tensor, tensors: Sequence[torch.Tensor], need_stash: bool, reuse_stash: bool = False,
                out: Optional[torch.Tensor] = None, out_col: int = 0, owner=None):
        """reuse_stash: keep the stash in a buffer owned by the runner (recompute mode: it only lives until the matching
        backward of the same chunk, and re-allocating ~100 GB blocks every chunk fragments the caching allocator); `owner`
        tags who may still rely on that buffer's contents (`_chunk_stash_owner`).
        out / out_col: write the streams into columns [out_col, out_col + B) of an existing [rows, B_total] block."""
        x = self._check_x(x)
        B = x.shape[0]
        if out is None:
            out = torch.empty((self.ch.n_rows, B), dtype=torch.float32, device=x.device)
            out_col = 0
        elif out.dim() != 2 or out.shape[0] != self.ch.n_rows or out_col < 0 or out_col + B > out.shape[1] or not out.is_contiguous():
            raise ValueError("output block does not hold these columns")
        if B == 0:
            return x, out, None
        plan = self.plan(B)
        self._bind(tensors)
        stash = None
        if need_stash and reuse_stash:
            if self._chunk_stash is None or self._chunk_stash.numel() < plan.stash_bytes // 4 or self._chunk_stash.device != x.device:
                self._chunk_stash = None
                self._chunk_stash = torch.empty(plan.stash_bytes // 4, dtype=torch.float32, device=x.device)
            stash = self._chunk_stash
            self._chunk_stash_owner = owner
        elif need_stash:
            stash = torch.empty(plan.stash_bytes // 4, dtype=torch.float32, device=x.device)
        # a forward that keeps no stash (inference, value-only groups, the first pass of recompute mode) only needs the
        # packed weights; the full workspace (gradient slabs, Hbar stream) belongs to a forward/backward pair
        ws = self._ws(plan.workspace_bytes if need_stash else plan.workspace_forward_bytes, x.device)
        with _lib.on_device(x.device):
            rc = _lib.lib().tde_jet_forward(C.byref(self._net), C.byref(self._spec), x.data_ptr(), B,
                                            out.data_ptr() + 4 * out_col, out.shape[1],
                                            stash.data_ptr() if need_stash else None, ws.data_ptr(), self._stream(x.device))
        _lib.check(rc, "tde_jet_forward")
        return x, out, stash

    def backward(self, x: torch.Tensor, tensors: Sequence[torch.Tensor], stash: torch.Tensor, seeds: torch.Tensor,
                 flat_grad: Optional[torch.Tensor] = None, grad_scale: float = 1.0, seeds_ld: int = 0,
                 point_range: Optional[tuple] = None) -> torch.Tensor:
        """seeds: [rows, B] block; with seeds_ld > 0 it is a column slice of a wider contiguous block whose row
        pitch is seeds_ld (recompute mode passes views, no copy).
        point_range = (lo, hi): the seeds are zero outside these points, sweep only the tiles that hold them
        (tde_jet_backward_range; the fused GradNorm's per-row gradients)."""
        B = x.shape[0]
        if flat_grad is None:
            flat_grad = torch.zeros(self.n_params, dtype=torch.float32, device=x.device)
        if B == 0:
            return flat_grad
        if seeds_ld == 0:
            seeds = seeds.contiguous()
            seeds_ld = B
        elif seeds.stride(1) != 1 or seeds.stride(0) != seeds_ld:
            raise ValueError("seed slice does not have the declared row pitch")
        plan = self.plan(B)
        self._bind(tensors)
        ws = self._ws(plan.workspace_bytes, x.device)
        with _lib.on_device(x.device):
            if point_range is None:
                rc = _lib.lib().tde_jet_backward(C.byref(self._net), C.byref(self._spec), x.data_ptr(), B, stash.data_ptr(),
                                                 seeds.data_ptr(), seeds_ld, grad_scale, flat_grad.data_ptr(), ws.data_ptr(),
                                                 self._stream(x.device))
            else:
                rc = _lib.lib().tde_jet_backward_range(C.byref(self._net), C.byref(self._spec), x.data_ptr(), B, stash.data_ptr(),
                                                       seeds.data_ptr(), seeds_ld, int(point_range[0]), int(point_range[1]),
                                                       grad_scale, flat_grad.data_ptr(), ws.data_ptr(), self._stream(x.device))
        _lib.check(rc, "tde_jet_backward")
        return flat_grad

    def split_flat(self, flat: torch.Tensor) -> List[torch.Tensor]:
        views, o = [], 0
        for sp in self.lins:
            for shape in ((sp.out_features, sp.in_features), (sp.out_features,)):
                n = shape[0] * (shape[1] if len(shape) > 1 else 1)
                views.append(flat[o:o + n].view(shape))
                o += n
        return views


# Largest activation stash kept from forward to backward, as a fraction of the free device memory at call
# time.  Bigger batches switch to recompute mode: the forward keeps nothing and the backward re-runs the
# forward chunk by chunk (+1/3 FLOPs, bounded memory) -- the cylinder config (2e6 pts/GPU x 10x512) needs it.
STASH_FRACTION_OF_FREE = 0.7


class _JetFunction(torch.autograd.Function):
    """out[J*n_out, B] = jets(net, x); differentiable w.r.t. the Linear parameters (not x: no caller needs it)."""

    @staticmethod
    def forward(ctx, x, runner: JetRunner, *params):
        need = any(ctx.needs_input_grad[2:])
        ctx.chunk = 0
        if need and x.shape[0] > 0 and x.is_cuda:
            key = (x.shape[0], STASH_FRACTION_OF_FREE)
            if key not in runner.chunking:      # decided once per batch size so every step allocates the same blocks
                free, _ = torch.cuda.mem_get_info(x.device)
                free += torch.cuda.memory_reserved(x.device) - torch.cuda.memory_allocated(x.device)
                budget = STASH_FRACTION_OF_FREE * free
                # stash + workspace (the tensor-core backward keeps an Hbar stream of the stash's size in the workspace)
                need_bytes = lambda n: runner.plan(n).stash_bytes + runner.plan(n).workspace_bytes      # noqa: E731
                chunk = 0
                if need_bytes(x.shape[0]) > budget:
                    tile = runner.plan(x.shape[0]).tile_points
                    n_chunks = int(-(-need_bytes(x.shape[0]) // budget))
                    while True:      # fixed costs (weights, gradient slabs) do not shrink with the chunk: re-check the chunk's own plan
                        chunk = max(tile, -(-(-(-x.shape[0] // n_chunks)) // tile) * tile)
                        if chunk == tile or need_bytes(chunk) <= budget:
                            break
                        n_chunks += 1
                runner.chunking[key] = chunk
            ctx.chunk = runner.chunking[key]
        ctx.kept = None
        if need and ctx.chunk > 0:
            # recompute mode.  The first pass keeps nothing -- except for the LAST chunk, whose stash goes into the runner's
            # chunk buffer: the backward starts with that chunk and need not run its forward again (1 / n_chunks of the
            # recompute).  Another forward through this runner before the backward simply takes the buffer over (owner tag).
            B = x.shape[0]
            lo_last = ((B - 1) // ctx.chunk) * ctx.chunk
            xc = runner._check_x(x)
            out = torch.empty((runner.ch.n_rows, B), dtype=torch.float32, device=x.device)
            if lo_last > 0:
                runner.forward(xc[:lo_last], params, need_stash=False, out=out, out_col=0)
            ctx.kept = object()
            runner.forward(xc[lo_last:], params, need_stash=True, reuse_stash=True, out=out, out_col=lo_last, owner=ctx.kept)
            stash = None
        else:
            xc, out, stash = runner.forward(x, params, need_stash=need and ctx.chunk == 0)
        ctx.runner = runner
        ctx.stash = stash
        ctx.need = need
        ctx.save_for_backward(xc, *params)
        # what a caller needs to run further reverse sweeps on this forward without autograd (fused GradNorm): picked up by
        # Jet_engine._jets right after apply()
        runner._last_forward = (xc, stash, params) if stash is not None else None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, *params = ctx.saved_tensors
        runner: JetRunner = ctx.runner
        B = x.shape[0]
        if not ctx.need and B > 0:
            raise RuntimeError("backward through a jet forward that kept no stash")
        if ctx.chunk == 0:
            flat = runner.backward(x, params, ctx.stash, grad_out)
        else:   # recompute mode
            flat = torch.zeros(runner.n_params, dtype=torch.float32, device=x.device)
            grad_out = grad_out.contiguous()
            starts = list(range(0, B, ctx.chunk))
            if ctx.kept is not None and runner._chunk_stash_owner is ctx.kept:
                lo = starts.pop()          # the forward left this chunk's stash in the runner's buffer
                runner.backward(x[lo:B], params, runner._chunk_stash, grad_out[:, lo:B], flat_grad=flat, seeds_ld=B)
            for lo in starts:
                hi = min(B, lo + ctx.chunk)
                xc, _, stash = runner.forward(x[lo:hi], params, need_stash=True, reuse_stash=True)
                runner.backward(xc, params, stash, grad_out[:, lo:hi], flat_grad=flat, seeds_ld=B)
                del stash
        return (None, None, *runner.split_flat(flat))


class JetForwardRecord:
    """Handle on one stash-keeping forward: lets `grad_balance.GradNorm` run range-restricted reverse sweeps on it."""
    __slots__ = ('runner', 'net', 'x', 'stash', 'params', 'versions')

    def __init__(self, runner, net, x, stash, params):
        self.runner, self.net, self.x, self.stash, self.params = runner, net, x, stash, params
        self.versions = [p._version for p in params]     # like autograd's saved-tensor check: the stash belongs to THESE weights

    def gradient(self, seeds: torch.Tensor, point_range=None, flat_grad: Optional[torch.Tensor] = None) -> torch.Tensor:
        """flat parameter gradient (net.parameters() order of the Linear layers) of sum(seeds * block)."""
        if any(p._version != v for p, v in zip(self.params, self.versions)):
            raise RuntimeError("the network parameters were modified in place after this forward (optimizer.step()?): "
                               "its activation stash no longer matches them")
        if point_range is not None:
            lo, hi = point_range
            if hi <= lo:
                return flat_grad if flat_grad is not None else torch.zeros(self.runner.n_params, dtype=torch.float32, device=self.x.device)
        return self.runner.backward(self.x, self.params, self.stash, seeds, flat_grad=flat_grad, point_range=point_range)


class JetOutput(dict):
    """The engine's result dict ({'all': {...}, group: {...}}) plus the raw stream block it views, so the loss
    side can run the fused residual kernel on it: `block[J*n_out, B]`, `stream_row[name]`, `ranges[group]`."""
    block: torch.Tensor = None
    stream_row: Dict[str, int] = None
    ranges: Dict[str, tuple] = None


def _group_name(key, group):
    return getattr(group, 'name', key)


def _group_points(group) -> torch.Tensor:
    return group.inputs['input'] if hasattr(group, 'inputs') else group


class Jet_engine(_EngineBase):
    """Taylor-mode replacement for AD_engine (same constructor shape: net, derivatives)."""

    def __init__(self, net: torch.nn.Module, derivatives: Dict[str, tuple], **kwargs):
        super().__init__()
        self.net = net
        self.derivatives = derivatives
        self.output_vars = {k: v[0] for k, v in derivatives.items() if len(v) == 1}   # Engines/base.py:56-58
        lins, act = extract_layers(net)
        d_in, n_out = lins[0].in_features, lins[-1].out_features
        self._full = JetRunner(lins, act, JetChannels({k: tuple(v) for k, v in derivatives.items()}, d_in, n_out))
        self._value_only = None     # built on first use (groups excluded by target_groups)
        self._lins, self._act = lins, act

    # -- helpers -----------------------------------------------------------------------------
    def _jets(self, runner: JetRunner, x: torch.Tensor) -> torch.Tensor:
        runner._last_forward = None
        block = _JetFunction.apply(x, runner, *runner.effective_tensors())
        if runner._last_forward is not None:
            block._tde_forward = JetForwardRecord(runner, self.net, *runner._last_forward)
            runner._last_forward = None
        return block

    def _named(self, runner: JetRunner, block: torch.Tensor, names, lo=None, hi=None) -> Dict[str, torch.Tensor]:
        if lo is None:
            return {n: block[runner.ch.row(n)] for n in names}
        return {n: block[runner.ch.row(n), lo:hi] for n in names}

    @staticmethod
    def _cat_groups(groups):
        names = [_group_name(k, g) for k, g in groups]
        pts = [_group_points(g) for _, g in groups]
        sizes = [p.shape[0] for p in pts]
        return torch.cat(pts), names, sizes

    def _differentiate(self, groups) -> Dict[str, Dict[str, torch.Tensor]]:
        x, names, sizes = self._cat_groups(groups)
        block = self._jets(self._full, x)
        keys = list(self.derivatives.keys())
        out = JetOutput({'all': self._named(self._full, block, keys)})
        out.block = block
        out.stream_row = {k: self._full.ch.row(k) for k in keys}
        out.ranges = {}
        lo = 0
        for name, n in zip(names, sizes):
            out[name] = self._named(self._full, block, keys, lo, lo + n)
            out.ranges[name] = (lo, lo + n)
            lo += n
        return out

    def _values_only(self, groups) -> Dict[str, Dict[str, torch.Tensor]]:
        if not groups:
            return {}
        if self._value_only is None:
            d_in, n_out = self._lins[0].in_features, self._lins[-1].out_features
            self._value_only = JetRunner(self._lins, self._act,
                                         JetChannels({k: (i,) for k, i in self.output_vars.items()}, d_in, n_out))
        x, names, sizes = self._cat_groups(groups)
        block = self._jets(self._value_only, x)
        out, lo = {}, 0
        for name, n in zip(names, sizes):
            out[name] = self._named(self._value_only, block, self.output_vars.keys(), lo, lo + n)
            lo += n
        return out

    # -- the engine contract -----------------------------------------------------------------
    def calculate(self, x: Union[torch.Tensor, dict], target_groups: Union[str, Sequence[str], None] = None,
                  **kwargs) -> Dict[str, Dict[str, torch.Tensor]]:
        if isinstance(x, torch.Tensor):
            block = self._jets(self._full, x)
            return {'all': self._named(self._full, block, list(self.derivatives.keys()))}
        if isinstance(x, dict):
            items = list(x.items())
            if target_groups is None:
                return self._differentiate(items)
            targets = [target_groups] if isinstance(target_groups, str) else list(target_groups)
            for t in targets:
                if t not in x:
                    raise KeyError(t)
            out = self._values_only([(k, g) for k, g in items if _group_name(k, g) not in targets])
            out.update(self._differentiate([(t, x[t]) for t in targets]))
            return out
        raise ValueError(f'input should be of type dict or torch.tensor. Got instead {type(x)}')


class FD_engine(_EngineBase):
    """Finite-difference engine with the reference's contract (src/torch_DE/continuous/Engines/FD.py:4-133): central
    first / second differences (non-mixed, order <= 2) from value-only network passes, step clipped by a signed-distance
    function.  All 2*d+1 stencil evaluations of every point go through ONE value-only (J=1) jet-kernel launch, and the
    result stays differentiable w.r.t. the network parameters."""

    def __init__(self, net: torch.nn.Module, derivatives: Dict[str, tuple], dxs, sdf=None, **kwargs) -> None:
        super().__init__()
        self.net = net
        self.derivatives = derivatives
        self.output_vars = {k: v[0] for k, v in derivatives.items() if len(v) == 1}
        self.dims = len(dxs)
        self.dxs = [float(d) for d in dxs]
        self.sdf = sdf
        for name, idx in derivatives.items():
            if len(idx) > 3 or (len(idx) == 3 and idx[1] != idx[2]):
                raise ValueError('Only upto second order non mixed derivatives are currently supported')
        lins, act = extract_layers(net)
        d_in, n_out = lins[0].in_features, lins[-1].out_features
        assert d_in == self.dims, f'Engine is for a PINN of dimension {self.dims}. Got a network with {d_in} inputs instead'
        self._values = JetRunner(lins, act, JetChannels({k: (i,) for k, i in self.output_vars.items()}, d_in, n_out))

    def finite_diff(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        B = x.shape[0]
        steps = []
        for dx in self.dxs:
            h = torch.full((B,), dx, device=x.device, dtype=x.dtype)
            if self.sdf is not None:
                h = torch.minimum(self.sdf(x).to(x.dtype), h)
            steps.append(h)
        pts = [x]
        for i, h in enumerate(steps):
            lo, hi = x.clone(), x.clone()
            lo[:, i] -= h
            hi[:, i] += h
            pts += [lo, hi]
        block = _JetFunction.apply(torch.cat(pts), self._values, *self._values.effective_tensors())   # [n_out, (2d+1)B]
        out = {}
        for name, idx in self.derivatives.items():
            row = block[self._values.ch.row(next(k for k, v in self.output_vars.items() if v == idx[0]))]
            u2 = row[:B]
            if len(idx) == 1:
                out[name] = u2
                continue
            j = idx[1]
            u1, u3 = row[(1 + 2 * j) * B:(2 + 2 * j) * B], row[(2 + 2 * j) * B:(3 + 2 * j) * B]
            h = steps[j]
            out[name] = (u3 - u1) / (2 * h) if len(idx) == 2 else (u1 - 2 * u2 + u3) / h.pow(2)
        return out

    def calculate(self, x, target_groups=None, **kwargs) -> Dict[str, Dict[str, torch.Tensor]]:
        if target_groups is not None:
            target_groups = [target_groups] if isinstance(target_groups, str) else list(target_groups)
        if isinstance(x, dict):
            items = list(x.items())
            chosen = items if target_groups is None else [(t, x[t]) for t in target_groups]
            out = {}
            others = [] if target_groups is None else [(k, g) for k, g in items if _group_name(k, g) not in target_groups]
            if others:
                xo, names, sizes = Jet_engine._cat_groups(others)
                blk = _JetFunction.apply(xo, self._values, *self._values.effective_tensors())
                lo = 0
                for name, n in zip(names, sizes):
                    out[name] = {k: blk[self._values.ch.row(k), lo:lo + n] for k in self.output_vars}
                    lo += n
            xd, names, sizes = Jet_engine._cat_groups(chosen)
            derivs = self.finite_diff(xd)
            lo = 0
            for name, n in zip(names, sizes):
                out[name] = {k: v[lo:lo + n] for k, v in derivs.items()}
                lo += n
            return out
        if isinstance(x, torch.Tensor):
            return {target_groups[0] if target_groups else 'all': self.finite_diff(x)}
        raise ValueError(f'input should be of type dict or torch.tensor. Got instead {type(x)}')
