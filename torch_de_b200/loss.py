"""Loss table with the reference's `Loss_handler` / `Loss` surface (src/torch_DE/utils/loss.py:18-313),
without pandas: rows are plain records, so a training step does no DataFrame masking (`.loc/.apply`,
loss.py:192-199, is the reference's dominant host cost at small batch sizes).

Semantics kept: row loss = aggregate(weight * |residual|**power) with aggregate = mean|sum (loss.py:25-59);
`sum()` adds the rows (:60-64); `grouped_losses(by)` sums rows per 'loss_type' | 'group' | 'variable' (:72-83).
"""
from typing import Callable, Dict, List, Union

import torch
from torch import Tensor

from .equations import DE_func
from . import fused_loss as FL


class GroupedLosses(dict):
    """dict of name -> scalar Tensor; `.index` for code written against the pandas Series the reference returns."""
    @property
    def index(self):
        return list(self.keys())


class RowStack(Tensor):
    """The stacked row losses.  Iterating it (`*loss.individual_losses()`) yields the original row tensors, which remember the
    stream block and point range they were computed from (grad_balance.GradNorm's fused path); everything else is a Tensor."""
    def __iter__(self):
        rows = getattr(self, '_rows', None)
        return iter(rows) if rows is not None else super().__iter__()


class Loss:
    def __init__(self, rows: List[dict], aggregation_method: str = 'mean', power: Union[int, None] = 2):
        self.losses = rows
        self.aggregation_method = aggregation_method
        self.power = power
        self._agg = torch.mean if aggregation_method == 'mean' else torch.sum
        self._done = False

    def error_func(self, r: Tensor) -> Tensor:
        return torch.abs(r).pow(self.power) if self.power is not None else r

    def _materialise_points(self, row):
        """Rows reduced by the fused kernel keep no per-point tensors; the first request launches the kernel's point pass for
        the row's whole group (tde_residual_loss_points) and fills residual + weighting of every row of that group.  These
        are detached values for samplers and diagnostics (R3_sampler, Causal_weighting), as in sampling.py:89 `no_grad`."""
        if row.get('residual') is None and row.get('_points') is not None:
            run, i = row['_points']          # (point pass of the row's group -- shared by its rows --, row index in the group)
            res, wts = run()
            row['residual'], row['weighting'] = res[i], wts[i]

    def point_error(self) -> List[Tensor]:
        for row in self.losses:
            self._materialise_points(row)
            if row.get('residual') is None:
                raise RuntimeError("per-point residuals of this row are not available")
            if row.get('point_error') is None:
                row['point_error'] = self.error_func(row['residual'])
        return [row['point_error'] for row in self.losses]

    def weighted_point_error(self) -> List[Tensor]:
        self.point_error()
        for row in self.losses:
            if row.get('weighted_error') is None:
                row['weighted_error'] = row['weighting'] * row['point_error']
        return [row['weighted_error'] for row in self.losses]

    def residual_rows(self, group: str, weighted: bool = True) -> List[Tensor]:
        """Per-point (weighted) errors of the 'residual' rows of one group, in table order: what the reference's R3_sampler
        looks up as `table['Residual'][group]` (utils/sampling.py:72-76)."""
        errs = self.weighted_point_error() if weighted else self.point_error()
        return [e for e, row in zip(errs, self.losses) if row['group'] == group and str(row['loss_type']).lower() == 'residual']

    def aggregated_loss(self) -> List[Tensor]:
        if not self._done:
            for row in self.losses:
                if row.get('aggregated_loss') is None:      # rows reduced by the fused kernel arrive aggregated
                    if row.get('point_error') is None:
                        row['point_error'] = self.error_func(row['residual'])
                    row['weighted_error'] = row['weighting'] * row['point_error']
                    row['aggregated_loss'] = self._agg(row['weighted_error'])
                if row.get('_src') is not None:
                    row['aggregated_loss']._tde_row = row['_src']      # (stream block, point range of the row's group)
            self._done = True
        return [row['aggregated_loss'] for row in self.losses]

    def sum(self) -> Tensor:
        return sum(self.aggregated_loss())

    def individual_losses(self) -> Tensor:
        rows = list(self.aggregated_loss())
        stacked = torch.stack(rows).as_subclass(RowStack)
        stacked._rows = rows           # `*loss.individual_losses()` (Stokes_Cylinder.py:108) hands GradNorm the row tensors themselves
        return stacked

    def grouped_losses(self, groupby: str) -> GroupedLosses:
        if groupby not in ('loss_type', 'group', 'variable'):
            raise ValueError(f'groupby must be one of the following strings: "loss_type", "group" or "variable". Got {groupby} instead')
        self.aggregated_loss()
        out = GroupedLosses()
        for row in self.losses:
            k = row[groupby]
            out[k] = out[k] + row['aggregated_loss'] if k in out else row['aggregated_loss']
        return out

    def backward(self):
        self.sum().backward()

    def __len__(self):
        return len(self.losses)

    def print_losses(self, epoch, groupby='loss_type'):
        total = self.sum()
        if groupby is None:
            names = [f"{r['loss_type']}__{r['group']}__{r['variable']}" for r in self.losses]
            vals = self.aggregated_loss()
        else:
            g = self.grouped_losses(groupby)
            names, vals = list(g.keys()), list(g.values())
        body = '\t'.join(f'{n}: {float(v):.3E}' for n, v in zip(names, vals))
        print(f'Epoch {epoch}:- Total Loss {float(total):.3E}\t {body}')


class Loss_handler:
    def __init__(self, dataset=None, fused: bool = True, revalidate: bool = True) -> None:
        self.rows: List[dict] = []
        self.fused = fused                 # use tde_residual_loss when the engine output allows it
        # Re-trace the rows symbolically at every call (pure Python, no tensor work) and recompile a group whose trace
        # changed: a Python number captured by a residual closure (an annealed viscosity ...) is a constant of the compiled
        # program.  Tensor scalars never need this (they are re-read on the device); `revalidate=False` skips the check.
        self.revalidate = revalidate
        self.fusion_report: Dict[str, str] = {}   # group -> 'fused' | reason it stayed on the torch path
        self._programs: Dict[tuple, object] = {}
        self.update_dataset(dataset)

    def update_dataset(self, dataset):
        self.dataset = dataset
        if dataset is None:
            self._group_names = None
        elif hasattr(dataset, 'group_names'):
            self._group_names = list(dataset.group_names())
        elif isinstance(dataset, dict):
            self._group_names = list(dataset.keys())
        else:
            self._group_names = list(dataset)

    def __len__(self):
        return len(self.rows)

    def __call__(self, group_input: Dict, group_output: Dict, **kwargs) -> Loss:
        return self.calculate(group_input, group_output, **kwargs)

    def calculate(self, batched_input, batched_output, power: int = 2, aggregation_method='mean', fused=None) -> Loss:
        use_fused = self.fused if fused is None else fused
        fused_groups = {}
        if (use_fused and power == 2 and aggregation_method == 'mean'
                and getattr(batched_output, 'block', None) is not None and batched_output.block.is_cuda):
            fused_groups = self._fused_rows(batched_input, batched_output)
        block = getattr(batched_output, 'block', None)
        ranges = getattr(batched_output, 'ranges', None) or {}
        n_all = block.shape[1] if block is not None else 0
        rows = []
        group_rows: Dict[str, list] = {}      # group -> [point pass, rows handed out so far]
        for spec in self.rows:
            src = None if block is None else (block, (0, n_all) if spec['custom'] else ranges.get(spec['group']))
            if src is not None and src[1] is None:
                src = None
            if not spec['custom'] and spec['group'] in fused_groups:
                row = dict(spec)
                row['_src'] = src
                row['residual'] = None
                row['weighting'] = None
                row['aggregated_loss'] = fused_groups[spec['group']].pop(0)
                g = spec['group']
                if g not in group_rows:
                    group_rows[g] = [self._point_pass(block, self._launched[g]), 0]
                row['_points'] = (group_rows[g][0], group_rows[g][1])
                group_rows[g][1] += 1
                rows.append(row)
                continue
            row = dict(spec)
            row['_src'] = src
            if spec['custom']:
                row['residual'] = spec['evaluation'](batched_input, batched_output)
                row['weighting'] = 1.
            else:
                g = spec['group']
                if g not in batched_input:
                    continue
                gi, go = batched_input[g], batched_output[g]
                row['residual'] = spec['evaluation'](gi, go)
                row['weighting'] = spec['weighting_function'](gi, go)
            rows.append(row)
        return Loss(rows, aggregation_method, power)

    @staticmethod
    def _point_pass(block, launch):
        """Closure that runs the loss kernel's point pass for one fused group once and caches the result.  It must not refer
        back to the rows that hold it: a reference cycle would keep the stream block -- and through it the activation stash of
        the step, tens of GB -- alive until Python's cycle collector happens to run."""
        cache = []

        def run():
            if not cache:
                gp, aux, lo, n = launch
                with torch.no_grad():
                    cache.append(FL.point_rows(block.detach(), gp, aux, lo, n))
            return cache[0]
        return run

    # -- fused path ---------------------------------------------------------------------------
    def _fused_rows(self, batched_input, out) -> Dict[str, list]:
        """Launch the fused kernel for every group whose rows trace into a register program.
        Returns {group: [row loss tensors in table order]}; other groups stay on the torch path."""
        names = list(out.stream_row.keys())
        n_stream_rows = out.block.shape[0]
        launches, owners = [], []
        self._launched = {}
        for g, (lo, hi) in out.ranges.items():
            specs = [r for r in self.rows if r['group'] == g and not r['custom']]
            if not specs or g not in batched_input:
                continue
            group = batched_input[g]
            key = (g, tuple(names), len(specs))
            gp = self._programs.get(key)
            try:
                # rows whose weighting function was traced into the program (SDF lookup and the like) need no torch evaluation
                weights = [None if gp and gp.traced_weight[i] else FL.eval_weight(sp, group, names) for i, sp in enumerate(specs)]
                if gp and self.revalidate and FL.signature_of(group, specs, out.stream_row) != gp.signature:
                    gp = None      # a Python constant behind a residual closure changed: compile the rows again
                    weights = [FL.eval_weight(sp, group, names) for sp in specs]
                if gp is None:
                    gp = FL.trace_group(group, specs, out.stream_row, n_stream_rows)
                    for i, w in enumerate(weights):
                        if isinstance(w, torch.Tensor) and not gp.traced_weight[i]:
                            gp.weight_slot(i)
                    self._programs[key] = gp
                    self.fusion_report[g] = 'fused'
                elif gp is False:
                    continue
            except FL.TraceError as exc:
                self._programs[key] = False
                self.fusion_report[g] = f'torch path: {exc}'
                continue
            n = hi - lo
            aux = None
            if gp.aux_names:
                cols = []
                for a in gp.aux_names:
                    if a.startswith('__w'):
                        w = weights[int(a[3:])]
                        cols.append(w.to(out.block.device) if isinstance(w, torch.Tensor) else torch.full((n,), float(w), device=out.block.device))
                    elif a.startswith('__s'):       # tensor scalar of a residual: re-read every launch, no host sync
                        cols.append(gp.scalars[int(a[3:])].detach().to(device=out.block.device, dtype=torch.float32).reshape(1).expand(n))
                    else:
                        cols.append(group.batchables[a].to(torch.float32))
                aux = torch.stack(cols).contiguous()
            for i, w in enumerate(weights):
                if gp.weight_aux[i] is None and not gp.traced_weight[i]:
                    if isinstance(w, torch.Tensor):
                        raise RuntimeError(f"weighting of group '{g}' changed from a constant to a per-point tensor")
                    gp.prog.w_const[i] = float(w)
            if gp.grids:
                gp.bind_grids(out.block.device)
            launches.append((gp, aux, lo, n))
            self._launched[g] = launches[-1]
            owners.append((g, gp.prog.n_rows))
        if not launches:
            return {}
        rows = FL.fused_rows(out.block, launches)
        result, off = {}, 0
        for g, k in owners:
            result[g] = [rows[off + i] for i in range(k)]
            off += k
        return result

    def set_terms(self, loss_type, group, var_dict: Dict[str, Callable], weighting, custom: bool = False):
        self._programs.clear()
        if self._group_names is not None and not custom:
            assert group in self._group_names, f"group '{group}' is not in the dataset"
        if not isinstance(weighting, dict):
            weighting = {k: weighting for k in var_dict}
        for (var_name, fn), w in zip(var_dict.items(), weighting.values()):
            if any(r['group'] == group and r['loss_type'] == loss_type and r['variable'] == var_name for r in self.rows):
                raise ValueError(f'variable name {var_name} already exists in group {group} as a {loss_type} term')
            wf = w if callable(w) else (lambda gi, go, _w=w: _w)
            self.rows.append({'loss_type': loss_type, 'group': group, 'variable': var_name, 'evaluation': fn,
                              'weighting_function': wf, 'weighting_spec': w, 'custom': custom})

    @staticmethod
    def create_residual_from_rhs(var_name, rhs):
        rhs_func = rhs if callable(rhs) else DE_func(lambda _rhs=rhs, **kwargs: _rhs)
        fn = lambda group_input, group_output: group_output[var_name] - rhs_func(group_input, group_output)  # noqa: E731
        fn.rhs_spec = (var_name, rhs)
        return fn

    def add_boundary(self, group, bound_dict: Dict[str, Union[float, Callable]], weighting=1):
        self.set_terms('boundary', group, {v: self.create_residual_from_rhs(v, rhs) for v, rhs in bound_dict.items()}, weighting)

    def add_initial_condition(self, group: str, ic_dict: dict, weighting=1):
        self.set_terms('initial condition', group, {v: self.create_residual_from_rhs(v, rhs) for v, rhs in ic_dict.items()}, weighting)

    def add_residual(self, group, residuals: Dict[str, Callable], weighting=1):
        self.set_terms('residual', group, residuals, weighting)

    def add_data_constraint(self, loss_type, group, data_vars: List[str], weighting=1.):
        def make(var_name):
            return lambda gi, go: go[var_name] - gi.batchables[var_name]
        self.set_terms(loss_type, group, {v: make(v) for v in data_vars}, weighting)

    def add_periodic(self, group_1: str, group_2: str, variable: str):
        def periodic(batched_input, batched_output):
            return batched_output[group_1][variable] - batched_output[group_2][variable]
        self.set_terms('periodic', group_1, {variable: periodic}, weighting=1., custom=True)

    def add_custom_function(self, loss_type: str, group: str, func_dict: Dict[str, Callable]):
        self.set_terms(loss_type, group, func_dict, weighting=1., custom=True)
