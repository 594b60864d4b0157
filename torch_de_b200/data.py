"""Input containers the engine accepts (the loader itself is host glue and out of scope).

`PINN_dict` / `PINN_group` keep the attribute surface of the reference (src/torch_DE/utils/data.py:12-112)
that the hot path touches: `group.name`, `group.inputs['input']` ([N, D]), per-variable columns,
`group.batchables` / `group.unbatchables` (what a DE_func receives as keyword arguments), `.to()`.
No tensordict dependency: plain dicts of tensors.
"""
from typing import Dict, Optional, Sequence

import torch


class PINN_dict(dict):
    def to(self, *args, **kwargs):
        for k in list(self.keys()):
            self[k] = self[k].to(*args, **kwargs)
        return self


class PINN_group:
    def __init__(self, name: str, inputs: torch.Tensor, input_vars: Sequence[str],
                 batchable_kwargs: Optional[Dict[str, torch.Tensor]] = None,
                 unbatched_kwargs: Optional[dict] = None, batch_size: Optional[int] = None):
        if inputs.dim() != 2 or inputs.shape[1] != len(input_vars):
            raise ValueError(f"group '{name}': inputs must be [N, {len(input_vars)}], got {tuple(inputs.shape)}")
        self.name = name
        self.input_vars = list(input_vars)
        self.N, self.D = inputs.shape
        self.batch_size = self.N if batch_size is None else batch_size
        self.batchable_kwargs = dict(batchable_kwargs or {})
        for k, v in self.batchable_kwargs.items():
            if k == 'input' or k in self.input_vars:
                raise ValueError(f"Shared keys detected: {{{k!r}}}")
            if len(v) != self.N:
                raise ValueError('All values for batchable_kwargs must have the same length in the batch dimension')
        self.unbatchables = dict(unbatched_kwargs or {})
        self._set_inputs(inputs)

    def _set_inputs(self, inputs):
        self.inputs = {'input': inputs}
        self.inputs.update({v: inputs[:, i] for i, v in enumerate(self.input_vars)})
        self.batchables = {**self.inputs, **self.batchable_kwargs}

    def __len__(self):
        return int(self.N)

    def to(self, *args, **kwargs):
        self.batchable_kwargs = {k: v.to(*args, **kwargs) for k, v in self.batchable_kwargs.items()}
        self.unbatchables = {k: (v.to(*args, **kwargs) if hasattr(v, 'to') else v) for k, v in self.unbatchables.items()}
        self._set_inputs(self.inputs['input'].to(*args, **kwargs))
        return self

    def subgroup(self, idx):
        return PINN_group(self.name, self.inputs['input'][idx], self.input_vars,
                          {k: v[idx] for k, v in self.batchable_kwargs.items()}, self.unbatchables, self.batch_size)


def make_batch(input_vars: Sequence[str], groups: Dict[str, torch.Tensor], **kwargs) -> PINN_dict:
    """{group name: points[N,D]} -> PINN_dict of PINN_group (what PINN_dataset.__getitem__ yields, data.py:168-169)."""
    return PINN_dict({name: PINN_group(name, pts, input_vars, **kwargs) for name, pts in groups.items()})
