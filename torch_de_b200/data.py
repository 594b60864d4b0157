"""Input containers the engine accepts, and a device-resident batch iterator.

`PINN_dict` / `PINN_group` keep the attribute surface of the reference (src/torch_DE/utils/data.py:12-112)
that the hot path touches: `group.name`, `group.inputs['input']` ([N, D]), per-variable columns,
`group.batchables` / `group.unbatchables` (what a DE_func receives as keyword arguments), `.to()`.
No tensordict dependency: plain dicts of tensors.

`PINN_sampler` keeps the reference's batch-count rule (data.py:203-255); `DeviceLoader` applies it to groups that already
live on the GPU, so a training step starts without a host->device copy.
"""
import warnings
from typing import Dict, Iterator, Optional, Sequence

import torch


class PINN_dict(dict):
    def to(self, *args, **kwargs):
        for k in list(self.keys()):
            self[k] = self[k].to(*args, **kwargs)
        return self


class PINN_group:
    def __init__(self, name: str, inputs: torch.Tensor, input_vars: Sequence[str],
                 batchable_kwargs: Optional[Dict[str, torch.Tensor]] = None,
                 unbatched_kwargs: Optional[dict] = None, batch_size: Optional[int] = None, shuffle: bool = False):
        if inputs.dim() != 2 or inputs.shape[1] != len(input_vars):
            raise ValueError(f"group '{name}': inputs must be [N, {len(input_vars)}], got {tuple(inputs.shape)}")
        self.name = name
        self.input_vars = list(input_vars)
        self.N, self.D = inputs.shape
        self.batch_size = self.N if batch_size is None else batch_size
        self.shuffle = shuffle
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
                          {k: v[idx] for k, v in self.batchable_kwargs.items()}, self.unbatchables, self.batch_size, self.shuffle)


def make_batch(input_vars: Sequence[str], groups: Dict[str, torch.Tensor], **kwargs) -> PINN_dict:
    """{group name: points[N,D]} -> PINN_dict of PINN_group (what PINN_dataset.__getitem__ yields, data.py:168-169)."""
    return PINN_dict({name: PINN_group(name, pts, input_vars, **kwargs) for name, pts in groups.items()})


class PINN_sampler:
    """Batch indices per group with the reference's rule (utils/data.py:203-255): every group yields as many batches as the
    group with the most batches (len // batch_size); a group with fewer repeats its index list (a fresh permutation when
    `group.shuffle`) and tops it up from the list's start.  Index tensors are created on the device the group lives on."""

    def __init__(self, groups: Dict[str, PINN_group]):
        self.groups = groups
        self._warned = False

    @staticmethod
    def make_indices(groups: Dict[str, PINN_group]) -> Dict[str, torch.Tensor]:
        n_batches = {g.name: len(g) // g.batch_size for g in groups.values()}
        most = max(n_batches.values())
        out = {}
        for g in groups.values():
            dev = g.inputs['input'].device
            base = torch.randperm(len(g), device=dev) if g.shuffle else torch.arange(len(g), device=dev)
            idx = base.repeat(most // n_batches[g.name])
            out[g.name] = torch.cat([idx, idx[:(most % n_batches[g.name]) * g.batch_size]])
        return out

    def __len__(self) -> int:
        counts = [(len(g) // g.batch_size, len(g) % g.batch_size) for g in self.groups.values()]
        most, rest = max(counts, key=lambda c: c[0])
        if rest and not self._warned:
            warnings.warn("The group with the most number of batches is not divisible by it's batch size. "
                          "The last batch will be ignored. This warning will no longer appear")
        self._warned = True
        return most

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        self.indices = self.make_indices(self.groups)
        for i in range(len(self)):
            yield {g.name: self.indices[g.name][i * g.batch_size:(i + 1) * g.batch_size] for g in self.groups.values()}


class DeviceLoader:
    """for batch in DeviceLoader(groups): ...  -- `batch` is a PINN_dict of sub-groups (what the reference's DataLoader yields,
    data.py:168-169, 258-275), cut from device-resident groups by `PINN_sampler`: no per-step host work or copy."""

    def __init__(self, groups: Dict[str, PINN_group]):
        self.groups = groups
        self.sampler = PINN_sampler(groups)

    def __len__(self) -> int:
        return len(self.sampler)

    def __iter__(self) -> Iterator[PINN_dict]:
        for idx in self.sampler:
            yield PINN_dict({name: self.groups[name].subgroup(i) for name, i in idx.items()})
