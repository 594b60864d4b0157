"""Derivative spec handling: 'u_xx' -> index tuple -> jet channel.

`derivative_index_map` mirrors DE_Getter.set_vars/set_derivatives/get_deriv_index of the reference
(src/torch_DE/continuous/DE_module.py:53-122): outputs are registered first as (i,), every requested
derivative registers itself and its lower-order prefixes (Deriv.previous_Deriv, symbols/DE_Symbols.py:109-114).

`JetChannels` turns the set of index tuples into the Taylor-mode channel layout the kernels use:
channel 0 = value, then one channel per differentiated input (first order), then one per unordered
pair (second order).  Only the (n_first, pair-mask) combinations compiled into libtde_jet.so exist, so
the requested set is permuted / widened to the smallest compiled superset.
"""
import itertools
from typing import Dict, Sequence, Tuple

# (n_first, pair mask) instantiated in csrc/jet_launch.h (TDE_SPEC_LIST)
COMPILED_SPECS = {0: [0], 1: [0, 1], 2: [0, 1, 5, 7], 3: [0, 9, 63]}


def derivative_index_map(input_vars: Sequence[str], output_vars: Sequence[str],
                         derivatives: Sequence[str], table: Dict[str, Tuple[int, ...]] = None) -> Dict[str, Tuple[int, ...]]:
    in_pos = {str(v): i for i, v in enumerate(input_vars)}
    out_pos = {str(v): i for i, v in enumerate(output_vars)}
    table = {} if table is None else table
    for o, i in out_pos.items():
        table.setdefault(o, (i,))
    for name in derivatives:
        name = str(name)
        parts = name.split('_')
        if len(parts) != 2:
            raise ValueError(f"derivative '{name}' must look like '<output>_<inputs>', e.g. 'u_xx'")
        out, ins = parts
        assert out in out_pos, f'Output Variable {out} does not exist'
        for v in ins:
            assert v in in_pos, f'Variable {v} is not an input Variable'
        chain = list(ins)
        while chain:
            key = out + '_' + ''.join(chain)
            if key not in table:
                table[key] = (out_pos[out],) + tuple(in_pos[v] for v in chain)
            chain.pop()
    return table


def _pair_rank(nf, a, b):
    return sum(nf - i for i in range(a)) + (b - a)


class JetChannels:
    """Channel layout for a set of index tuples {name: (out, in...)}."""

    def __init__(self, index_map: Dict[str, Tuple[int, ...]], d_in: int, n_out: int):
        order = max((len(t) for t in index_map.values()), default=1) - 1
        if order > 2:
            raise NotImplementedError("jet kernels carry derivatives up to second order (north star); "
                                      f"order {order} was requested")
        firsts, pairs = set(), set()
        for t in index_map.values():
            if len(t) == 2:
                firsts.add(t[1])
            elif len(t) == 3:
                firsts.update(t[1:])
                pairs.add(tuple(sorted(t[1:])))
        if len(firsts) > 3:
            raise NotImplementedError("at most 3 differentiated inputs are supported")
        nf = len(firsts)
        best = None
        for perm in itertools.permutations(sorted(firsts)):
            pos = {v: i for i, v in enumerate(perm)}
            need = 0
            for i, j in pairs:
                a, b = sorted((pos[i], pos[j]))
                need |= 1 << _pair_rank(nf, a, b)
            for mask in COMPILED_SPECS[nf]:
                if mask & need == need:
                    cost = bin(mask).count('1')
                    if best is None or cost < best[0]:
                        best = (cost, perm, mask)
        if best is None:
            raise NotImplementedError(f"no compiled jet spec covers first={sorted(firsts)} pairs={sorted(pairs)}")
        _, perm, mask = best
        self.d_in, self.n_out = d_in, n_out
        self.first_in = list(perm)
        self.n_first = nf
        self.pairs = [(a, b) for a in range(nf) for b in range(a, nf) if mask >> _pair_rank(nf, a, b) & 1]
        self.mask = mask
        self.n_channels = 1 + nf + len(self.pairs)
        pos = {v: i for i, v in enumerate(perm)}
        self._row = {}
        for name, t in index_map.items():
            if len(t) == 1:
                c = 0
            elif len(t) == 2:
                c = 1 + pos[t[1]]
            else:
                a, b = sorted((pos[t[1]], pos[t[2]]))
                c = 1 + nf + self.pairs.index((a, b))
            self._row[name] = c * n_out + t[0]

    def row(self, name: str) -> int:
        """Row of the [J*n_out, B] stream block that holds `name`."""
        return self._row[name]

    @property
    def n_rows(self):
        return self.n_channels * self.n_out
