"""`DE_Getter`: derivative-spec front end with the reference's surface (src/torch_DE/continuous/DE_module.py:25-199).

Same constructor, `set_vars`, `set_derivatives`, `set_deriv_method`, `calculate`/`__call__`; the default
(and only built-in) method is the jet engine.  Any object with a `calculate` method can be installed
as the engine, as in the reference (:148-149).
"""
from typing import Sequence

import torch

from .engine import FD_engine, Jet_engine, _EngineBase
from .networks import extract_layers
from .spec import derivative_index_map


class DE_Getter:
    def __init__(self, net: torch.nn.Module, input_vars: Sequence[str] = None, output_vars: Sequence[str] = None,
                 derivatives: Sequence[str] = None, *args, **kwargs) -> None:
        self.net = net
        self.derivatives = {}
        self.deriv_method = None
        if input_vars is not None and output_vars is not None:
            self.set_vars(input_vars, output_vars)
        if derivatives is not None:
            self.set_derivatives(derivatives)

    def set_vars(self, input_vars, output_vars, net_check: bool = True):
        self.input_vars = [str(v) for v in input_vars]
        self.output_vars = [str(v) for v in output_vars]
        self.input_vars_idx = {v: i for i, v in enumerate(self.input_vars)}
        self.output_vars_idx = {v: i for i, v in enumerate(self.output_vars)}
        self.derivatives.update({o: (i,) for i, o in enumerate(self.output_vars)})
        if net_check:
            # the reference round-trips the net through the CPU to probe shapes (:68-81); read them instead
            try:
                lins, _ = extract_layers(self.net)
            except NotImplementedError:
                lins = []
            if lins:
                assert lins[-1].out_features == len(self.output_vars), \
                    (f'The output of the network of size {lins[-1].out_features} does not match the number of '
                     f'output variables given {len(self.output_vars)}')
                if lins[0].in_features != len(self.input_vars):
                    print(f'The number of input vars provided {len(self.input_vars)} does not match the input size of the network')

    def set_derivatives(self, derivatives) -> None:
        derivative_index_map(self.input_vars, self.output_vars, [str(d) for d in derivatives], self.derivatives)
        self.set_deriv_method('jet')

    def set_deriv_method(self, deriv_method, **kwargs):
        if not isinstance(self.derivatives, dict):
            raise ValueError(f'The derivatives to extract has not been set properly instead a type of {type(self.derivatives)} was found')
        if isinstance(deriv_method, str):
            if deriv_method in ('jet', 'AD'):      # 'AD' keeps tutorial scripts working; it selects the jet kernels
                self.deriv_method = Jet_engine(self.net, self.derivatives, **kwargs)
            elif deriv_method == 'FD':
                kwargs.setdefault('dxs', [1e-3 for _ in range(len(self.input_vars))])      # DE_module.py:138
                self.deriv_method = FD_engine(self.net, self.derivatives, **kwargs)
            else:
                raise NotImplementedError(f"derivative method '{deriv_method}' does not exist ('jet' | 'AD' | 'FD')")
        elif isinstance(deriv_method, _EngineBase) or hasattr(deriv_method, 'calculate'):
            self.deriv_method = deriv_method
        else:
            raise TypeError('deriv_method should be an engine class or appropriate string')

    def calculate(self, x, **kwargs) -> dict:
        return self.deriv_method.calculate(x, **kwargs)

    def __call__(self, *args, **kwds) -> dict:
        return self.calculate(*args, **kwds)
