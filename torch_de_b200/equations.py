"""Residual adapters (host side of SURVEY.md 8a rows a9-a10).

`DE_func` keeps the calling convention of src/torch_DE/Equations/de_func.py:5-22: the wrapped function is
called as f(**input.batchables, **input.unbatchables, **output) and must accept **kwargs.
`get_derivatives` lists the derivative names a residual's signature asks for (Equations/Get_Derivatives.py:2-37).
The Navier-Stokes residuals follow Equations/NavierStokes.py:8-80.
"""
import inspect
from functools import partial, wraps
from typing import Callable, Dict, List, Sequence, Tuple


def DE_func(func: Callable) -> Callable:
    kinds = [p.kind for p in inspect.signature(func).parameters.values()]
    assert inspect.Parameter.VAR_KEYWORD in kinds, \
        'To use DE_func, you need to have the **kwargs declared in your function input e.g foo(x,y,**kwargs)'

    @wraps(func)
    def wrapper(network_input, network_output):
        return func(**network_input.batchables, **network_input.unbatchables, **network_output)
    wrapper.is_decorated = True
    wrapper.base_func = func
    return wrapper


def get_derivatives(input_vars: Sequence[str], output_vars: Sequence[str], *equations, merge: bool = True) -> tuple:
    ins, outs = set(input_vars), set(output_vars)
    per_eq = []
    for eq in equations:
        wanted = set()
        for name in inspect.signature(eq).parameters:
            if '_' not in name:
                continue
            parts = name.split('_')
            if len(parts) != 2 or parts[0] not in outs or not parts[1] or any(v not in ins for v in parts[1]):
                continue
            wanted.add(name)
        per_eq.append(wanted)
    if merge:
        return tuple(set().union(*per_eq)) if per_eq else tuple()
    return tuple(per_eq)


def NavierStokes_x(u, v, w, u_x, u_y, u_z, u_t, u_xx, u_yy, u_zz, p_x, Re, **kwargs):
    return u_t + u * u_x + v * u_y + w * u_z + p_x - 1 / Re * (u_xx + u_yy + u_zz)


def NavierStokes_y(u, v, w, v_x, v_y, v_z, v_t, v_xx, v_yy, v_zz, p_y, Re, **kwargs):
    return v_t + u * v_x + v * v_y + w * v_z + p_y - 1 / Re * (v_xx + v_yy + v_zz)


def NavierStokes_z(u, v, w, w_x, w_y, w_z, w_t, w_xx, w_yy, w_zz, p_z, Re, **kwargs):
    return w_t + u * w_x + v * w_y + w * w_z + p_z - 1 / Re * (w_xx + w_yy + w_zz)


def incompressible(u_x, v_y, w_z, **kwargs):
    return u_x + v_y + w_z


def get_NavierStokes(dims: int = 2, steady_state: bool = False, Re=None) -> Tuple[tuple, tuple, tuple, Dict[str, Callable]]:
    if dims > 3 or dims < 1:
        raise ValueError('Dims can only be int type and between 1 and 3')
    frozen = {'u_t': 0, 'v_t': 0, 'w_t': 0} if steady_state else {}
    eqs = [partial(f, **frozen, Re=Re) for f in (NavierStokes_x, NavierStokes_y, NavierStokes_z)]
    inc = incompressible
    if dims == 2:
        eqs[0] = partial(eqs[0], w=0, u_z=0, u_zz=0)
        eqs[1] = partial(eqs[1], w=0, v_z=0, v_zz=0)
        inc = partial(incompressible, w_z=0)
    elif dims == 1:
        eqs[0] = partial(eqs[0], v=0, w=0, u_z=0, u_y=0, u_yy=0, u_zz=0)
        inc = partial(incompressible, v_y=0, w_z=0)
    input_vars = ('x', 'y', 'z')[:dims] + (tuple() if steady_state else ('t',))
    output_vars = ('u', 'v', 'w')[:dims] + ('p',)
    names = ('NavierStokes_x', 'NavierStokes_y', 'NavierStokes_z')[:dims] + ('incompressible',)
    equations = tuple(eqs[:dims]) + (inc,)
    derivatives = get_derivatives(input_vars, output_vars, *equations)
    return input_vars, output_vars, derivatives, {n: DE_func(e) for n, e in zip(names, equations)}
