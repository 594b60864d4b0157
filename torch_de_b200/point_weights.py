"""Device-side point weights and resampling.

* `RegularGridInterpolator` -- bilinear lookup of a field given on a regular 2-D grid (signed-distance / point-weight
  maps), same constructor / `set_device` / `__call__(x, y)` / `interpolate([x, y])` as the reference
  (torch_DE/utils/GridInterpolator.py:11-83); the lookup is one launch of `tde_grid_interp2d` on the device the points
  live on, no host round trip.
* `R3_sampler` -- retain / resample / release of collocation points from the per-point residuals the loss kernel already
  emits (torch_DE/utils/sampling.py:19-112): everything stays on the device of `x`.
* `sample_from_tensor` -- sampling.py:12-17.
"""
import ctypes as C
from typing import Callable, Optional, Union

import torch

from . import _lib


def sample_from_tensor(num_points: int, t: torch.Tensor, causal: bool = False, time_col: int = -1) -> torch.Tensor:
    """`num_points` rows of `t` drawn with replacement; `causal` sorts them by the time column (sampling.py:12-17)."""
    picked = t[torch.randint(low=0, high=t.shape[0], size=(num_points,), device=t.device)]
    return picked if not causal else picked[picked[:, time_col].sort()[-1]]


class RegularGridInterpolator:
    """2-D linear interpolation on a regular grid: `points = (gx[nx], gy[ny])` ascending, `values[nx, ny]`."""

    def __init__(self, points, values):
        assert isinstance(points, (tuple, list))
        assert isinstance(values, torch.Tensor)
        self.points = [p for p in points]
        self.values = values
        self.ms = list(values.shape)
        self.n = len(self.points)
        assert len(self.ms) == self.n
        if self.n != 2:
            raise NotImplementedError("the device lookup covers the 2-D grids the reference uses (GridInterpolator.py:17)")
        for i, p in enumerate(self.points):
            assert isinstance(p, torch.Tensor)
            assert p.shape[0] == self.values.shape[i]
        self.device = 'cpu'

    def set_device(self, device):
        """Move the grid to `device` (GridInterpolator.py:36-43)."""
        self.points = [p.to(device=device) for p in self.points]
        self.values = self.values.to(device=device)
        self.device = device

    def __call__(self, x, y: torch.Tensor, **kwargs):
        from . import fused_loss as FL
        if isinstance(x, FL.Sym) or isinstance(y, FL.Sym):
            # called by the loss tracer on symbolic coordinates (a weighting such as DE_func(domain.create_sdf()),
            # Stokes_Cylinder.py:34,79): the lookup becomes one op of the fused loss kernel instead of a launch + an aux row
            return FL.grid_lookup(self, x, y)
        return self.interpolate([x, y])

    def interpolate(self, points_to_interp):
        assert self.points is not None and self.values is not None
        assert len(points_to_interp) == len(self.points)
        x, y = points_to_interp
        assert x.shape[0] == y.shape[0]
        if not x.is_cuda:
            raise RuntimeError("RegularGridInterpolator runs on CUDA tensors only (sm_100a kernel; there is no CPU fallback)")
        with torch.no_grad():
            gx, gy = (p.to(device=x.device, dtype=torch.float32).contiguous() for p in self.points)
            vals = self.values.to(device=x.device, dtype=torch.float32).contiguous()
            xq = x.detach().to(torch.float32).contiguous()
            yq = y.detach().to(torch.float32).contiguous()
            out = torch.empty_like(xq)
            rc = _lib.lib().tde_grid_interp2d(gx.data_ptr(), gx.shape[0], gy.data_ptr(), gy.shape[0], vals.data_ptr(),
                                              xq.data_ptr(), yq.data_ptr(), xq.numel(), out.data_ptr(),
                                              torch.cuda.current_stream(x.device).cuda_stream)
            _lib.check(rc, "tde_grid_interp2d")
            return out


class R3_sampler:
    """Retain-Resample-Release sampler (reference: utils/sampling.py:19-112), on the device of the points.

    Points whose residual measure (sum of |residual| rows, times the causal gate) reaches the batch mean are kept, the rest are
    replaced by `sampler(n_missing, **kwargs)`.  `time_interval=(t0, t1)` switches the causal gate on (a tanh ramp in t / t1
    whose position `gamma` advances as the mean measure falls); `causal=True` returns the released set sorted by its last
    (time) column.  Attribute names follow the reference so user code that tunes `alpha / gamma / nu / eps / dmax` keeps working."""

    def __init__(self, sampler: Callable, *, group: Optional[str] = None, device: str = 'cpu', causal: bool = False,
                 time_interval=None) -> None:
        self.sampler, self.group, self.device, self.causal, self.time_interval = sampler, group, device, causal, time_interval
        self.alpha, self.gamma, self.nu, self.eps, self.dmax = 5, -0.5, 1e-3, 20, 0.1
        self._plot_args = None
        self.g = self._gate if time_interval is not None else (lambda t, gamma: t)

    def _gate(self, t: torch.Tensor, gamma) -> torch.Tensor:
        return 0.5 * (1 - torch.tanh(self.alpha * (t / self.time_interval[-1] - gamma)))

    @staticmethod
    def F_measure(*res, device='cpu') -> torch.Tensor:
        total = res[0].abs()
        for r in res[1:]:
            total = total + r.abs()
        return total.to(device)

    def _residual_rows(self, res, loss_type):
        if not hasattr(res, 'residual_rows'):
            if not (hasattr(res, 'weighted_error') and hasattr(res, 'point_error')):
                return res                                    # already an iterable of per-point residual vectors
            # an object with the table accessors the reference's sampler calls (sampling.py:72-76)
            if loss_type not in ('weighted', 'point error'):
                raise ValueError('loss_type accepts only strings weighted and point error')
            table = res.weighted_error() if loss_type == 'weighted' else res.point_error()
            return list(table['Residual'][self.group].values())
        if loss_type not in ('weighted', 'point error'):
            raise ValueError('loss_type accepts only strings weighted and point error')
        # a Loss object: the 'residual' rows of my group (sampling.py:72-76); under the fused loss path the per-point values
        # come from the loss kernel's point pass
        rows = res.residual_rows(self.group, weighted=loss_type == 'weighted')
        if not rows:
            raise KeyError(f"the loss table has no 'residual' rows for group {self.group!r}")
        return rows

    def retain(self, x, Res, mean):
        return x[Res >= mean]

    def resample(self, num_points, **kwargs):
        return self.sampler(num_points, **kwargs)

    @torch.no_grad()
    def RRR_sample(self, x: Union[dict, torch.Tensor], res, loss_type='weighted', **kwargs) -> torch.Tensor:
        rows = self._residual_rows(res, loss_type)
        pts = x[self.group] if isinstance(x, dict) else x
        measure = self.F_measure(*rows, device=self.device) * self.g(pts[:, -1], self.gamma).to(self.device)
        pts = pts.to(self.device)
        level = measure.mean()
        kept = self.retain(pts, measure, level)
        fresh = self.resample(pts.shape[0] - kept.shape[0], **kwargs)
        self._plot_args = [kept, fresh, measure]
        if self.time_interval is not None:
            self.gamma = self.gamma + self.nu * min(torch.exp(-self.eps * level), self.dmax)
        released = torch.cat([kept, fresh.to(self.device)], dim=0)
        return released[released[:, -1].sort()[1]] if self.causal else released

    __call__ = RRR_sample
