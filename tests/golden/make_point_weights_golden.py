"""Generates tests/golden/point_weights.npz from the UNMODIFIED reference (run in the build container only):
    python tests/golden/make_point_weights_golden.py
RegularGridInterpolator (utils/GridInterpolator.py) on a non-uniform 2-D grid incl. the edge cases the reference handles
(outside the grid, exactly on grid lines, on a corner), and R3_sampler.RRR_sample (utils/sampling.py) with a deterministic
resampler, with and without the causal gate."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(__file__))
import ref_shim  # noqa: E402

ref_shim.install()

from torch_DE.utils.GridInterpolator import RegularGridInterpolator
from torch_DE.utils.sampling import R3_sampler

torch.manual_seed(5)
gx = torch.tensor([-1.0, -0.5, 0.1, 0.2, 0.9, 1.5])
gy = torch.cumsum(torch.rand(9) * 0.3 + 0.05, 0) - 1.0
vals = torch.randn(6, 9)
q = torch.rand(400, 2) * torch.tensor([3.2, 3.6]) - torch.tensor([1.4, 1.6])
q[:6, 0] = gx                      # exactly on grid lines
q[6:15, 1] = gy
q[15] = torch.tensor([gx[0], gy[0]])
q[16] = torch.tensor([gx[-1], gy[-1]])
interp = RegularGridInterpolator([gx, gy], vals)
out = interp(q[:, 0].clone(), q[:, 1].clone())

out_r3 = {}
for name, kw in (('plain', {}), ('gate', {'time_interval': (0.0, 2.0)})):
    x = torch.rand(300, 3) * 2.0
    res = [torch.randn(300), torch.randn(300) * 0.3]
    fresh = torch.full((300, 3), -7.0)
    fresh[:, -1] = torch.linspace(0, 1, 300)
    sampler = R3_sampler(lambda n, fresh=fresh: fresh[:n], **kw)
    new = sampler(x, res)
    out_r3[name] = dict(x=x.numpy(), r0=res[0].numpy(), r1=res[1].numpy(), fresh=fresh.numpy(), new=new.numpy(),
                        gamma=np.float32(sampler.gamma))
np.savez_compressed(os.path.join(os.path.dirname(__file__), 'point_weights.npz'), gx=gx.numpy(), gy=gy.numpy(), vals=vals.numpy(),
                    q=q.numpy(), out=out.numpy(),
                    **{f'{k}_{f}': v for k, d in out_r3.items() for f, v in d.items()})
print('written', out.shape, {k: d['new'].shape for k, d in out_r3.items()})
