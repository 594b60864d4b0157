"""Times tde_grid_interp2d (bilinear SDF / point-weight lookup) against its HBM roofline: 12 algorithmic bytes per point
(x, y in, weight out; the grid itself stays in L2).  usage: python tools/grid_interp_bench.py [n_points]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import torch_de_b200 as T

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
g = torch.Generator().manual_seed(0)
gx = torch.cumsum(torch.rand(513, generator=g) + 0.01, 0)
gy = torch.cumsum(torch.rand(513, generator=g) + 0.01, 0)
it = T.RegularGridInterpolator((gx, gy), torch.randn(513, 513, generator=g))
it.set_device('cuda')
x = (torch.rand(n, device='cuda') * (gx[-1] - gx[0]) + gx[0]).contiguous()
y = (torch.rand(n, device='cuda') * (gy[-1] - gy[0]) + gy[0]).contiguous()
for _ in range(3):
    it(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 20
for _ in range(reps):
    it(x, y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
peaks = os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')
hbm = json.load(open(peaks)).get('hbm_gbs', 6500.0) if os.path.exists(peaks) else 6500.0
gbs = 12.0 * n / (ms * 1e-3) * 1e-9
print(json.dumps({"kernel": "grid_interp2d_kernel", "points": n, "ms": ms, "points_per_s": n / (ms * 1e-3), "achieved_GBs": gbs,
                  "peak_GBs": hbm, "frac": gbs / hbm}))
