"""Eager vs CUDA-graph replay of the same training step (torch_de_b200.GraphedStep) for the launch-bound configurations.
usage: python tools/graph_bench.py [workload ...]      (default: spring burgers allen_cahn)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import torch_de_b200 as T
from torch_de_b200 import workloads as WL

for name in (sys.argv[1:] or ['spring', 'burgers', 'allen_cahn']):
    wl = WL.WORKLOADS[name]()
    n_col, n_bnd = max(1, wl.default_points // 8), max(1, wl.default_boundary // 8)
    torch.manual_seed(1234)
    net = wl.make_net().cuda()
    pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
    groups = {g: x.cuda() for g, x in wl.sample(n_col, n_bnd, seed=1234).items()}
    batch = T.make_batch(wl.input_vars, groups)
    losses = wl.build_losses(list(groups.keys()))
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)

    def step():
        opt.zero_grad(set_to_none=True)
        out = pinn.calculate(batch)
        total = wl.combine(losses(batch, out).grouped_losses('loss_type'))
        total.backward()
        opt.step()
        return total

    def time_it(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    eager_ms = time_it(step)
    graphed = T.GraphedStep(step)
    graph_ms = time_it(graphed)
    n_pts = sum(x.shape[0] for x in groups.values())
    print(json.dumps({"workload": name, "points_per_step": n_pts, "eager_ms": eager_ms, "graph_ms": graph_ms,
                      "eager_points_per_s": n_pts / (eager_ms * 1e-3), "graph_points_per_s": n_pts / (graph_ms * 1e-3),
                      "speedup": eager_ms / graph_ms}))
