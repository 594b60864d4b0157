"""Whole-step CUDA graph for launch-bound configurations.

A PINN step on a small network (Spring, Burgers: 0.06-0.6 ms of kernels) is dominated by host work: ~15 launches through
Python, autograd and ctypes cost ~1 ms.  Every launch of this package goes to `torch.cuda.current_stream()` with no host
synchronisation, host->device copy or allocation outside torch's caching allocator, so the step -- `DE_Getter.calculate`,
the fused residual/loss kernels, `backward()`, the optimizer -- can be captured once and replayed.

    opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)
    def step():
        opt.zero_grad(set_to_none=True)
        out = pinn.calculate(batch)              # batch: tensors that stay at the same addresses (refill them in place)
        total = combine(losses(batch, out))
        total.backward()
        opt.step()
        return total
    graphed = GraphedStep(step)
    for it in range(n): loss = graphed()         # `loss` is a device tensor overwritten by every replay
"""
from typing import Callable

import torch


class GraphedStep:
    def __init__(self, step: Callable[[], torch.Tensor], warmup: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up on a side stream: plans, chunking decisions and allocator pools settle
            for _ in range(warmup):
                step()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = step()

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        return self.result
