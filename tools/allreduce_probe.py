"""How long does the gradient all-reduce take on its own?  (1.85 MB flat gradient of the 8x256 network, NCCL, one process per GPU)
usage: python -m torch.distributed.run --nproc-per-node N tools/allreduce_probe.py"""
import os
import torch
import torch.distributed as dist

rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
for n in (462_339, 4_000_000):
    buf = torch.ones(n, device='cuda')
    for _ in range(5):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f'all_reduce of {n * 4 / 1e6:.2f} MB x {dist.get_world_size()} ranks: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call (back to back)')
    # one call after a long busy kernel on every rank (what the training step does)
    a = torch.randn(8192, 8192, device='cuda')
    ts = []
    for _ in range(5):
        for _ in range(3):
            a @ a
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        dist.all_reduce(buf)
        s1.record()
        torch.cuda.synchronize()
        ts.append(s0.elapsed_time(s1) * 1e3)
    if rank == 0:
        print(f'  after a busy stretch: {sorted(ts)[len(ts) // 2]:.1f} us (median of 5)')
dist.destroy_process_group()
