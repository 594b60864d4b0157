"""World-size-2 gloo test of the data-parallel plumbing (FlatGrads aliasing + one all-reduce of the flat gradient)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torch_de_b200.dp import FlatGrads, shard


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(2, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    flat = FlatGrads(net)
    x_all = torch.rand(64, 2, generator=torch.Generator().manual_seed(1))
    x = shard(x_all, rank, world)
    net(x).pow(2).mean().backward()           # local mean over the shard
    assert flat.check_aliasing()
    flat.all_reduce_mean()
    if rank == 0:
        ret['flat'] = flat.flat.clone()
    flat.zero()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in net.parameters())
    # a caller that resets the gradients to None (optimizer.zero_grad() default) must not end up reducing a stale buffer
    for p in net.parameters():
        p.grad = None
    net(x).pow(2).mean().backward()
    assert not flat.check_aliasing()
    flat.all_reduce_mean()
    assert flat.check_aliasing()
    if rank == 0:
        ret['flat_after_reset'] = flat.flat.clone()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_equals_global_mean():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29517, ret), nprocs=world, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(2, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    x_all = torch.rand(64, 2, generator=torch.Generator().manual_seed(1))
    net(x_all).pow(2).mean().backward()       # global mean
    ref = torch.cat([p.grad.flatten() for p in net.parameters()])
    assert torch.allclose(ret['flat'], ref, rtol=1e-5, atol=1e-7)
    assert torch.allclose(ret['flat_after_reset'], ref, rtol=1e-5, atol=1e-7)


def test_proportional_shards_equalise_the_step_time():
    from torch_de_b200.dp import proportional_shards
    sizes = [500_000] * 4
    ms = [40.0, 41.0, 42.0, 40.5]                       # rank 2 is the slowest
    new = proportional_shards(ms, sizes)
    assert sum(new) == sum(sizes) and all(n % 32 == 0 for n in new[:0] + new[1:])     # (the remainder lands on the fastest rank)
    t_new = [n / (s / t) for n, s, t in zip(new, sizes, ms)]
    assert max(t_new) - min(t_new) < 0.01 and max(t_new) < max(ms)
    assert new[2] < new[1] < new[3] < new[0]
