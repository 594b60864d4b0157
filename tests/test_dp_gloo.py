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
