"""N>1 host logic on CPU: world_size-2 gloo run of the env sharding + flat-bucket gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pufferlib_b200 import distributed as pdist


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, l, w = pdist.init(backend='gloo')
    assert (r, w) == (rank, world)
    off, per = pdist.shard(128, rank, world)
    assert (off, per) == (rank * 64, 64)
    torch.manual_seed(0)                      # same init on every rank, then broadcast anyway
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    pdist.broadcast_parameters(model)
    bucket = pdist.GradBucket(model)
    x = torch.full((4, 5), float(rank + 1))
    for step in range(2):
        bucket.zero()
        model(x).sum().backward()
        local = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
        bucket.all_reduce_mean()
        torch.save({'local': local, 'avg': bucket.flat.clone()}, os.path.join(out_dir, f'r{rank}_s{step}.pt'))
    dist.destroy_process_group()


def test_gloo_world2_grad_bucket(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for step in range(2):
        a = torch.load(tmp_path / f'r0_s{step}.pt')
        b = torch.load(tmp_path / f'r1_s{step}.pt')
        assert torch.equal(a['avg'], b['avg'])                                   # every rank holds the same mean
        assert torch.allclose(a['avg'], (a['local'] + b['local']) / 2, atol=1e-6)
        assert not torch.equal(a['local'], b['local'])


def test_shard_validation():
    with pytest.raises(ValueError):
        pdist.shard(10, 0, 3)
    assert pdist.shard(131072, 7, 8) == (7 * 16384, 16384)
