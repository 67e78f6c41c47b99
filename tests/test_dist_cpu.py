"""World-size-2 gloo checks (CPU) of the multi-process logic: the flat gradient bucket's all-reduce average used by the
training step, and the image-index sharding used by the eval bench (no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tgsr_b200.train_step import FlatBucket
        torch.manual_seed(0)                                   # same weights on every rank, as replicas start
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, bias=False), torch.nn.BatchNorm2d(4), torch.nn.Linear(5, 2))
        w0 = [p.detach().clone() for p in net.parameters()]
        bucket = FlatBucket([net])
        for p, w in zip(net.parameters(), w0):                 # re-homing keeps the values
            assert torch.equal(p.detach(), w)
        bucket.zero_grad()
        x = torch.full((2, 3, 5, 5), float(rank + 1))
        net[0](x).sum().backward()                             # autograd accumulates straight into the flat buffer
        net[2](torch.ones(1, 5) * (rank + 1)).sum().backward()
        local = bucket.grad.clone()
        bucket.all_reduce()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(bucket.grad, sum(gathered))      # SUM; the 1/world average is the optimiser's grad_scale
        assert net[0].weight.grad.data_ptr() == bucket.grad.data_ptr()
        # eval sharding: contiguous image-index blocks, concatenation in rank order reproduces the global batch
        total = 10
        idx = torch.arange(total)
        per = total // world
        mine = idx[rank * per:(rank + 1) * per]
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        assert torch.equal(torch.cat(parts), idx)
        if rank == 0:
            out.put('ok')
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_gloo_world2():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 'ok'
