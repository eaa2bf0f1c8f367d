"""world_size-2 gloo test of the bucketed gradient all-reduce (host logic of the multi-GPU row):
sharded-batch gradients after the all-reduce == full-batch gradients of a single process."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.LayerNorm(64),
                               torch.nn.Linear(64, 8))


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from swin_b200.ddp import BucketedGradAllReduce
    net = _model()
    if rank == 1:                                   # rank 1 starts from different weights: broadcast must fix it
        for p in net.parameters():
            p.data.add_(1.0)
    ddp = BucketedGradAllReduce(net, bucket_mb=0.01, tail_kb=0.5)    # tiny buckets -> several buckets + a small tail bucket
    assert len(ddp.buckets) > 2
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 8, generator=g)
    for it in range(2):
        xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
        loss = ((net(xs) - ys) ** 2).sum() / 8.0     # global-batch mean => average of per-rank (sum/4)... see below
        (loss * world).backward()                    # per-rank loss is scaled so that AVG over ranks == full-batch grad
        ddp.finish()
        grads = [p.grad.clone() for p in net.parameters()]
        if it == 0:
            ddp.zero_grad()
    if rank == 0:
        ret.put([g.numpy() for g in grads])
    dist.destroy_process_group()


def test_bucketed_allreduce_matches_full_batch():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    net = _model()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 8, generator=g)
    (((net(x) - y) ** 2).sum() / 8.0).backward()
    for a, p in zip(got, net.parameters()):
        assert torch.allclose(torch.from_numpy(a), p.grad, rtol=1e-5, atol=1e-6)
