"""CPU, world_size 2, gloo: the data-parallel gradient averaging used by train_gan.train_step
(GradAllReduce: flat buffers, reverse-order buckets, hook-driven async all-reduce) gives every rank
the mean of the per-rank gradients, with BN statistics left per-rank like the reference's
DataParallel replicas."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mcgm_b200.train_gan import GradAllReduce
    torch.manual_seed(0)                                   # identical weights on both ranks
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(),
                              torch.nn.Linear(32, 4))
    red = GradAllReduce(net, bucket_bytes=2048)            # several buckets
    assert len(red.buckets) >= 2
    results = []
    for step in range(2):
        red.zero_grad()
        g = torch.Generator().manual_seed(100 + 10 * step + rank)   # per-rank shard
        x = torch.randn(8, 16, generator=g)
        net(x).square().mean().backward()
        red.finish()
        results.append([p.grad.clone() for p in net.parameters()])
        # the grads stay views of the flat buffer (zero_grad is one memset)
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(net.parameters(), red.views))
    # reference: average of both shards' gradients computed locally
    ok = True
    for step in range(2):
        ref = None
        for r in range(world):
            torch.manual_seed(0)
            net2 = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32),
                                       torch.nn.ReLU(), torch.nn.Linear(32, 4))
            g = torch.Generator().manual_seed(100 + 10 * step + r)
            net2(torch.randn(8, 16, generator=g)).square().mean().backward()
            gr = [p.grad for p in net2.parameters()]
            ref = gr if ref is None else [a + b for a, b in zip(ref, gr)]
        for got, want in zip(results[step], ref):
            ok &= torch.allclose(got, want / world, rtol=1e-5, atol=1e-7)
    # a disabled reducer (the net that receives no useful gradient in this sub-step) stays local
    red.zero_grad()
    red.enabled = False
    net(torch.full((2, 16), float(rank + 1))).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    red.enabled = True
    gathered = [torch.zeros_like(local[0]) for _ in range(world)]
    dist.all_gather(gathered, local[0])
    ok &= not torch.allclose(gathered[0], gathered[1])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_allreduce_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
