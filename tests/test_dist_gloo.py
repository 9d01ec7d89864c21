"""world_size-2 gloo test (CPU) of the multi-GPU host logic: pair sharding covers the batch exactly once and
the bucketed gradient all-reduce averages across ranks."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import anystereo_b200 as A
    lo, hi = A.shard_pairs(7, rank, world)
    owned = torch.zeros(7)
    owned[lo:hi] = 1
    dist.all_reduce(owned)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 1000, 3)]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    A.allreduce_gradients(params, bucket_bytes=2048)
    ok = bool((owned == 1).all())
    for i, p in enumerate(params):
        ok = ok and torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1)))
    # ranks that skipped a branch (grad None on ONE rank only), mixed dtypes, explicit world_size
    ps = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(6, dtype=torch.float64)),
          torch.nn.Parameter(torch.zeros(3))]
    ps[0].grad = torch.full_like(ps[0], float(rank + 1))
    ps[1].grad = torch.full_like(ps[1], 2.0 * (rank + 1))
    if rank == 0:
        ps[2].grad = torch.full_like(ps[2], 8.0)          # rank 1 has no gradient for this parameter
    A.allreduce_gradients(ps, world_size=world, bucket_bytes=1 << 20)
    ok = ok and torch.allclose(ps[0].grad, torch.full_like(ps[0], 1.5))
    ok = ok and ps[1].grad.dtype == torch.float64 and torch.allclose(ps[1].grad, torch.full_like(ps[1], 3.0))
    ok = ok and ps[2].grad is not None and torch.allclose(ps[2].grad, torch.full_like(ps[2], 4.0))
    # overlapped reducer: hooks fire during backward, finish() completes; equals the plain average
    torch.manual_seed(1)
    w1 = torch.nn.Parameter(torch.randn(8, 8))
    w2 = torch.nn.Parameter(torch.randn(8, 8))
    w3 = torch.nn.Parameter(torch.randn(8))                 # unused on this step: reduced as zeros
    red = A.GradientAllReducer([w1, w2, w3], bucket_bytes=64)
    x = torch.full((2, 8), float(rank + 1))
    h = x
    for _ in range(3):                                     # a parameter used several times (unrolled iterations)
        h = torch.tanh(h @ w1) @ w2
    h.sum().backward()
    local = [w1.grad.clone(), w2.grad.clone()]
    n_async = red.finish()
    gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
    for g, lst in zip(local, gathered):
        dist.all_gather(lst, g)
    ok = ok and n_async == len(red.buckets)
    ok = ok and torch.allclose(w1.grad, sum(gathered[0]) / world, atol=1e-6)
    ok = ok and torch.allclose(w2.grad, sum(gathered[1]) / world, atol=1e-6)
    ok = ok and w3.grad is not None and float(w3.grad.abs().sum()) == 0.0
    red.remove()
    q.put((rank, ok, (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res
    assert sorted(r[2] for r in res) == [(0, 4), (4, 7)]
