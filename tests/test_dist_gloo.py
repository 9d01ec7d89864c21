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
