"""Multi-GPU plumbing: the hot path shards by stereo pair.

Inference: every pair is independent through the whole path, so ranks take disjoint slices of the batch and
there is NO data-path collective.  Training (config 5): identical replicas, one NCCL all-reduce of the
gradients per step (what the reference gets implicitly from nn.DataParallel, train_continuous_IGEV.py:184).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_pairs(num_pairs: int, rank: int, world_size: int):
    """Contiguous, balanced slice [lo, hi) of the pair indices owned by ``rank``."""
    if world_size < 1 or not (0 <= rank < world_size) or num_pairs < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(num_pairs, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def allreduce_gradients(params, world_size=None, bucket_bytes=32 << 20):
    """Average gradients across ranks with bucketed all-reduce (NCCL on GPUs, gloo in CPU tests).

    The model is ~12.5 M parameters (50 MB fp32): a handful of buckets, sized for launch latency rather than
    link count (NVSwitch gives uniform peer bandwidth)."""
    if not dist.is_initialized():
        return
    ws = world_size or dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(ws)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
