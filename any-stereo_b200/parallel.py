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


def _buckets(params, bucket_bytes):
    """Deterministic buckets: parameters in the given (model) order, split by dtype, closed at ``bucket_bytes``."""
    out, cur, size, dt = [], [], 0, None
    for p in params:
        if cur and (p.dtype != dt or size >= bucket_bytes):
            out.append(cur)
            cur, size = [], 0
        dt = p.dtype
        cur.append(p)
        size += p.numel() * p.element_size()
    if cur:
        out.append(cur)
    return out


def _reduce_bucket(bucket, ws, async_op=False):
    flat = torch.cat([p.grad.reshape(-1) for p in bucket])
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=async_op)

    def finish():
        if work is not None:
            work.wait()
        flat.div_(ws)
        off = 0
        for p in bucket:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
    return finish


def allreduce_gradients(params, world_size=None, bucket_bytes=32 << 20):
    """Average gradients across ranks with bucketed all-reduce (NCCL on GPUs, gloo in CPU tests).

    The set of buckets is a function of the parameter LIST only, never of which gradients happen to exist on this
    rank: a parameter whose branch was skipped (``p.grad is None``; e.g. ``iter16=False`` on one rank) contributes
    zeros, so every rank issues the same collectives with the same sizes.  Buckets never mix dtypes.
    The model is ~12.5 M parameters (50 MB fp32): a handful of buckets, sized for launch latency rather than
    link count (NVSwitch gives uniform peer bandwidth)."""
    if not dist.is_initialized():
        return
    ws = dist.get_world_size() if world_size is None else int(world_size)
    if ws < 1:
        raise ValueError("world_size must be >= 1")
    params = [p for p in params if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    for bucket in _buckets(params, bucket_bytes):
        _reduce_bucket(bucket, ws)()


class GradientAllReducer:
    """DDP-style overlap of the gradient all-reduce with the backward pass (replaces nn.DataParallel's reduce-to-GPU0,
    train_continuous_IGEV.py:184).

    A post-accumulate hook on every parameter marks it ready; when the last parameter of a bucket is ready the bucket's
    all-reduce is launched asynchronously (NCCL runs it on its own stream, concurrently with the remaining backward
    kernels).  ``finish()`` -- call it after ``loss.backward()`` -- reduces whatever was not triggered (parameters
    without a gradient this step are reduced as zeros, so all ranks issue identical collectives), waits, averages and
    writes the results back into ``p.grad``.

    In the unrolled 16-iteration graph every parameter's gradient is complete only once the FIRST iteration has been
    back-propagated, so at most the last 1/16 of the backward pass can overlap; the buckets are ordered the way that
    iteration's backward completes them (``disp_head`` -> ``gru04`` -> ``encoder`` -> ``gru08`` -> ``gru16``)."""

    def __init__(self, params, world_size=None, bucket_bytes=4 << 20):
        self.ws = dist.get_world_size() if (world_size is None and dist.is_initialized()) else int(world_size or 1)
        self.params = [p for p in params if p.requires_grad]
        self.buckets = _buckets(self.params, bucket_bytes)
        self._where = {}
        for bi, b in enumerate(self.buckets):
            for p in b:
                self._where[id(p)] = bi
        self._pending = [len(b) for b in self.buckets]
        self._launched = {}
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def _hook(self, p):
        bi = self._where[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and dist.is_initialized() and self.ws > 1:
            self._launched[bi] = _reduce_bucket(self.buckets[bi], self.ws, async_op=True)

    def finish(self):
        if dist.is_initialized() and self.ws > 1:
            for bi, b in enumerate(self.buckets):
                if bi not in self._launched:
                    for p in b:
                        if p.grad is None:
                            p.grad = torch.zeros_like(p)
                    self._launched[bi] = _reduce_bucket(b, self.ws, async_op=True)
            for bi in sorted(self._launched):
                self._launched[bi]()
        overlapped = len(self._launched)
        self._launched = {}
        self._pending = [len(b) for b in self.buckets]
        return overlapped

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
