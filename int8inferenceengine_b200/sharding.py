"""Batch sharding of the INT8 forward across the GPUs of one box.

Images are independent (the reference itself parallelises over images, conv2d.cc:125), so the
path shards with NO data-path collective: every rank holds the full (replicated) weights and a
contiguous slice of the batch. The only exchange is the result: an all-gather of the fp32
logits and an all-reduce(SUM) of the top-1 agreement count — both tiny, latency-bound.
Backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of the batch owned by `rank`; the first `global_batch % world`
    ranks take one extra image."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(int(global_batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_logits(local_logits: torch.Tensor, global_batch: int, out: torch.Tensor | None = None):
    """All-gather the per-rank logits [b_r, C] into [global_batch, C] in rank order (handles
    uneven shards by padding to the largest shard)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local_logits
    rank = dist.get_rank()
    c = local_logits.shape[1]
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == bmax for lo, hi in sizes):
        if out is None:
            out = torch.empty(world * bmax, c, dtype=local_logits.dtype, device=local_logits.device)
        dist.all_gather_into_tensor(out, local_logits.contiguous())
        return out
    padded = torch.zeros(bmax, c, dtype=local_logits.dtype, device=local_logits.device)
    padded[: local_logits.shape[0]] = local_logits
    buf = torch.empty(world * bmax, c, dtype=local_logits.dtype, device=local_logits.device)
    dist.all_gather_into_tensor(buf, padded)
    parts = [buf[r * bmax: r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    del rank
    return torch.cat(parts, 0)


def reduce_count(local_count: int, device="cpu") -> int:
    """all-reduce(SUM) of an int64 count (e.g. top-1 agreement with a reference argmax)."""
    t = torch.tensor([int(local_count)], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
