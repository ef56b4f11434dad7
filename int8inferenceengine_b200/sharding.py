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


class ResultExchange:
    """Fused result exchange of one sharded step on GPUs: one pack kernel (top-1 agreement
    count + logits behind a 16-byte header), ONE `all_gather_into_tensor` over NCCL, one unpack
    kernel. Replaces all-gather(logits) + all-reduce(count) and the handful of tiny torch kernels
    that computed the count; handles uneven shards (each chunk carries its row count).
    CUDA only — the kernels live in libi8ie_sm100.so (include/i8ie_sm100.h, i8ie_top1_*)."""

    def __init__(self, global_batch: int, cols: int, device, overlap: bool = False):
        """overlap=True: the all-gather and the unpack run on a side stream (double-buffered chunks),
        so the forward of the next step never waits for a peer; results are valid after `wait()`."""
        from . import _lib
        self._L = _lib.load()
        self._check = _lib.check
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.global_batch, self.cols = int(global_batch), int(cols)
        spans = [shard_range(global_batch, r, self.world) for r in range(self.world)]
        self.rows = spans[self.rank][1] - spans[self.rank][0]
        cap = max(hi - lo for lo, hi in spans)
        while (cap * cols) % 4:
            cap += 1
        self.chunk = int(self._L.i8ie_top1_chunk_bytes(cap, cols))
        nbuf = 2 if overlap else 1
        self._packed = [torch.zeros(self.chunk, dtype=torch.uint8, device=device) for _ in range(nbuf)]
        self._gathered = [torch.zeros(self.world * self.chunk, dtype=torch.uint8, device=device) for _ in range(nbuf)]
        self.packed, self.gathered = self._packed[0], self._gathered[0]
        self.logits_all = torch.empty(self.global_batch, cols, dtype=torch.float32, device=device)
        self.agree = torch.zeros(1, dtype=torch.int64, device=device)
        self._side = torch.cuda.Stream(device=device) if overlap else None
        self._done = [None] * nbuf     # side-stream events: chunk buffers free / results written
        self._calls = 0

    def wait(self):
        """Make the current stream wait for every exchange issued so far (overlap mode)."""
        for ev in self._done:
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)

    def __call__(self, local_logits: torch.Tensor, ref_argmax: torch.Tensor | None = None):
        """local_logits: [rows, cols] fp32 CUDA (contiguous); ref_argmax: [rows] int64 CUDA or None.
        Returns (logits of all ranks [global_batch, cols], agreement count tensor [1] int64) —
        static buffers, valid until the next call; everything is stream-ordered, nothing syncs
        (overlap mode: valid once `wait()` has been called on the consuming stream)."""
        if not local_logits.is_cuda or local_logits.dtype != torch.float32 or not local_logits.is_contiguous():
            raise ValueError("ResultExchange needs a contiguous fp32 CUDA tensor")
        if tuple(local_logits.shape) != (self.rows, self.cols):
            raise ValueError(f"expected logits of shape {(self.rows, self.cols)}, got {tuple(local_logits.shape)}")
        cur = torch.cuda.current_stream()
        ref_ptr = None
        if ref_argmax is not None:
            if ref_argmax.dtype != torch.int64 or ref_argmax.numel() != self.rows or not ref_argmax.is_cuda:
                raise ValueError("ref_argmax must be int64 CUDA with one entry per local row")
            ref_ptr = ref_argmax.data_ptr()
        k = self._calls % len(self._packed)
        self._calls += 1
        packed, gathered = self._packed[k], self._gathered[k]
        if self._done[k] is not None:      # the exchange that last used this chunk buffer has finished
            cur.wait_event(self._done[k])
        # pack on the caller's stream: it consumes the logits there (local work, no peer involved)
        self._check(self._L.i8ie_top1_pack(local_logits.data_ptr(), ref_ptr, self.rows, self.cols,
                                           packed.data_ptr(), cur.cuda_stream), "top1_pack")
        run_on = cur
        if self._side is not None:
            self._side.wait_event(cur.record_event())
            run_on = self._side
        with torch.cuda.stream(run_on):
            if self.world > 1:
                dist.all_gather_into_tensor(gathered, packed)
                src = gathered
            else:
                src = packed
            self._check(self._L.i8ie_top1_unpack(src.data_ptr(), self.world, self.chunk, self.logits_all.data_ptr(),
                                                 self.agree.data_ptr(), run_on.cuda_stream), "top1_unpack")
            if self._side is not None:
                self._done[k] = run_on.record_event()
        return self.logits_all, self.agree
