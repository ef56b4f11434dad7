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


class PeerExchange:
    """The same result exchange over NVLink / NVSwitch PEER MEMORY instead of NCCL
    (csrc/peer_exchange.cu): the pack kernel pushes this rank's [count | logits] chunk into every
    rank's gather buffer with peer stores and publishes a step number; the unpack kernel waits for
    all ranks' step numbers and unpacks from local memory. Two kernels, no host call and no NCCL
    collective per step, so a whole step (forward + exchange) is capturable as ONE CUDA graph —
    see ShardedStep. torch.distributed is used once, at construction, to swap the 64-byte CUDA IPC
    handles. Same call contract as ResultExchange."""

    def __init__(self, global_batch: int, cols: int, device):
        import ctypes as C
        from . import _lib
        self._C = C
        self._L = _lib.load()
        self._check = _lib.check
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.global_batch, self.cols = int(global_batch), int(cols)
        self.device = torch.device(device)
        spans = [shard_range(global_batch, r, self.world) for r in range(self.world)]
        self.rows = spans[self.rank][1] - spans[self.rank][0]
        cap = max(hi - lo for lo, hi in spans)
        while (cap * cols) % 4:
            cap += 1
        self.chunk = int(self._L.i8ie_top1_chunk_bytes(cap, cols))
        nbytes = int(self._L.i8ie_peer_exchange_bytes(self.world, self.chunk))
        # Every step below that can fail on ONE rank (allocation, IPC export, mapping a peer) is followed by an
        # exchange of the outcome, so that all ranks raise together instead of some of them waiting in a collective.
        mine = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        err = None
        try:
            with torch.cuda.device(self.device):
                self._check(self._L.i8ie_peer_alloc(nbytes, C.byref(mine), handle), "peer_alloc")
        except Exception as e:  # noqa: BLE001
            err = str(e)
        self._mine = mine.value
        handles = [(err, bytes(handle))]
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, (err, bytes(handle)))
        self._opened = []
        bad = [f"rank {r}: {h[0]}" for r, h in enumerate(handles) if h[0] is not None]
        bases = (C.c_void_p * self.world)()
        if not bad:
            try:
                for p in range(self.world):
                    if p == self.rank:
                        bases[p] = self._mine
                        continue
                    ptr = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(handles[p][1])
                    with torch.cuda.device(self.device):
                        self._check(self._L.i8ie_peer_open(buf, C.byref(ptr)), f"peer_open(rank {p})")
                    bases[p] = ptr.value
                    self._opened.append(ptr.value)
            except Exception as e:  # noqa: BLE001
                err = str(e)
            if self.world > 1:
                outcomes = [None] * self.world
                dist.all_gather_object(outcomes, err)
                bad = [f"rank {r}: {o}" for r, o in enumerate(outcomes) if o is not None]
            elif err is not None:
                bad = [err]
        if bad:
            with torch.cuda.device(self.device):
                for ptr in self._opened:
                    self._L.i8ie_peer_close(ptr)
                if self._mine:
                    self._L.i8ie_peer_free(self._mine)
            self._opened, self._mine = [], None
            raise RuntimeError("peer-memory exchange unavailable (" + "; ".join(bad) + ")")
        self._bases = bases
        self.seq = torch.zeros(2, dtype=torch.int64, device=self.device)    # [0] = step counter
        self.logits_all = torch.empty(self.global_batch, cols, dtype=torch.float32, device=self.device)
        self.agree = torch.zeros(1, dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.barrier()          # every buffer exists, is zeroed and is mapped everywhere before the first push

    def wait(self):
        pass

    def __call__(self, local_logits: torch.Tensor, ref_argmax: torch.Tensor | None = None):
        if not local_logits.is_cuda or local_logits.dtype != torch.float32 or not local_logits.is_contiguous():
            raise ValueError("PeerExchange needs a contiguous fp32 CUDA tensor")
        if tuple(local_logits.shape) != (self.rows, self.cols):
            raise ValueError(f"expected logits of shape {(self.rows, self.cols)}, got {tuple(local_logits.shape)}")
        ref_ptr = None
        if ref_argmax is not None:
            if ref_argmax.dtype != torch.int64 or ref_argmax.numel() != self.rows or not ref_argmax.is_cuda:
                raise ValueError("ref_argmax must be int64 CUDA with one entry per local row")
            ref_ptr = ref_argmax.data_ptr()
        st = torch.cuda.current_stream().cuda_stream
        self._check(self._L.i8ie_top1_pack_push(local_logits.data_ptr(), ref_ptr, self.rows, self.cols, self._bases,
                                                self.world, self.rank, self.chunk, self.seq.data_ptr(), st),
                    "top1_pack_push")
        self._check(self._L.i8ie_top1_wait_unpack(self._mine, self.world, self.chunk, self.seq.data_ptr(),
                                                  self.logits_all.data_ptr(), self.agree.data_ptr(), st),
                    "top1_wait_unpack")
        return self.logits_all, self.agree

    def close(self):
        """Unmaps the peers' buffers and frees the local one (collective: all ranks call it)."""
        if self._mine is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier()          # nobody is still pushing into a buffer that is about to go away
        with torch.cuda.device(self.device):
            for ptr in self._opened:
                self._L.i8ie_peer_close(ptr)
            self._opened = []
            if self.world > 1 and dist.is_initialized():
                dist.barrier()
            self._L.i8ie_peer_free(self._mine)
        self._mine = None


def make_exchange(global_batch: int, cols: int, device, kind: str | None = None):
    """Result exchange for the sharded step: 'peer' (NVLink peer memory, default on GPUs) or 'nccl'
    (ResultExchange). I8IE_EXCHANGE overrides. If the peer buffers cannot be set up (no CUDA IPC in the
    container, no peer access), every rank learns it together and all of them fall back to the NCCL exchange."""
    import os
    import warnings
    kind = kind or os.environ.get("I8IE_EXCHANGE", "peer")
    if kind == "peer":
        try:
            return PeerExchange(global_batch, cols, device)
        except RuntimeError as e:
            warnings.warn(f"i8ie: {e}; using the NCCL result exchange")
            return ResultExchange(global_batch, cols, device)
    if kind == "nccl":
        return ResultExchange(global_batch, cols, device)
    raise ValueError(f"unknown exchange kind {kind!r}")


class ShardedStep:
    """One batch-sharded step = the rank's quantised forward + the result exchange, captured as ONE
    CUDA graph per input buffer: a step is a single host enqueue (graph launch), whatever the number
    of kernels, and nothing on the host sits between the forward and the exchange.

    model     an i8ie.Module (converted); its own graph replay is bypassed (the forward is captured here)
    inputs    list of i8ie.Tensor device batches [rows, ...] (static buffers: replays read them in place)
    ref_args  list of int64 CUDA tensors [rows] (reference argmax per input) or None
    exchange  PeerExchange (captured with the forward). With a ResultExchange (NCCL) only the forward is
              captured and the exchange is enqueued after every replay: an NCCL collective inside the
              capture hung on the B200 box (torch 2.11 / NCCL 2.28), so it is not attempted."""

    def __init__(self, model, inputs, ref_args, exchange):
        self.model, self.inputs, self.exchange = model, list(inputs), exchange
        self.in_graph = isinstance(exchange, PeerExchange)
        self.ref_args = list(ref_args) if ref_args is not None else [None] * len(self.inputs)
        self.graphs = []
        self.kernels_per_step = 0
        self.replays = 0
        from . import _lib
        rows, cols = exchange.rows, exchange.cols
        saved = model.graph
        model.graph = False
        try:
            pool = None
            for x, ref in zip(self.inputs, self.ref_args):
                for _ in range(2):          # eager warm-up: plans, offsets, packed weights, scratch sizes
                    out = model(x)
                    exchange(out.data.buf.view(rows, cols), ref)
                torch.cuda.synchronize()
                before = _lib.launch_count()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    out = model(x)
                    if self.in_graph:
                        exchange(out.data.buf.view(rows, cols), ref)
                self.kernels_per_step = int(_lib.launch_count() - before)
                pool = pool or g.pool()
                self.graphs.append((g, out))
        finally:
            model.graph = saved

    def __call__(self, i: int):
        """Runs step i (input i modulo the ring). Returns the exchange's static result buffers
        (logits of all ranks, agreement count), valid until the next step."""
        j = i % len(self.graphs)
        self.graphs[j][0].replay()
        self.replays += 1
        if not self.in_graph:
            self.exchange(self.graphs[j][1].data.buf.view(self.exchange.rows, self.exchange.cols), self.ref_args[j])
        return self.exchange.logits_all, self.exchange.agree

    def local_logits(self, i: int):
        """The rank's own dequantised logits of the last replay of input i (static graph output)."""
        return self.graphs[i % len(self.graphs)][1].data.buf

    def launches(self):
        return self.replays * self.kernels_per_step
