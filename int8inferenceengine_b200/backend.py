"""Backend object protocol of the B200 engine — the stand-in for the reference's pybind11
module `_CXX_i8ie` (src/pybind11.cc:37-55, conv2d.cc:144-165, fully_connected.cc:54-72,
functional.cc:66-82): functions tensor / quantize / dequantize / relu / max_pool2d, layer
classes Linear / Conv2d, and tensor objects with numpy / zero_point / scale / sum /
ref_count / reshape.

Device memory is owned by torch tensors (glue); every INT8 op is a call into
libi8ie_sm100.so through the C ABI (include/i8ie_sm100.h). There is no CPU fallback.

Layouts. The API is logical NCHW (the reference's, conv2d.cc:64-67). Physically,
  fp32 tensors          dense NCHW
  u8 4-D activations    NHWC with channel pitch cp = round_up(C, 16)
  u8 2-D activations    rows with pitch round_up(K, 16)  (NHWC with H = W = 1)
  other u8 / s8         dense
`numpy()` always returns the logical NCHW array (a D2H copy; the reference's is a
zero-copy view of host memory, pybind11.cc:14-15 — mutation through the view is not
carried over).
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import warnings
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import I8ieError, check

NUM_SAMPLES = 1000  # include/calibrator.h:4

# Bumped whenever something a captured CUDA graph has baked in changes (a layer's (scale, zero_point),
# its weights, its fuse_relu flag): api.Module keys its graph cache on it, so a stale graph is never replayed.
_GRAPH_EPOCH = [0]


def graph_epoch():
    return _GRAPH_EPOCH[0]


def _bump_epoch():
    _GRAPH_EPOCH[0] += 1
_F32_MAX_EXACT = 1 << 24


def _r16(v):
    return (int(v) + 15) // 16 * 16


def _act_pitch(c):
    """Channel pitch (bytes) of a u8 NHWC activation with c channels. The tcgen05 conv kernel
    consumes one tap's channel run in K blocks of 128/64/32 bytes; wider blocks mean 4x/2x
    fewer TMA requests per byte, so channels are padded up to the widest block that costs at
    most ~1/3 extra K (pad lanes hold the zero point, the matching weight lanes are zero)."""
    c = int(c)
    if c <= 16:
        return 16
    if c <= 32:
        return 32
    for blk in (128, 64):
        p = (c + blk - 1) // blk * blk
        if p * 3 <= c * 4 + 2:
            return p
    return (c + 31) // 32 * 32


_SLACK = 128


def _u8_empty(numel, device):
    """u8 activation buffer of exactly `numel` bytes with >= 128 readable bytes behind it in the same
    allocation: a row-mode convolution's last K block reads a few bytes past the last pixel (against
    zero weights), which must stay inside mapped memory."""
    return torch.empty(int(numel) + _SLACK, dtype=torch.uint8, device=device)[:int(numel)]


def _has_slack(t):
    st = t.untyped_storage()
    return st.nbytes() - (t.storage_offset() + t.numel()) * t.element_size() >= _SLACK


def _need_cuda():
    if not torch.cuda.is_available():
        raise I8ieError("no CUDA device visible: the i8ie B200 backend has no CPU fallback")
    return _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _log(msg):
    # the reference prints these through test_utils.h print() to stderr (layer.cc:30,38,42)
    print(msg, file=sys.stderr)


class _Storage:
    """Stands in for the py::capsule that owns a Tensor's buffer (tensor.h:28,46,52,60)."""
    __slots__ = ("t", "views")

    def __init__(self, t):
        self.t = t
        self.views = 0


class _SlotStorage:
    """Storage of a CUDA-graph input: the graph's first kernel reads the source ADDRESS from a
    device slot at run time (the "_indirect" entry points), so a replay works on whatever buffer
    the caller passed without a staging copy. Any consumer that needs the bytes at a fixed
    address gets them materialised into a static buffer by one in-graph copy kernel."""

    def __init__(self, slot, numel, device):
        self.slot, self.numel, self.device = slot, int(numel), device
        self.views = 0
        self._static = None

    @property
    def materialised(self):
        return self._static is not None

    @property
    def t(self):
        if self._static is None:
            L = _need_cuda()
            self._static = torch.empty(self.numel, dtype=torch.float32, device=self.device)
            if (self.numel * 4) % 16 == 0:
                check(L.i8ie_copy_indirect(self.slot.data_ptr(), self._static.data_ptr(), self.numel * 4, _stream()),
                      "copy_indirect")
            else:
                raise I8ieError("graph input of %d floats is not a multiple of 16 bytes" % self.numel)
        return self._static


class _ChunkedStorage:
    """Storage whose bytes are still arriving from pinned host memory: `tensor()` splits a large
    pinned batch into row chunks copied on a side stream, one event per chunk. `Module.__call__`
    consumes the chunks one by one (the forward of chunk k overlaps the copy of chunk k+1; images
    are independent, so the result does not depend on the split); any other access simply orders
    the current stream behind the last copy."""

    def __init__(self, t, chunks):
        self._t = t
        self.views = 0
        self.chunks = chunks          # [(row_begin, row_end, event)] or None once ordered

    @property
    def t(self):
        if self.chunks:
            torch.cuda.current_stream().wait_event(self.chunks[-1][2])   # one copy stream: the last event covers all
            self.chunks = None
        return self._t


_COPY_STREAM = None
H2D_CHUNK_MIN_BYTES = 8 << 20      # below this a single copy is already short
H2D_CHUNK_MIN_ROWS = 16


def _h2d_chunks(rows, nbytes):
    if nbytes < H2D_CHUNK_MIN_BYTES or os.environ.get("I8IE_NO_H2D_CHUNKS"):
        return 1
    for c in (8, 6, 5, 4, 3, 2):
        if rows % c == 0 and rows // c >= H2D_CHUNK_MIN_ROWS:
            return c
    return 1


def _tensor_from_pinned(h):
    """Pinned host tensor -> device, as row chunks on the copy stream (see _ChunkedStorage)."""
    global _COPY_STREAM
    shape = list(h.shape)
    flat = h.reshape(-1)
    c = _h2d_chunks(shape[0], flat.numel() * 4) if len(shape) >= 2 else 1
    if c == 1:
        return TensorF32(_Storage(flat.to("cuda", non_blocking=True)), shape)
    if _COPY_STREAM is None:
        _COPY_STREAM = torch.cuda.Stream()
    cur = torch.cuda.current_stream()
    dev = torch.empty(flat.numel(), dtype=torch.float32, device="cuda")
    dev.record_stream(_COPY_STREAM)
    _COPY_STREAM.wait_stream(cur)      # the block may be recycled from work still queued on this stream
    rows = shape[0] // c
    per = flat.numel() // shape[0]
    chunks = []
    with torch.cuda.stream(_COPY_STREAM):
        for k in range(c):
            a, b = k * rows, (k + 1) * rows
            dev[a * per:b * per].copy_(flat[a * per:b * per], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(_COPY_STREAM)
            chunks.append((a, b, ev))
    return TensorF32(_ChunkedStorage(dev, chunks), shape)


# ---- pageable host arrays (what an unmodified reference script passes): staged through pinned memory ----
_STAGE = {"bufs": [], "next": 0, "pool": None}
_STAGE_THREADS = max(1, min(8, (os.cpu_count() or 1) // 2))


def _staging(numel):
    """One of two pinned staging buffers (so the DMA of one batch can still be running while the next
    batch is being staged); waits until the buffer's previous DMA has finished."""
    st = _STAGE
    if len(st["bufs"]) < 2:
        st["bufs"].append({"t": None, "ev": None})
    b = st["bufs"][st["next"] % len(st["bufs"])]
    st["next"] += 1
    if b["ev"] is not None:
        b["ev"].synchronize()
    if b["t"] is None or b["t"].numel() < numel:
        b["t"] = torch.empty(numel, dtype=torch.float32).pin_memory()
    return b


def _parallel_copy(dst, src):
    """dst[:] = src for large contiguous float32 numpy views, split over a few threads (memcpy releases the GIL)."""
    n = src.size
    if _STAGE_THREADS == 1 or n < (1 << 20):
        np.copyto(dst, src)
        return
    if _STAGE["pool"] is None:
        from concurrent.futures import ThreadPoolExecutor
        _STAGE["pool"] = ThreadPoolExecutor(max_workers=_STAGE_THREADS)
    step = (n + _STAGE_THREADS - 1) // _STAGE_THREADS
    futs = [_STAGE["pool"].submit(np.copyto, dst[i:i + step], src[i:i + step]) for i in range(0, n, step)]
    for f in futs:
        f.result()


def _tensor_from_pageable(a):
    """Contiguous float32 numpy array -> device. The bytes are copied SYNCHRONOUSLY into pinned staging
    memory (when this returns the caller may modify the array, like tensor.h:40-47), chunk by chunk,
    and each chunk's host->device DMA is queued as soon as it is staged; Module.__call__ consumes the
    chunks as they land, exactly as for a pinned source (see _ChunkedStorage)."""
    global _COPY_STREAM
    shape = list(a.shape)
    flat = a.reshape(-1)
    c = _h2d_chunks(shape[0], flat.size * 4) if len(shape) >= 2 else 1
    if _COPY_STREAM is None:
        _COPY_STREAM = torch.cuda.Stream()
    stage = _staging(flat.size)
    st_t = stage["t"][:flat.size]
    st_np = st_t.numpy()
    cur = torch.cuda.current_stream()
    dev = torch.empty(flat.size, dtype=torch.float32, device="cuda")
    dev.record_stream(_COPY_STREAM)
    _COPY_STREAM.wait_stream(cur)
    rows = shape[0] // c if c > 1 else shape[0] if shape else 1
    per = flat.size // shape[0] if (shape and shape[0]) else flat.size
    chunks = []
    for k in range(c):
        lo, hi = (k * rows * per, (k + 1) * rows * per) if c > 1 else (0, flat.size)
        _parallel_copy(st_np[lo:hi], flat[lo:hi])
        with torch.cuda.stream(_COPY_STREAM):
            dev[lo:hi].copy_(st_t[lo:hi], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(_COPY_STREAM)
        chunks.append((k * rows, (k + 1) * rows, ev) if c > 1 else (0, shape[0] if shape else 1, ev))
    stage["ev"] = chunks[-1][2]
    if c == 1:
        cur.wait_event(chunks[-1][2])
        return TensorF32(_Storage(dev), shape)
    return TensorF32(_ChunkedStorage(dev, chunks), shape)


def pending_chunks(x):
    """[(sub-batch tensor, copy-done event)] of a float tensor still arriving in chunks, else None."""
    st = x._st
    if not isinstance(st, _ChunkedStorage) or not st.chunks or len(x._shape) < 2 or x._shape[0] != st.chunks[-1][1]:
        return None
    per = st._t.numel() // x._shape[0]
    return [(TensorF32(_Storage(st._t[a * per:b * per]), [b - a] + x._shape[1:]), ev) for a, b, ev in st.chunks]


def host_buffer_free(x):
    """Blocks until the asynchronous host->device copy behind a tensor made from a pinned CPU
    tensor has finished, i.e. until the caller may overwrite that pinned buffer."""
    st = x._storage
    if isinstance(st, _ChunkedStorage):
        if st.chunks:
            st.chunks[-1][2].synchronize()
        else:
            torch.cuda.current_stream().synchronize()
    elif st is not None:
        torch.cuda.current_stream().synchronize()


def wait_event(ev):
    torch.cuda.current_stream().wait_event(ev)


def chunks_consumed(x):
    x._st.chunks = None


def concat_rows(parts):
    """Stacks per-chunk results (same trailing shape) along the batch dimension."""
    shape = [sum(p._shape[0] for p in parts)] + parts[0]._shape[1:]
    return TensorF32(_Storage(torch.cat([p.buf for p in parts])), shape)


def _src_ptrs(x):
    """(direct pointer or None, slot pointer or None, device) of a float tensor's bytes."""
    st = x._st
    if isinstance(st, _SlotStorage) and not st.materialised:
        return None, st.slot.data_ptr(), st.device
    return st.t.data_ptr(), None, st.t.device


class _TensorBase:
    torch_dtype = None
    np_dtype = None

    def __init__(self, storage, shape, layout="dense", geom=None, scale=1.0, zp=0):
        self._storage = storage
        if storage is not None:
            storage.views += 1
        self._shape = [int(s) for s in shape]
        self._layout = layout          # 'dense' | 'nhwc'
        self._geom = geom              # (n, c, h, w, cp) when layout == 'nhwc'
        self._scale = np.float32(scale)  # tensor.h:153
        self._zp = int(zp)               # tensor.h:154

    def __del__(self):
        try:
            if self._storage is not None:
                self._storage.views -= 1
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    @property
    def _st(self):
        return self._storage

    # -- the methods src/pybind11.cc:12-30 binds --------------------------------------
    def zero_point(self):
        return self._zp

    def scale(self):
        return float(self._scale)

    def ref_count(self):
        return self._st.views

    def numpy(self):
        out = self._dense_buf().view(self._shape).cpu().numpy()
        _lib.check_tc_error()     # the D2H copy synchronised: a pipeline fault upstream raises here
        return out

    def sum(self):
        # pybind11.cc:18-25: sequential fp32 running sum
        a = self.numpy().astype(np.float32).ravel()
        return float(np.cumsum(a, dtype=np.float32)[-1]) if a.size else 0.0

    def reshape(self, shape):
        # Tensor::reshape, tensor.h:106-133 (one negative entry inferred; zeros rejected)
        shape = [int(s) for s in shape]
        size = int(np.prod(self._shape)) if self._shape else 1
        midx, sz = -1, 1
        for i, s in enumerate(shape):
            if s < 0:
                if midx != -1:
                    raise RuntimeError("std::exception")
                midx = i
            elif s == 0:
                raise RuntimeError("std::exception")
            else:
                sz *= s
        if midx >= 0:
            if size % sz != 0:
                raise RuntimeError("std::exception")
            shape[midx] = size // sz
            sz *= shape[midx]
        if sz != size:
            raise RuntimeError("std::exception")
        return self._view(shape)

    # -- internals ---------------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self._shape)

    @property
    def buf(self):
        return self._st.t

    def _dense_buf(self):
        """Flat torch tensor holding the logical (row-major NCHW) element order."""
        return self._st.t

    def _view(self, shape):
        # tensor_view (tensor.h:94-104): shares the capsule, carries scale / zero_point
        return type(self)(self._st, shape, "dense", None, self._scale, self._zp)


class TensorF32(_TensorBase):
    """Tensor<float> (`6TensorIfE`)."""
    torch_dtype = torch.float32
    np_dtype = np.float32


class TensorS8(_TensorBase):
    """Tensor<s8_t> (`6TensorIcE`) — quantised weights."""
    torch_dtype = torch.int8
    np_dtype = np.int8


class TensorU8(_TensorBase):
    """Tensor<u8_t> (`6TensorIhE`) — quantised activations, NHWC / padded rows on device."""
    torch_dtype = torch.uint8
    np_dtype = np.uint8

    def __init__(self, storage, shape, layout="dense", geom=None, scale=1.0, zp=0, deferred=None):
        super().__init__(storage, shape, layout, geom, scale, zp)
        self._deferred = deferred      # a _Deferred* op that has not been launched yet

    @property
    def _st(self):
        # Deferred tensors launch their kernel the first time the bytes are needed. Until then a
        # following relu / flatten / first-layer conv can still fold itself into that launch.
        if self._storage is None:
            self._resolve(*self._deferred.launch())
        return self._storage

    def _resolve(self, st, layout, geom):
        self._storage = st
        st.views += 1
        self._layout, self._geom = layout, geom
        self._deferred = None

    def _pending(self, kind=None):
        d = self._deferred if self._storage is None else None
        return d if (d is not None and (kind is None or d.kind == kind)) else None

    def _dense_buf(self):
        self._st  # noqa: B018 - materialise
        if self._layout == "dense":
            return self._st.t
        n, c, h, w, cp = self._geom
        if cp == c and h * w == 1:
            return self._st.t
        L = _need_cuda()
        out = torch.empty(n * c * h * w, dtype=torch.uint8, device=self._st.t.device)
        check(L.i8ie_u8_nhwc_to_nchw(self._st.t.data_ptr(), out.data_ptr(), n, c, h, w, cp, _stream()),
              "u8_nhwc_to_nchw")
        return out

    def _view(self, shape):
        pool = self._pending("pool")
        if pool is not None and not pool.out_nchw:
            # flatten of a not-yet-launched pool: let the pool kernel write the logical (NCHW)
            # order directly instead of NHWC + a re-layout pass (tensor.h:106-133 semantics)
            return TensorU8(None, shape, "dense", None, self._scale, self._zp, deferred=pool.as_nchw())
        self._st  # noqa: B018 - materialise
        if self._layout == "dense":
            return TensorU8(self._st, shape, "dense", None, self._scale, self._zp)
        n, c, h, w, cp = self._geom
        if cp == c and h * w == 1:  # rows without padding: physical == logical
            return TensorU8(self._st, shape, "dense", None, self._scale, self._zp)
        # physical NHWC differs from the logical order: materialise the logical order once
        return TensorU8(_Storage(self._dense_buf()), shape, "dense", None, self._scale, self._zp)

    def _as_padded_nhwc(self, n, c, h, w, cpx, pad):
        """This tensor as a PHYSICALLY PADDED NHWC buffer [n][h + 2*pad][w + 2*pad][cpx] whose border holds
        the zero point (the operand layout of a row-mode convolution plan). A not-yet-launched max-pool
        writes it directly; anything else is pad-copied once (cached: tensors are immutable)."""
        cache = self.__dict__.setdefault("_padded", {})
        hit = cache.get((cpx, pad))
        if hit is not None:
            return hit
        L = _need_cuda()
        pool = self._pending("pool")
        if pool is not None and not pool.out_nchw:
            out = pool.launch_padded(cpx, pad, self._zp)
        else:
            buf, cp = self._as_nhwc(n, c, h, w)
            if pad == 0 and cp == cpx and _has_slack(buf):
                cache[(cpx, pad)] = buf      # no border to add: the NHWC tensor is already the operand
                return buf
            out = torch.empty(n * (h + 2 * pad) * (w + 2 * pad) * cpx + 128, dtype=torch.uint8, device=buf.device)
            check(L.i8ie_maxpool_u8_nhwc_padded(buf.data_ptr(), out.data_ptr(), n, h, w, c, cp, 1, 1, cpx, pad,
                                                self._zp, _stream()), "u8_nhwc pad-copy")
        cache[(cpx, pad)] = out
        return out

    def _as_nhwc(self, n, c, h, w, want_cp=None):
        """(buffer, cp) of this tensor as NHWC. An NHWC tensor of the right shape is used as it
        is, whatever its pitch (unless want_cp insists); anything else is converted once."""
        self._st  # noqa: B018 - materialise
        if self._layout == "nhwc" and self._geom[:4] == (n, c, h, w) and (want_cp is None or self._geom[4] == want_cp):
            return self._st.t, self._geom[4]
        cp = want_cp or _act_pitch(c)
        if self._layout == "dense" and h * w == 1 and cp == c:
            return self._st.t, cp
        L = _need_cuda()
        src = self._dense_buf()
        out = _u8_empty(n * h * w * cp, src.device)
        check(L.i8ie_u8_nchw_to_nhwc(src.data_ptr(), out.data_ptr(), n, c, h, w, cp, self._zp, _stream()),
              "u8_nchw_to_nhwc")
        return out, cp


class _DeferredLayer:
    """conv / fc launch that can still absorb a following relu<u8> (functional.cc:15-26 is
    max(y, zero_point): identical whether applied by a second kernel or in the epilogue)."""
    kind = "layer"

    def __init__(self, layer, x, relu=False):
        self.layer, self.x, self.relu = layer, x, relu

    def with_relu(self):
        return _DeferredLayer(self.layer, self.x, True)

    def launch(self, out_cp=None):
        kw = {"out_cp": out_cp} if out_cp is not None else {}
        y = self.layer._forward_u8(self.x, relu=self.relu or self.layer.fuse_relu, **kw)
        return y._st, y._layout, y._geom


def _launch_conv_narrow(x):
    """x is about to be read by a max-pool that chooses its own OUTPUT pitch (padded layout of a row-mode
    convolution, or the dense NCHW order of a flatten). If x is a convolution that has not been launched yet,
    let it store its real channels only (pitch round_up(C, 16)) instead of the K-block-friendly pitch a following
    convolution would want: AlexNet conv1 writes 96 instead of 128 bytes per pixel, and pool1 reads as much."""
    pend = x._pending("layer")
    if pend is not None and isinstance(pend.layer, Conv2d) and len(x._shape) == 4:
        c = x._shape[1]
        if _r16(c) < _act_pitch(c) and not os.environ.get("I8IE_NO_NARROW_PITCH"):
            x._resolve(*pend.launch(out_cp=_r16(c)))


class _DeferredPool:
    kind = "pool"

    def __init__(self, x, k, s, out_nchw=False):
        self.x, self.k, self.s, self.out_nchw = x, k, s, out_nchw

    def as_nchw(self):
        return _DeferredPool(self.x, self.k, self.s, True)

    def launch(self):
        L = _need_cuda()
        x, k, s = self.x, self.k, self.s
        n, c, h, w = x._shape
        oh, ow = (h - k) // s + 1, (w - k) // s + 1
        if self.out_nchw:
            _launch_conv_narrow(x)
        buf, cp = x._as_nhwc(n, c, h, w)
        out = _u8_empty(n * oh * ow * (c if self.out_nchw else cp), buf.device)
        check(L.i8ie_maxpool_u8_nhwc(buf.data_ptr(), out.data_ptr(), n, h, w, c, cp, k, s,
                                     1 if self.out_nchw else 0, _stream()), "maxpool_u8_nhwc")
        if self.out_nchw:
            return _Storage(out), "dense", None
        return _Storage(out), "nhwc", (n, c, oh, ow, cp)

    def launch_padded(self, out_cp, out_pad, zp):
        """The pool written straight into the physically padded layout a row-mode convolution reads
        (border = zero point, channel pitch out_cp); + 128 readable bytes behind the tensor."""
        L = _need_cuda()
        x, k, s = self.x, self.k, self.s
        n, c, h, w = x._shape
        oh, ow = (h - k) // s + 1, (w - k) // s + 1
        _launch_conv_narrow(x)
        buf, cp = x._as_nhwc(n, c, h, w)
        out = torch.empty(n * (oh + 2 * out_pad) * (ow + 2 * out_pad) * out_cp + 128, dtype=torch.uint8,
                          device=buf.device)
        check(L.i8ie_maxpool_u8_nhwc_padded(buf.data_ptr(), out.data_ptr(), n, h, w, c, cp, k, s, out_cp, out_pad,
                                            int(zp), _stream()), "maxpool_u8_nhwc_padded")
        return out


class _DeferredQuant:
    """Input quantise (module.py:20) that a first-layer stem convolution can fuse."""
    kind = "quant"

    def __init__(self, src, scale, zp, geom):
        self.src, self.scale, self.zp, self.geom = src, scale, zp, geom

    def launch(self):
        L = _need_cuda()
        n, c, h, w, cp = self.geom
        ptr, slot, dev = _src_ptrs(self.src)
        out = _u8_empty(n * h * w * cp, dev)
        if slot is not None:
            check(L.i8ie_quantize_nchw_f32_nhwc_u8_indirect(slot, out.data_ptr(), n, c, h, w, cp, self.scale,
                                                            self.zp, _stream()), "quantize_nchw_f32_nhwc_u8_indirect")
        else:
            check(L.i8ie_quantize_nchw_f32_nhwc_u8(ptr, out.data_ptr(), n, c, h, w, cp, self.scale, self.zp,
                                                   _stream()), "quantize_nchw_f32_nhwc_u8")
        return _Storage(out), "nhwc", self.geom


def _new_u8_nhwc(buf, n, c, h, w, cp, scale, zp, two_d=False):
    shape = [n, c] if two_d else [n, c, h, w]
    return TensorU8(_Storage(buf), shape, "nhwc", (n, c, h, w, cp), scale, zp)


# ---- module-level functions (src/pybind11.cc:38-48, functional.cc:66-82) -----------------

def tensor(ndarray):
    """_CXX_i8ie.tensor: py::array_t<float> force-casts anything array-like to f32 and copies
    it in (tensor.h:40-47).

    numpy arrays and pageable CPU tensors are copied synchronously, like the reference. A PINNED
    CPU torch tensor is the fast path: its host->device copy is asynchronous (row chunks on a side
    stream, see _ChunkedStorage) and no host-side copy is made, so the caller must not overwrite
    the pinned buffer until the copy has finished — call `host_buffer_free(t)` (blocks until the
    last chunk has left the host buffer) before reusing it, or pass a numpy array instead."""
    _need_cuda()
    if isinstance(ndarray, torch.Tensor) and ndarray.device.type == "cpu":
        # same force-cast; a pinned f32 tensor goes to the device without a staging copy
        h = ndarray.detach().to(torch.float32).contiguous()
        if h.is_pinned():
            return _tensor_from_pinned(h)
        return TensorF32(_Storage(h.reshape(-1).to("cuda", non_blocking=True)), list(h.shape))
    a = np.ascontiguousarray(np.asarray(ndarray), dtype=np.float32)
    if a.size * 4 >= H2D_CHUNK_MIN_BYTES and not os.environ.get("I8IE_NO_H2D_CHUNKS"):
        return _tensor_from_pageable(a)
    t = torch.from_numpy(a.reshape(-1).copy()).to("cuda", non_blocking=False)
    return TensorF32(_Storage(t), list(a.shape))


def tensor_from_torch(t):
    """Extension (not in the reference): adopt a CUDA/CPU torch tensor without a host round trip."""
    _need_cuda()
    t = t.detach().to(device="cuda", dtype=torch.float32).contiguous()
    return TensorF32(_Storage(t.reshape(-1)), list(t.shape))


def quantize(x, scale, zero_point):
    """quantize(Tensor<float>&, scale, zp) — quantize_utils.cc:44-52 (unclamped, truncating)."""
    L = _need_cuda()
    if not isinstance(x, TensorF32):
        raise TypeError("quantize(): incompatible function arguments (expected a float tensor)")
    zp = int(zero_point)
    if not 0 <= zp <= 255:
        raise TypeError("quantize(): zero_point must fit unsigned char")
    scale = float(np.float32(scale))
    shp = x._shape
    if len(shp) in (2, 4):
        n, c = shp[0], shp[1]
        h, w = (shp[2], shp[3]) if len(shp) == 4 else (1, 1)
        geom = (n, c, h, w, _act_pitch(c))
        return TensorU8(None, shp, "nhwc", geom, scale, zp, deferred=_DeferredQuant(x, scale, zp, geom))
    ptr, slot, dev = _src_ptrs(x)
    numel = int(np.prod(shp)) if shp else 1
    out = torch.empty(numel, dtype=torch.uint8, device=dev)
    if slot is not None:
        check(L.i8ie_quantize_f32_u8_indirect(slot, out.data_ptr(), numel, scale, zp, _stream()),
              "quantize_f32_u8_indirect")
    else:
        check(L.i8ie_quantize_f32_u8(ptr, out.data_ptr(), numel, scale, zp, _stream()), "quantize_f32_u8")
    return TensorU8(_Storage(out), shp, "dense", None, scale, zp)


def dequantize(x):
    """dequantize(Tensor<u8>&) — quantize_utils.cc:54-58 -> :38-42."""
    L = _need_cuda()
    if not isinstance(x, TensorU8):
        raise TypeError("dequantize(): incompatible function arguments (expected a u8 tensor)")
    pend = x._pending("layer")
    if pend is not None and isinstance(pend.layer, Linear):
        # the model's last fc has not been launched yet: one entry point computes the u8 result and its
        # dequantised copy (classifier heads: in the same kernel)
        y, out = pend.layer._forward_u8(pend.x, relu=pend.relu or pend.layer.fuse_relu, deq=True)
        x._resolve(y._st, y._layout, y._geom)
        return TensorF32(_Storage(out), x._shape)
    numel = int(np.prod(x._shape)) if x._shape else 1
    out = torch.empty(numel, dtype=torch.float32, device=x.buf.device)
    if x._layout == "nhwc" and x._geom[2] * x._geom[3] == 1:
        n, c, _, _, cp = x._geom
        check(L.i8ie_dequantize_rows_u8_f32(x.buf.data_ptr(), out.data_ptr(), n, c, cp, x.scale(), x._zp,
                                            _stream()), "dequantize_rows_u8_f32")
    else:
        src = x._dense_buf()
        check(L.i8ie_dequantize_u8_f32(src.data_ptr(), out.data_ptr(), numel, x.scale(), x._zp, _stream()),
              "dequantize_u8_f32")
    return TensorF32(_Storage(out), x._shape)


def relu(x):
    """relu<T> — functional.cc:5-26. u8: max(x, zero_point); fp32: max(x, 0)."""
    L = _need_cuda()
    if isinstance(x, TensorU8):
        pend = x._pending("layer")
        if pend is not None and not pend.relu:
            return TensorU8(None, x._shape, x._layout, x._geom, x._scale, x._zp, deferred=pend.with_relu())
        out = torch.empty_like(x.buf)
        check(L.i8ie_relu_u8(x.buf.data_ptr(), out.data_ptr(), x.buf.numel(), x._zp, _stream()), "relu_u8")
        return TensorU8(_Storage(out), x._shape, x._layout, x._geom, x._scale, x._zp)
    if isinstance(x, TensorF32):
        out = torch.empty_like(x.buf)
        check(L.i8ie_relu_f32(x.buf.data_ptr(), out.data_ptr(), x.buf.numel(), _stream()), "relu_f32")
        return TensorF32(_Storage(out), x._shape)
    raise TypeError("relu(): unsupported tensor type")


def max_pool2d(x, kernel_size, strides):
    """max_pool2d<T> — functional.cc:36-64 (no padding, floor output size)."""
    L = _need_cuda()
    k, s = int(kernel_size), int(strides)
    if len(x._shape) != 4:
        raise RuntimeError("max_pool2d expects a 4-d tensor")
    n, c, h, w = x._shape
    if k < 1 or s < 1 or h < k or w < k:
        raise RuntimeError("max_pool2d: window larger than input")
    oh, ow = (h - k) // s + 1, (w - k) // s + 1
    if isinstance(x, TensorU8):
        cp = x._geom[4] if x._layout == "nhwc" else _act_pitch(c)
        return TensorU8(None, [n, c, oh, ow], "nhwc", (n, c, oh, ow, cp), x._scale, x._zp,
                        deferred=_DeferredPool(x, k, s))
    if isinstance(x, TensorF32):
        y = torch.empty(n * c * oh * ow, dtype=torch.float32, device=x.buf.device)
        check(L.i8ie_maxpool_f32_nchw(x.buf.data_ptr(), y.data_ptr(), n, c, h, w, k, s, _stream()), "maxpool_f32_nchw")
        return TensorF32(_Storage(y), [n, c, oh, ow])
    raise TypeError("max_pool2d(): unsupported tensor type")


# ---- calibrator (include/calibrator.h, src/calibrator.cc) ---------------------------------

class Calibrator:
    """Activation-range collector. The reference keeps a 1000-slot sample with random
    replacement from a random_device-seeded mt19937 (calibrator.cc:6-23) and takes min/max of the
    sorted slots (get_range(1), :24-37).

    mode "minmax" (default): the range is the true min/max of everything seen — a warp-shuffle /
    block reduction on the device (i8ie_minmax_f32), deterministic — which is what the reference
    computes whenever <=1000 values were seen; for exactly that regime the first 1000 values are
    also kept so the zero-filled-tail quirk is reproduced. Beyond 1000 values the
    reference's sample is ~1000 of the LAST few thousand values in memory order (each later value
    overwrites a uniformly chosen slot with probability 1000/2001), so its ranges are narrower and
    differ from run to run; this mode's scales are systematically wider than that.

    mode "reference" (opt-in: `Calibrator.default_mode = "reference"` or I8IE_CALIBRATOR=reference):
    a seeded emulation of that sampling process, exact in distribution per slot — the element a slot
    holds at the end is the last one that hit it, i.e. `G` positions before the end of the stream
    with G ~ Geometric(1/2001) — so a model calibrated here gets ranges distributed like the
    reference's (and reproducible for a fixed `Calibrator.seed`)."""

    default_mode = os.environ.get("I8IE_CALIBRATOR", "minmax")
    seed = 0
    _instances = 0

    def __init__(self, mode=None):
        self.mode = mode or Calibrator.default_mode
        if self.mode not in ("minmax", "reference"):
            raise ValueError(f"unknown calibrator mode {self.mode!r}")
        self.out_cnt = 0
        self.head = []          # first <=1000 samples, host copies
        self.mn = None
        self.mx = None
        self._ws = None
        self._out = None
        self.slots = None       # "reference" mode: the 1000-slot sample (host)
        self._rng = np.random.default_rng([Calibrator.seed, Calibrator._instances])
        Calibrator._instances += 1

    def new_range(self, device):
        """Device {min, max} pair a forward kernel folds its outputs into (fused calibrator pass)."""
        return torch.tensor([np.finfo(np.float32).max, -np.finfo(np.float32).max], dtype=torch.float32, device=device)

    @staticmethod
    def replacement_indices(n, first, rng):
        """calibrator.cc:13-21 for the elements [first, n) of one sample() call, in distribution: for
        every slot the index of the LAST element that hit it (uniform slot pick out of 2 * 1000 + 1, so
        a given slot is hit with probability 1/2001 per element), or -1 if none did.
        Returns int64[1000]."""
        g = rng.geometric(1.0 / (2 * NUM_SAMPLES + 1), size=NUM_SAMPLES) - 1     # misses counted from the end
        idx = (n - 1) - g
        return np.where(idx >= first, idx, -1).astype(np.int64)

    def _sample_slots(self, t, take):
        """"reference" mode bookkeeping for one sample() call; `take` leading elements went to free slots."""
        n = t.numel()
        if self.slots is None:
            self.slots = np.zeros(NUM_SAMPLES, np.float32)      # value-initialised (layer.cc:33)
        flat = t.reshape(-1)
        if take:
            self.slots[self.out_cnt:self.out_cnt + take] = flat[:take].cpu().numpy()
        if n > take:
            idx = self.replacement_indices(n, take, self._rng)
            hit = np.nonzero(idx >= 0)[0]
            if hit.size:
                src = torch.from_numpy(idx[hit]).to(flat.device)
                self.slots[hit] = flat.index_select(0, src).cpu().numpy()

    def _take(self, t):
        n = t.numel()
        take = min(NUM_SAMPLES - self.out_cnt, n) if self.out_cnt < NUM_SAMPLES else 0
        if self.mode == "reference":
            self._sample_slots(t, take)
        elif take:
            self.head.append(t.reshape(-1)[:take].cpu().numpy())

    def sample_range(self, t, mm):
        """sample() for a layer output whose {min, max} the producing kernel already reduced."""
        n = t.numel()
        if n == 0:
            return
        self._take(t)
        mn, mx = (float(v) for v in mm.cpu().numpy())
        self.mn = mn if self.mn is None else min(self.mn, mn)
        self.mx = mx if self.mx is None else max(self.mx, mx)
        self.out_cnt += n

    def sample(self, t):
        L = _need_cuda()
        n = t.numel()
        if n == 0:
            return
        self._take(t)
        if self._ws is None:
            self._ws = torch.zeros(int(L.i8ie_minmax_workspace_bytes()), dtype=torch.uint8, device=t.device)
            self._out = torch.empty(2, dtype=torch.float32, device=t.device)
        check(L.i8ie_minmax_f32(t.data_ptr(), n, self._out.data_ptr(), self._ws.data_ptr(), _stream()),
              "minmax_f32")
        mn, mx = (float(v) for v in self._out.cpu().numpy())
        self.mn = mn if self.mn is None else min(self.mn, mn)
        self.mx = mx if self.mx is None else max(self.mx, mx)
        self.out_cnt += n

    def get_range(self):
        L = _lib.load()
        if self.out_cnt == 0:
            raise RuntimeError("convert() after prepare() needs at least one calibration forward pass")
        cnt = min(self.out_cnt, NUM_SAMPLES)
        if self.mode == "reference":
            buf = np.sort(self.slots)                       # calibrator.cc:25
            mn, mx = buf[0], buf[cnt - 1]                   # :26-27 with quantile 1
        elif self.out_cnt <= NUM_SAMPLES:
            buf = np.zeros(NUM_SAMPLES, np.float32)        # value-initialised slots (layer.cc:33)
            head = np.concatenate(self.head)
            buf[:head.size] = head
            buf.sort()                                      # calibrator.cc:25
            mn, mx = buf[0], buf[cnt - 1]                   # :26-27 with quantile 1
        else:
            mn, mx = self.mn, self.mx
        sc, zp = C.c_float(), C.c_uint8()
        check(L.i8ie_range_from_minmax_host(float(mn), float(mx), C.byref(sc), C.byref(zp)))
        return np.float32(sc.value), int(zp.value)


# ---- layers (include/layer.h, src/layer.cc, conv2d.cc, fully_connected.cc) ----------------

class _BaseLayer:
    # F4 extension (opt-in, not in the reference): set `layer.per_channel = True` before convert() to
    # quantise every output channel's weights (and its bias entry) with its own scale — layer.cc:6-26
    # applied per channel — and requantise with that channel's scale. The default (False) is the
    # reference's single per-tensor scale and stays bit-identical to it.
    per_channel = False

    def __init__(self, weight_shape, out_channel):
        _need_cuda()
        # BaseLayer ctor allocates uninitialised weight/bias (layer.h:9-12)
        self._w = np.zeros(weight_shape, np.float32)
        self._b = np.zeros((out_channel,), np.float32)
        self._w_dev = None
        self._b_dev = None
        self._cal = None
        self._is_preparing = False
        self._is_quantized = False
        self._scale = np.float32(1.0)   # layer.h:46
        self._zp = 0                    # layer.h:47
        self._w_scale = None
        self._qw_dev = None             # s8, reference row order [n, K]
        self._qb_dev = None
        self._w_packed = None           # s8, kernel layout
        self._oc_cache = {}
        self._plans = {}
        self._fuse_relu = False         # extension: fold a following relu<u8> into the epilogue

    @property
    def fuse_relu(self):
        return self._fuse_relu

    @fuse_relu.setter
    def fuse_relu(self, v):
        if bool(v) != self._fuse_relu:
            _bump_epoch()
        self._fuse_relu = bool(v)

    def load_weight(self, w):
        if self._w is None:
            raise RuntimeError("std::exception")      # layer.h:16-18
        self._w = np.ascontiguousarray(np.asarray(w), dtype=np.float32)
        self._w_dev = None
        _bump_epoch()

    def load_bias(self, b):
        if self._w is None:
            raise RuntimeError("std::exception")      # layer.h:22-24
        self._b = np.ascontiguousarray(np.asarray(b), dtype=np.float32)
        self._b_dev = None
        _bump_epoch()

    def prepare(self):
        if self._is_quantized:
            _log("already quantized")                 # layer.cc:29-32
            return
        self._cal = Calibrator()
        self._is_preparing = True

    def set_qparams(self, scale, zero_point):
        """Extension hook (not in the reference): inject a calibrated (scale, zero_point),
        e.g. the ones a reference run produced, before convert(). The reference's own
        calibrator is randomised beyond 1000 samples, so parity runs pin ranges this way."""
        self._scale = np.float32(scale)
        self._zp = int(zero_point)
        self._cal = None
        self._is_preparing = False
        self._injected = True
        self._oc_cache = {}
        _bump_epoch()

    def convert(self):
        L = _need_cuda()
        if self._is_quantized:
            _log("already quantized")                 # layer.cc:37-40
            return
        if not self._is_preparing:
            if not getattr(self, "_injected", False):
                _log("No prepared, use default config")   # layer.cc:41-42
        else:
            self._scale, self._zp = self._cal.get_range()  # layer.cc:44
            self._cal = None
        qw = np.empty(self._w.shape, np.int8)
        qb = np.empty(self._b.shape, np.int8)
        sc = C.c_float()
        self._w_scales = None
        self._w_scales_dev = None
        if self.per_channel:
            n = self._w.shape[0]
            per = self._w.size // n
            scales = np.empty(n, np.float32)
            w2, qw2 = self._w.reshape(n, per), qw.reshape(n, per)
            for j in range(n):
                check(L.i8ie_quantize_weight_host(w2[j].ctypes.data, per, self._b[j:j + 1].ctypes.data, 1,
                                                  qw2[j].ctypes.data, qb[j:j + 1].ctypes.data, C.byref(sc)),
                      "quantize_weight_host")
                scales[j] = sc.value
            if not np.all(np.isfinite(scales)) or scales.min() <= 0:
                raise I8ieError("per_channel: a channel's weights and bias are all equal (zero scale)")
            self._w_scales = scales
            self._w_scales_dev = torch.from_numpy(scales).cuda()
            sc.value = float(scales.max())
        else:
            check(L.i8ie_quantize_weight_host(self._w.ctypes.data, self._w.size, self._b.ctypes.data, self._b.size,
                                              qw.ctypes.data, qb.ctypes.data, C.byref(sc)), "quantize_weight_host")
        self._w_scale = np.float32(sc.value)
        self._qw_dev = torch.from_numpy(qw.reshape(qw.shape[0], -1)).cuda()
        self._qb_dev = torch.from_numpy(qb).cuda()
        self._qw_shape = qw.shape
        self._pack()
        self._is_preparing = False
        self._is_quantized = True
        _bump_epoch()
        self._w = None          # layer.cc:52-53 frees the fp32 parameters
        self._b = None
        self._w_dev = None
        self._b_dev = None

    # extension accessors used by parity tests
    def q_weight(self):
        return TensorS8(_Storage(self._qw_dev.reshape(-1)), list(self._qw_shape), scale=self._w_scale)

    def q_bias(self):
        return TensorS8(_Storage(self._qb_dev), [self._qb_dev.numel()], scale=self._w_scale)

    def weight_scales(self):
        """Per-output-channel weight scales (numpy f32 [n]) of a layer converted with per_channel = True, else None."""
        return None if getattr(self, "_w_scales", None) is None else self._w_scales.copy()

    def _offsets(self, in_zp, in_scale, is_conv):
        key = (int(in_zp), float(in_scale))
        hit = self._oc_cache.get(key)
        if hit is None:
            L = _lib.load()
            n, k = self._qw_dev.shape
            oc = torch.empty(n, dtype=torch.int32, device="cuda")
            bf = torch.empty(n, dtype=torch.float32, device="cuda")
            check(L.i8ie_zp_offsets(self._qw_dev.data_ptr(), self._qb_dev.data_ptr(), n, k, int(in_zp),
                                    float(in_scale), int(is_conv), oc.data_ptr(), bf.data_ptr(), _stream()),
                  "zp_offsets")
            hit = (oc, bf)
            self._oc_cache[key] = hit
        return hit

    def _fp32_params(self):
        if self._w_dev is None:
            self._w_dev = torch.from_numpy(self._w).cuda()
        if self._b_dev is None:
            self._b_dev = torch.from_numpy(self._b).cuda()
        return self._w_dev, self._b_dev

    def __call__(self, x):
        if isinstance(x, TensorF32):
            if self._w is None:
                raise RuntimeError("layer was converted: fp32 parameters are released (layer.cc:52-53)")
            return self._forward_f32(x)
        if isinstance(x, TensorU8):
            if not self._is_quantized:
                raise RuntimeError("layer is not converted: call prepare()/convert() before a u8 forward")
            shape, geom = self._out_meta(x)     # validates the input shape now, launches later
            return TensorU8(None, shape, "nhwc", geom, self._scale, self._zp, deferred=_DeferredLayer(self, x))
        raise TypeError("__call__(): incompatible function arguments")


class Linear(_BaseLayer):
    """Linear — include/fully_connected.h, src/fully_connected.cc."""

    def __init__(self, in_channel, out_channel):
        super().__init__((int(out_channel), int(in_channel)), int(out_channel))

    def _pack(self):
        n, k = self._qw_dev.shape
        self._ldw = _r16(k)
        self._n_pad = _r16(n)
        wp = torch.zeros(self._n_pad, self._ldw, dtype=torch.int8, device="cuda")
        wp[:n, :k] = self._qw_dev
        self._drop_tiled()
        self._w_packed = wp
        # small-batch fc streams its weights from HBM: keep a second copy in contiguous, pre-swizzled
        # 16 KB blocks next to the K-major one (include/i8ie_sm100.h, i8ie_fc_weight_tiled_attach)
        L = _lib.load()
        nbytes = int(L.i8ie_fc_weight_tiled_bytes(self._n_pad, self._ldw))
        if nbytes > 0 and not os.environ.get("I8IE_NO_FC_TILED"):
            wt = torch.empty(nbytes, dtype=torch.int8, device="cuda")
            check(L.i8ie_fc_weight_tiled_attach(wp.data_ptr(), self._n_pad, self._ldw, wt.data_ptr(), _stream()),
                  "fc_weight_tiled_attach")
            self._w_tiled = wt

    def _drop_tiled(self):
        if getattr(self, "_w_tiled", None) is not None and self._w_packed is not None:
            try:
                _lib.load().i8ie_fc_weight_tiled_detach(self._w_packed.data_ptr())
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
        self._w_tiled = None

    def __del__(self):
        self._drop_tiled()

    def _forward_f32(self, x):
        # Linear::forward_prop(Tensor<float>&&), fully_connected.cc:5-21: fp32 FMA GEMM
        # with the bias add and the calibrator's min/max (fully_connected.cc:17-19) in the epilogue
        L = _lib.load()
        w, b = self._fp32_params()
        if len(x._shape) != 2 or x._shape[1] != w.shape[1]:
            raise RuntimeError(f"Linear: input shape {tuple(x._shape)} does not match weight {tuple(w.shape)}")
        m, k = x._shape
        n = w.shape[0]
        y = torch.empty(m * n, dtype=torch.float32, device=w.device)
        mm = self._cal.new_range(w.device) if self._is_preparing else None
        check(L.i8ie_linear_f32(x.buf.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), m, n, k,
                                mm.data_ptr() if mm is not None else None, _stream()), "linear_f32")
        if mm is not None:
            self._cal.sample_range(y, mm)
        return TensorF32(_Storage(y), [m, n])

    def _out_meta(self, x):
        if len(x._shape) != 2:
            raise RuntimeError("Linear expects a 2-d tensor")
        m, k = x._shape
        n, kw = self._qw_dev.shape
        if k != kw:
            raise RuntimeError(f"Linear: input has {k} features, weight expects {kw}")
        return [m, n], (m, n, 1, 1, _r16(n))

    def _forward_u8(self, x, acc_out=None, impl=0, relu=None, deq=False):
        # Linear::forward_prop(Tensor<u8>&&), fully_connected.cc:22-52
        # deq=True: also returns the dequantised result [m, n] fp32 (i8ie_fc_u8_deq)
        L = _lib.load()
        self._out_meta(x)
        m, k = x._shape
        n = self._qw_dev.shape[0]
        buf, ldx = x._as_nhwc(m, k, 1, 1, want_cp=_r16(k))
        oc, bf = self._offsets(x._zp, x.scale(), False)
        ldy = _r16(n)
        out = torch.empty(m * ldy, dtype=torch.uint8, device=buf.device)
        flags = 1 if (self.fuse_relu if relu is None else relu) else 0
        if deq:
            pc = getattr(self, "_w_scales_dev", None)
            f32 = torch.empty(m * n, dtype=torch.float32, device=buf.device)
            check(L.i8ie_fc_u8_deq(buf.data_ptr(), ldx, self._w_packed.data_ptr(), self._ldw, self._n_pad,
                                   out.data_ptr(), ldy, m, n, k, oc.data_ptr(), bf.data_ptr(), x.scale(),
                                   float(self._w_scale), pc.data_ptr() if pc is not None else None,
                                   float(self._w_scales.min()) if pc is not None else 0.0,
                                   float(self._w_scales.max()) if pc is not None else 0.0,
                                   float(self._scale), self._zp, flags, impl, f32.data_ptr(), _stream()), "fc_u8_deq")
            return _new_u8_nhwc(out, m, n, 1, 1, ldy, self._scale, self._zp, two_d=True), f32
        if getattr(self, "_w_scales_dev", None) is not None:
            check(L.i8ie_fc_u8_pc(buf.data_ptr(), ldx, self._w_packed.data_ptr(), self._ldw, self._n_pad,
                                  out.data_ptr(), ldy, m, n, k, oc.data_ptr(), bf.data_ptr(), x.scale(),
                                  self._w_scales_dev.data_ptr(), float(self._w_scales.min()), float(self._w_scales.max()),
                                  float(self._scale), self._zp, flags,
                                  acc_out.data_ptr() if acc_out is not None else None, impl, _stream()), "fc_u8_pc")
        else:
            check(L.i8ie_fc_u8(buf.data_ptr(), ldx, self._w_packed.data_ptr(), self._ldw, self._n_pad,
                               out.data_ptr(), ldy, m, n, k, oc.data_ptr(), bf.data_ptr(), x.scale(),
                               float(self._w_scale), float(self._scale), self._zp, flags,
                               acc_out.data_ptr() if acc_out is not None else None, impl, _stream()), "fc_u8")
        return _new_u8_nhwc(out, m, n, 1, 1, ldy, self._scale, self._zp, two_d=True)


class Conv2d(_BaseLayer):
    """Conv2d — include/conv2d.h, src/conv2d.cc."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        if int(stride) == 0:
            raise RuntimeError("std::exception")      # conv2d.h:12-14
        super().__init__((int(out_channels), int(in_channels), int(kernel_size), int(kernel_size)),
                         int(out_channels))
        self._stride = int(stride)
        self._pad = int(padding)

    def _pack(self):
        self._packed = {}       # input channel pitch -> packed [kc_pad, kh, kw, cp] s8 weights
        self._kc_pad = _r16(self._qw_shape[0])

    def _packed_for(self, cp):
        wp = self._packed.get(cp)
        if wp is None:
            L = _lib.load()
            kc, c, kh, kw = self._qw_shape
            wp = torch.empty(self._kc_pad * kh * kw * cp, dtype=torch.int8, device="cuda")
            check(L.i8ie_pack_conv_weight(self._qw_dev.data_ptr(), wp.data_ptr(), kc, c, kh, kw, self._kc_pad,
                                          cp, _stream()), "pack_conv_weight")
            self._packed[cp] = wp
        return wp

    def _forward_f32(self, x):
        # Conv2d::forward_prop(Tensor<float>&&), conv2d.cc:63-98: fp32 implicit GEMM
        # (no im2col buffer) with the bias add and the calibrator's min/max (conv2d.cc:94-96) fused
        L = _lib.load()
        w, b = self._fp32_params()
        if len(x._shape) != 4 or x._shape[1] != w.shape[1]:
            raise RuntimeError(f"Conv2d: input shape {tuple(x._shape)} does not match weight {tuple(w.shape)}")
        n, c, h, wd = x._shape
        kc, _, kh, kw = w.shape
        if h + 2 * self._pad < kh or wd + 2 * self._pad < kw:
            raise RuntimeError("Conv2d: kernel larger than the padded input")
        oh = (h - kh + 2 * self._pad) // self._stride + 1
        ow = (wd - kw + 2 * self._pad) // self._stride + 1
        y = torch.empty(n * kc * oh * ow, dtype=torch.float32, device=w.device)
        mm = self._cal.new_range(w.device) if self._is_preparing else None
        check(L.i8ie_conv2d_f32(x.buf.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), n, c, h, wd, kc, kh, kw,
                                self._stride, self._pad, mm.data_ptr() if mm is not None else None, _stream()),
              "conv2d_f32")
        if mm is not None:
            self._cal.sample_range(y, mm)
        return TensorF32(_Storage(y), [n, kc, oh, ow])

    def _plan(self, n, c, h, w, cp, impl, out_cp=None):
        kc, _, kh, kw = self._qw_shape
        out_cp = out_cp or _act_pitch(kc)
        key = (n, c, h, w, cp, impl, out_cp)
        p = self._plans.get(key)
        if p is None:
            L = _lib.load()
            p = L.i8ie_conv2d_plan_create(n, c, h, w, cp, kc, kh, kw, self._stride, self._pad, out_cp,
                                          self._packed_for(cp).data_ptr(), self._kc_pad, impl)
            if not p:
                raise I8ieError(f"conv2d_plan_create failed: {_lib.last_error()}")
            if getattr(self, "_w_scales_dev", None) is not None:
                check(L.i8ie_conv2d_plan_set_channel_scales(p, self._w_scales_dev.data_ptr(), float(self._w_scales.min()),
                                                            float(self._w_scales.max())), "conv2d_plan_set_channel_scales")
            self._plans[key] = p
        return p

    def __del__(self):
        try:
            L = _lib.load()
            for p in self._plans.values():
                L.i8ie_conv2d_plan_destroy(p)
        except Exception:  # noqa: BLE001
            pass

    def _out_meta(self, x):
        if len(x._shape) != 4:
            raise RuntimeError("Conv2d expects a 4-d tensor")
        n, c, h, w = x._shape
        kc, cw, kh, kw = self._qw_shape
        if c != cw:
            raise RuntimeError(f"Conv2d: input has {c} channels, weight expects {cw}")
        if h + 2 * self._pad < kh or w + 2 * self._pad < kw:
            raise RuntimeError("Conv2d: kernel larger than padded input")
        oh = (h - kh + 2 * self._pad) // self._stride + 1
        ow = (w - kw + 2 * self._pad) // self._stride + 1
        return [n, kc, oh, ow], (n, kc, oh, ow, _act_pitch(kc))

    def _forward_u8(self, x, acc_out=None, impl=0, relu=None, out_cp=None):
        # Conv2d::forward_prop(Tensor<u8>&&), conv2d.cc:100-142
        # out_cp: channel pitch of the output (default _act_pitch(kc); a pool consumer may ask for round_up(kc, 16))
        L = _lib.load()
        (n, kc, oh, ow), _ = self._out_meta(x)
        _, c, h, w = x._shape
        kh, kw = self._qw_shape[2], self._qw_shape[3]
        relu = self.fuse_relu if relu is None else relu
        quant = x._pending("quant") if (impl == 0 and acc_out is None) else None
        if quant is not None:
            # the input is a not-yet-launched quantise of an fp32 image: a stem-eligible first
            # layer consumes the fp32 image directly (quantise fused into its operand staging)
            y = self.forward_quantize_fused(quant.src, quant.scale, quant.zp, relu=relu, out_cp=out_cp)
            if y is not None:
                return y
        out_cp = out_cp or _act_pitch(kc)
        oc, _ = self._offsets(x._zp, x.scale(), True)
        # row mode (stride 1, channel count not a multiple of 128): physically padded input, K = (filter row,
        # contiguous kw * cp run) — AlexNet conv2 runs 20 K blocks instead of 25
        cpx = 0
        if impl == 0:
            cpx = int(L.i8ie_conv2d_row_mode_cp(c, _act_pitch(c), kh, kw, self._stride, self._pad, out_cp))
        elif impl == 4:      # forced (tests / probes): plan creation checks the hard constraints
            cpx = _r16(c)
        if cpx:
            xp = x._as_padded_nhwc(n, c, h, w, cpx, self._pad)
            out = _u8_empty(n * oh * ow * out_cp, xp.device)
            plan = self._plan(n, c, h, w, cpx, 4, out_cp)
            self._last_impl = 4
            check(L.i8ie_conv2d_u8(plan, xp.data_ptr(), out.data_ptr(), oc.data_ptr(), x.scale(),
                                   float(self._w_scale), float(self._scale), x._zp, self._zp, 1 if relu else 0,
                                   acc_out.data_ptr() if acc_out is not None else None, _stream()), "conv2d_u8 (row mode)")
            return _new_u8_nhwc(out, n, kc, oh, ow, out_cp, self._scale, self._zp)
        buf, cp = x._as_nhwc(n, c, h, w)
        out = _u8_empty(n * oh * ow * out_cp, buf.device)
        plan = self._plan(n, c, h, w, cp, impl, out_cp)
        self._last_impl = int(L.i8ie_conv2d_plan_impl(plan))   # 1 SIMT, 2 tcgen05 im2col, 3 tcgen05 stem
        flags = 1 if relu else 0
        check(L.i8ie_conv2d_u8(plan, buf.data_ptr(), out.data_ptr(), oc.data_ptr(), x.scale(),
                               float(self._w_scale), float(self._scale), x._zp, self._zp, flags,
                               acc_out.data_ptr() if acc_out is not None else None, _stream()), "conv2d_u8")
        return _new_u8_nhwc(out, n, kc, oh, ow, out_cp, self._scale, self._zp)

    def forward_quantize_fused(self, x, in_scale, in_zp, acc_out=None, relu=None, out_cp=None):
        """Module.__call__'s input quantise (module.py:20) fused into this convolution
        (i8ie_conv2d_f32_u8). Only stem-eligible layers (C <= 4, stride 4/8) support it;
        returns None otherwise so the caller falls back to quantize() + __call__()."""
        L = _lib.load()
        if not self._is_quantized or not isinstance(x, TensorF32) or len(x._shape) != 4:
            return None
        n, c, h, w = x._shape
        kc, cw, kh, kw = self._qw_shape
        if c != cw or c > 4 or self._stride not in (4, 8):
            return None
        out_cp = out_cp or _act_pitch(kc)
        plan = self._plan(n, c, h, w, _r16(c), 0, out_cp)
        if int(L.i8ie_conv2d_plan_impl(plan)) != 3:
            return None
        in_scale = float(np.float32(in_scale))
        oh = (h - kh + 2 * self._pad) // self._stride + 1
        ow = (w - kw + 2 * self._pad) // self._stride + 1
        oc, _ = self._offsets(int(in_zp), in_scale, True)
        ptr, slot, dev = _src_ptrs(x)
        out = _u8_empty(n * oh * ow * out_cp, dev)
        flags = 1 if (self.fuse_relu if relu is None else relu) else 0
        fn = L.i8ie_conv2d_f32_u8_indirect if slot is not None else L.i8ie_conv2d_f32_u8
        check(fn(plan, slot if slot is not None else ptr, in_scale, int(in_zp), out.data_ptr(), oc.data_ptr(),
                 float(self._w_scale), float(self._scale), self._zp, flags,
                 acc_out.data_ptr() if acc_out is not None else None, _stream()),
              "conv2d_f32_u8")
        self._last_impl = 3
        return _new_u8_nhwc(out, n, kc, oh, ow, out_cp, self._scale, self._zp)


# ---- CUDA-graph replay of a whole quantised forward (used by api.Module.__call__) ------------

MAX_DIRECT_GRAPHS = int(os.environ.get("I8IE_DIRECT_GRAPHS", "16"))   # graphs with the input address baked in, per (shape, device, epoch)


class _GraphOutStorage:
    """Storage of a CUDA-graph result. It aliases the graph's static output buffer until that
    graph is replayed again; the replay first gives a result that is still referenced its own copy
    (copy-on-overwrite). A result therefore behaves like the fresh tensor the reference returns,
    while the usual loop `y = model(x).numpy()` — result consumed and dropped before the next
    call — costs no device copy per step."""
    __slots__ = ("_t", "views", "_own", "__weakref__")

    def __init__(self, t):
        self._t, self.views, self._own = t, 0, False

    @property
    def t(self):
        return self._t

    def detach(self):
        if not self._own:
            self._t = self._t.clone()      # stream-ordered before the replay that overwrites the buffer
            self._own = True


def graphable(x):
    return isinstance(x, TensorF32) and torch.cuda.is_available()


def _capture(fn, x, st, slot):
    """Captures fn(x) into a CUDA graph. The warm-up calls before this have already created every
    plan / offset table / packed weight, so nothing allocates through the C ABI or synchronises
    while the stream is capturing. slot is None: x's buffer address is baked into the graph
    (replayed only for inputs at that address); otherwise the input is addressed through the device
    slot (_SlotStorage) and a replay works on any buffer. All graphs of one model state share a
    memory pool (they never run concurrently); each keeps its own output buffer alive."""
    from .api import Tensor
    dev = x.buf.device
    torch.cuda.synchronize()
    before = _lib.launch_count()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, pool=st.get("pool")):
        src = _Storage(x.buf) if slot is None else _SlotStorage(slot, x.buf.numel(), dev)
        out = fn(Tensor(TensorF32(src, x._shape)))
        out_buf = out.data.buf
    kernels = _lib.launch_count() - before
    if not isinstance(out.data, TensorF32):
        raise RuntimeError("forward() did not return a dequantised tensor")
    st.setdefault("pool", g.pool())
    return {"graph": g, "out_buf": out_buf, "out_shape": list(out.data._shape), "kernels": int(kernels),
            "last_out": None}


def capture_forward(fn, x):
    """First capture for an input shape: the graph for this input address (state for run_graphed)."""
    st = {"direct": {}, "indirect": None, "slot": None, "launched": 0, "pool": None}
    if x.buf.data_ptr() % 16 == 0:
        st["direct"][x.buf.data_ptr()] = _capture(fn, x, st, None)
    else:
        _indirect_graph(fn, x, st)
    st["graph"] = True
    return st


def _indirect_graph(fn, x, st):
    if st["indirect"] is None:
        st["slot"] = torch.zeros(2, dtype=torch.int64, device=x.buf.device)   # 16-byte slot; [0] = source address
        keep = x.buf if x.buf.data_ptr() % 16 == 0 else x.buf.clone()
        st["slot"][0] = keep.data_ptr()
        st["indirect"] = _capture(fn, TensorF32(_Storage(keep), x._shape), st, st["slot"])
    return st["indirect"]


def replay_forward(st, x, fn):
    """Runs the captured forward on x. Inputs at an address seen before replay the graph that has
    that address baked in: one graph launch, nothing else on the stream. A new address gets its own
    graph (up to MAX_DIRECT_GRAPHS — a serving loop cycles through a few allocator blocks); beyond
    that, and for buffers that are not 16-byte aligned, the slot graph runs: one tiny fill kernel
    passes the address (and an unaligned buffer is copied first)."""
    ptr = x.buf.data_ptr()
    e = st["direct"].get(ptr)
    if e is None and ptr % 16 == 0 and len(st["direct"]) < MAX_DIRECT_GRAPHS:
        e = st["direct"][ptr] = _capture(fn, x, st, None)   # (the caller's buffer is not kept alive: a graph only
        #                                                      replays for a live tensor at exactly this address)
    if e is None:
        e = _indirect_graph(fn, x, st)
        src = x.buf if ptr % 16 == 0 else x.buf.clone()
        st["slot"][:1].fill_(src.data_ptr())      # stream-ordered: the value travels as a kernel argument
    prev = e["last_out"]() if e["last_out"] is not None else None
    if prev is not None:
        prev.detach()                              # a live result of this graph keeps its values
    e["graph"].replay()
    st["launched"] += e["kernels"]
    stor = _GraphOutStorage(e["out_buf"])
    e["last_out"] = weakref.ref(stor)
    return TensorF32(stor, e["out_shape"])
