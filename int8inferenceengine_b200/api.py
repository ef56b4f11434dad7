"""The `i8ie` user API on the B200 backend — same names, signatures and behaviour as the
reference's Python package (i8ie/__init__.py:1-32, tensor.py:4-37, layer.py:5-35,
module.py:6-35), so existing scripts run unchanged with `import i8ie`.
"""
from __future__ import annotations

from . import backend as _B
from .workloads import INPUT_SCALE, INPUT_ZP

__all__ = ["tensor", "argmax", "relu", "max_pool2d", "Linear", "Conv2d", "Tensor", "quantize",
           "dequantize", "Module"]


class Tensor:
    """i8ie/tensor.py:4-37 — thin handle around a backend tensor (`.data`)."""

    def __init__(self, data):
        self.data = data

    def __repr__(self):
        return ((self.numpy() - self.zero_point) * self.scale).__repr__()

    def __eq__(self, obj):
        # returns a RAW backend float tensor, as the reference does (tensor.py:11-12);
        # the notebooks call `.sum()` on it.
        return _B.tensor(self.numpy() == obj.numpy())

    __hash__ = None

    def reshape(self, *args):
        return Tensor(self.data.reshape(list(args)))

    def numpy(self):
        return self.data.numpy()

    def sum(self):
        return self.numpy().sum()

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def scale(self):
        return self.data.scale()

    @property
    def zero_point(self):
        return self.data.zero_point()

    @property
    def dtype(self):
        pass


def tensor(ndarray):
    return Tensor(_B.tensor(ndarray))


def argmax(x, *args, **kwargs):
    return tensor(x.numpy().argmax(*args, **kwargs))


def relu(x):
    return Tensor(_B.relu(x.data))


def max_pool2d(x, kernel_size, stride):
    return Tensor(_B.max_pool2d(x.data, kernel_size, stride))


def quantize(x, scale, zero_point):
    return Tensor(_B.quantize(x.data, scale, zero_point))


def dequantize(x):
    return Tensor(_B.dequantize(x.data))


class Layer:
    """i8ie/layer.py:5-19."""

    def __call__(self, x):
        return Tensor(self.layer(x.data))

    def load_weight(self, weight):
        self.layer.load_weight(weight)

    def load_bias(self, bias):
        self.layer.load_bias(bias)

    def prepare(self):
        self.layer.prepare()

    def convert(self):
        self.layer.convert()


class Linear(Layer):
    def __init__(self, in_channels, out_channels):
        self.layer = _B.Linear(in_channels, out_channels)


class Conv2d(Layer):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        self.layer = _B.Conv2d(in_channels, out_channels, kernel_size, stride, padding)


class Module:
    """i8ie/module.py:6-35."""

    def __init__(self):
        self.is_quant = False

    def load(self, state_dict):
        for key in state_dict:
            name, attr = key.split(".")
            if attr == "weight":
                self.__dict__[name].load_weight(state_dict[key])
            elif attr == "bias":
                self.__dict__[name].load_bias(state_dict[key])

    # Extension (not in the reference): after two eager calls with the same input shape the
    # quantised forward is captured into a CUDA graph and replayed (the reference runs one op
    # at a time under the GIL). Results are identical as long as forward() is a pure function of
    # its input and the layers' state: the cache is keyed on (shape, device, backend.graph_epoch()),
    # and the epoch moves whenever a layer's (scale, zero_point), weights or fuse_relu change, so a
    # stale graph is never replayed. A forward() with Python-side state or data-dependent control
    # flow must set `graph = False`.
    graph = True

    def __call__(self, x):
        if self.is_quant and self.graph and getattr(self, "record", None) is None and _B.graphable(x.data):
            parts = None if self.__dict__.get("_no_chunks") else _B.pending_chunks(x.data)
            if parts:
                return self._call_chunked(x, parts)
            return self._call_graphed(x)
        return self._call_eager(x)

    def _call_chunked(self, x, parts):
        # Extension: the batch is still arriving from pinned host memory in row chunks
        # (backend._ChunkedStorage). Images are independent and the quantisation parameters are
        # per tensor, so running the chunks one after the other gives the same bits as one call
        # while the forward of chunk k overlaps the host->device copy of chunk k+1.
        outs = []
        for sub, ev in parts:
            _B.wait_event(ev)
            outs.append(self._call_graphed(Tensor(sub)).data)
        _B.chunks_consumed(x.data)
        if any(not isinstance(o, _B.TensorF32) or len(o._shape) < 1 or o._shape[0] != p[0]._shape[0]
               for o, p in zip(outs, parts)):
            # a forward() that does not keep the batch dimension cannot be split: run it whole from now on
            self.__dict__["_no_chunks"] = True
            return self._call_graphed(x)
        return Tensor(_B.concat_rows(outs))

    def _call_eager(self, x):
        if self.is_quant:
            x = Tensor(_B.quantize(x.data, INPUT_SCALE, INPUT_ZP))   # module.py:20 (hard-coded 0.025/127)
        x = self.forward(x)
        if self.is_quant:
            x = Tensor(_B.dequantize(x.data))
        return x

    def _call_graphed(self, x):
        cache = self.__dict__.setdefault("_graphs", {})
        key = (tuple(x.data.shape), x.data.buf.device.index, _B.graph_epoch())
        if cache and next(iter(cache))[2] != key[2]:
            cache.clear()                            # graphs of an older epoch have stale parameters baked in
        st = cache.setdefault(key, {"calls": 0, "graph": None})
        if st["graph"] is None:
            st["calls"] += 1
            if st["calls"] <= 2:
                return self._call_eager(x)          # warm-up: builds plans, offsets, packed weights
            try:
                st.update(_B.capture_forward(self._call_eager, x.data))
                # the graph holds raw pointers into the layers' plans and packed weights
                st["keep"] = [v for v in self.__dict__.values() if issubclass(type(v), Layer)]
            except Exception as e:  # noqa: BLE001 - e.g. a forward() that leaves the engine mid-way
                import warnings
                warnings.warn(f"i8ie: CUDA-graph capture of {type(self).__name__}.forward failed ({e}); "
                              "running eagerly")
                self.graph = False
                return self._call_eager(x)
        return Tensor(_B.replay_forward(st, x.data, self._call_eager))

    def graph_launches(self):
        """Kernels replayed through CUDA graphs so far (they bypass the C-ABI launch counter)."""
        return sum(st.get("launched", 0) for st in self.__dict__.get("_graphs", {}).values())

    def prepare(self):
        for _, val in self.__dict__.items():
            if issubclass(type(val), Layer):
                val.prepare()

    def convert(self):
        for _, val in self.__dict__.items():
            if issubclass(type(val), Layer):
                val.convert()
        self.is_quant = True
