"""Helpers for the -m gpu parity tests: everything goes through the C ABI via
int8inferenceengine_b200.backend / _lib, and is compared with the oracle."""
import ctypes

import numpy as np
import torch

from int8inferenceengine_b200 import _lib, backend as B


def lib():
    return _lib.load()


def stream():
    return torch.cuda.current_stream().cuda_stream


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def u8_tensor_from_nchw(q, scale, zp):
    """Backend u8 tensor (dense logical layout) from a numpy u8 array."""
    t = dev(q.reshape(-1))
    return B.TensorU8(B._Storage(t), list(q.shape), "dense", None, scale, zp)


def make_layer(kind, w, b, qp, stride=1, pad=0):
    if kind == "conv":
        kc, c, k, _ = w.shape
        L = B.Conv2d(c, kc, k, stride, pad)
    else:
        n, k = w.shape
        L = B.Linear(k, n)
    L.load_weight(w)
    L.load_bias(b)
    L.set_qparams(*qp)
    L.convert()
    return L
