"""ctypes binding of the C restatement (oracle/i8ie_oracle.c). TEST INFRASTRUCTURE ONLY.

All arrays are numpy, layouts are the reference's (NCHW activations, OIHW weights).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_i8ie.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_s8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_s32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build():
    """Compile the C restatement (gcc only; seconds)."""
    subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        i, f, i64, vp = C.c_int, C.c_float, C.c_int64, C.c_void_p
        L.orc_quantize_f32_u8.argtypes = [_f32p, _u8p, i64, f, i]
        L.orc_quantize_f32_u8_clamped.argtypes = [_f32p, _u8p, i64, f, i]
        L.orc_quantize_f32_s8_clamped.argtypes = [_f32p, _s8p, i64, f]
        L.orc_dequantize_u8_f32.argtypes = [_u8p, _f32p, i64, f, i]
        L.orc_dequantize_s32_f32.argtypes = [_s32p, _f32p, i64, f, f]
        L.orc_down_scale.argtypes = [_s32p, _u8p, i64, f, f, f, i]
        L.orc_quantize_weight.argtypes = [_f32p, i64, _f32p, i64, _s8p, _s8p]
        L.orc_quantize_weight.restype = f
        L.orc_conv_offsets.argtypes = [_s8p, _s8p, i, i, i, f, _s32p]
        L.orc_fc_offsets.argtypes = [_s8p, i, i, i, _s32p]
        L.orc_conv2d_u8.argtypes = [_u8p, i, i, i, i, _s8p, _s8p, i, i, i, i, i, f, i, f, f, i, _u8p, vp]
        L.orc_linear_u8.argtypes = [_u8p, i, i, _s8p, _s8p, i, f, i, f, f, i, _u8p, vp]
        L.orc_relu_u8.argtypes = [_u8p, _u8p, i64, i]
        L.orc_max_pool2d_u8.argtypes = [_u8p, i, i, i, i, i, i, _u8p]
        L.orc_get_range.argtypes = [_f32p, i64, C.POINTER(f), C.POINTER(C.c_uint8)]
        L.orc_get_range.restype = i
        L.orc_get_range_minmax.argtypes = [f, f, C.POINTER(f), C.POINTER(C.c_uint8)]
        L.orc_conv2d_f32.argtypes = [_f32p, i, i, i, i, _f32p, _f32p, i, i, i, i, i, _f32p]
        L.orc_linear_f32.argtypes = [_f32p, i, i, _f32p, _f32p, i, _f32p]
        L.orc_relu_f32.argtypes = [_f32p, _f32p, i64]
        L.orc_max_pool2d_f32.argtypes = [_f32p, i, i, i, i, i, i, _f32p]
        _lib = L
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def quantize(x, scale, zp):
    x = _c(x, np.float32)
    q = np.empty(x.shape, np.uint8)
    lib().orc_quantize_f32_u8(x, q, x.size, scale, int(zp))
    return q


def quantize_u8_clamped(x, scale, zp):
    x = _c(x, np.float32)
    q = np.empty(x.shape, np.uint8)
    lib().orc_quantize_f32_u8_clamped(x, q, x.size, scale, int(zp))
    return q


def quantize_s8_clamped(x, scale):
    x = _c(x, np.float32)
    q = np.empty(x.shape, np.int8)
    lib().orc_quantize_f32_s8_clamped(x, q, x.size, scale)
    return q


def dequantize(q, scale, zp):
    q = _c(q, np.uint8)
    x = np.empty(q.shape, np.float32)
    lib().orc_dequantize_u8_f32(q, x, q.size, scale, int(zp))
    return x


def down_scale(acc, sa, sb, sc, zp_c):
    acc = _c(acc, np.int32)
    out = np.empty(acc.shape, np.uint8)
    lib().orc_down_scale(acc, out, acc.size, sa, sb, sc, int(zp_c))
    return out


def quantize_weight(w, b):
    """-> (qw s8, qb s8, scale f32)  [layer.cc:6-26]"""
    w = _c(w, np.float32)
    b = _c(b, np.float32)
    qw = np.empty(w.shape, np.int8)
    qb = np.empty(b.shape, np.int8)
    s = lib().orc_quantize_weight(w, w.size, b, b.size, qw, qb)
    return qw, qb, np.float32(s)


# ---- F4 extension (NOT in the reference): per-output-channel weight scales -------------------------
# Pure compositions of the pinned functions above: layer.cc:6-26 applied to each output channel's
# weights + its bias entry on its own, and quantize_utils.cc:27-36 applied per channel with that
# channel's weight scale. The integer GEMM, the oc offsets and the fc float-bias add do not involve
# the weight scale, so the accumulators come from the unchanged restatement.
def quantize_weight_per_channel(w, b):
    """-> (qw s8, qb s8, scales f32[n])"""
    w = _c(w, np.float32)
    b = _c(b, np.float32)
    qw = np.empty(w.shape, np.int8)
    qb = np.empty(b.shape, np.int8)
    sc = np.empty(w.shape[0], np.float32)
    for j in range(w.shape[0]):
        qwj, qbj, sj = quantize_weight(w[j].reshape(-1), b[j:j + 1])
        qw[j] = qwj.reshape(w[j].shape)
        qb[j] = qbj[0]
        sc[j] = sj
    return qw, qb, sc


def conv2d_u8_pc(x, qw, qb, stride, pad, in_scale, in_zp, w_scales, out_scale, out_zp, want_acc=False):
    _, acc = conv2d_u8(x, qw, qb, stride, pad, in_scale, in_zp, np.float32(1.0), out_scale, out_zp, want_acc=True)
    n, _, kc = acc.shape
    oh = (x.shape[2] - qw.shape[2] + 2 * pad) // stride + 1
    ow = (x.shape[3] - qw.shape[3] + 2 * pad) // stride + 1
    out = np.empty((n, kc, oh, ow), np.uint8)
    for j in range(kc):
        out[:, j] = down_scale(np.ascontiguousarray(acc[:, :, j]), in_scale, w_scales[j], out_scale, out_zp).reshape(n, oh, ow)
    return (out, acc) if want_acc else out


def linear_u8_pc(x, qw, qb, in_scale, in_zp, w_scales, out_scale, out_zp, want_acc=False):
    _, acc = linear_u8(x, qw, qb, in_scale, in_zp, np.float32(1.0), out_scale, out_zp, want_acc=True)
    out = np.empty(acc.shape, np.uint8)
    for j in range(acc.shape[1]):
        out[:, j] = down_scale(np.ascontiguousarray(acc[:, j]), in_scale, w_scales[j], out_scale, out_zp)
    return (out, acc) if want_acc else out


def conv_offsets(qw, qb, in_zp, in_scale):
    qw = _c(qw, np.int8)
    kc = qw.shape[0]
    K = qw.size // kc
    oc = np.empty(kc, np.int32)
    lib().orc_conv_offsets(qw, _c(qb, np.int8), kc, K, int(in_zp), in_scale, oc)
    return oc


def fc_offsets(qw, in_zp):
    qw = _c(qw, np.int8)
    n, K = qw.shape
    oc = np.empty(n, np.int32)
    lib().orc_fc_offsets(qw, n, K, int(in_zp), oc)
    return oc


def conv2d_u8(x, qw, qb, stride, pad, in_scale, in_zp, w_scale, out_scale, out_zp, want_acc=False):
    """x u8 NCHW, qw s8 OIHW -> u8 NCHW (and the s32 accumulators [n, oh*ow, kc])."""
    x = _c(x, np.uint8)
    qw = _c(qw, np.int8)
    n, c, h, w = x.shape
    kc, c2, kh, kw = qw.shape
    assert c == c2
    oh = (h - kh + 2 * pad) // stride + 1
    ow = (w - kw + 2 * pad) // stride + 1
    out = np.empty((n, kc, oh, ow), np.uint8)
    acc = np.empty((n, oh * ow, kc), np.int32) if want_acc else None
    lib().orc_conv2d_u8(x, n, c, h, w, qw, _c(qb, np.int8), kc, kh, kw, stride, pad,
                        in_scale, int(in_zp), w_scale, out_scale, int(out_zp), out,
                        acc.ctypes.data if want_acc else None)
    return (out, acc) if want_acc else out


def linear_u8(x, qw, qb, in_scale, in_zp, w_scale, out_scale, out_zp, want_acc=False):
    x = _c(x, np.uint8)
    qw = _c(qw, np.int8)
    m, k = x.shape
    n, k2 = qw.shape
    assert k == k2
    out = np.empty((m, n), np.uint8)
    acc = np.empty((m, n), np.int32) if want_acc else None
    lib().orc_linear_u8(x, m, k, qw, _c(qb, np.int8), n, in_scale, int(in_zp), w_scale,
                        out_scale, int(out_zp), out, acc.ctypes.data if want_acc else None)
    return (out, acc) if want_acc else out


def relu_u8(x, zp):
    x = _c(x, np.uint8)
    out = np.empty(x.shape, np.uint8)
    lib().orc_relu_u8(x, out, x.size, int(zp))
    return out


def max_pool2d_u8(x, k, s):
    x = _c(x, np.uint8)
    n, c, h, w = x.shape
    out = np.empty((n, c, (h - k) // s + 1, (w - k) // s + 1), np.uint8)
    lib().orc_max_pool2d_u8(x, n, c, h, w, k, s, out)
    return out


def get_range(samples):
    """calibrator.cc:24-37 for <=1000 collected samples -> (scale f32, zp int)."""
    s = _c(np.asarray(samples).ravel(), np.float32)
    sc, zp = C.c_float(), C.c_uint8()
    if lib().orc_get_range(s, s.size, C.byref(sc), C.byref(zp)) != 0:
        raise ValueError("get_range: sample count must be in [1, 1000]")
    return np.float32(sc.value), int(zp.value)


def get_range_minmax(mn, mx):
    sc, zp = C.c_float(), C.c_uint8()
    lib().orc_get_range_minmax(float(np.float32(mn)), float(np.float32(mx)), C.byref(sc), C.byref(zp))
    return np.float32(sc.value), int(zp.value)


def conv2d_f32(x, w, b, stride, pad):
    x = _c(x, np.float32)
    w = _c(w, np.float32)
    n, c, h, wd = x.shape
    kc, _, kh, kw = w.shape
    oh = (h - kh + 2 * pad) // stride + 1
    ow = (wd - kw + 2 * pad) // stride + 1
    out = np.empty((n, kc, oh, ow), np.float32)
    lib().orc_conv2d_f32(x, n, c, h, wd, w, _c(b, np.float32), kc, kh, kw, stride, pad, out)
    return out


def linear_f32(x, w, b):
    x = _c(x, np.float32)
    w = _c(w, np.float32)
    out = np.empty((x.shape[0], w.shape[0]), np.float32)
    lib().orc_linear_f32(x, x.shape[0], x.shape[1], w, _c(b, np.float32), w.shape[0], out)
    return out


def relu_f32(x):
    x = _c(x, np.float32)
    out = np.empty_like(x)
    lib().orc_relu_f32(x, out, x.size)
    return out


def max_pool2d_f32(x, k, s):
    x = _c(x, np.float32)
    n, c, h, w = x.shape
    out = np.empty((n, c, (h - k) // s + 1, (w - k) // s + 1), np.float32)
    lib().orc_max_pool2d_f32(x, n, c, h, w, k, s, out)
    return out
