#!/usr/bin/env python3
"""Generates the golden vectors under tests/golden/ from the COMPILED REFERENCE.

Run in the build container (needs oracle/_ref, i.e. /root/reference + `make -C oracle ref`):
    PYTHONPATH=/root/repo python tests/golden/make_golden.py
Every expected value below is an output of the reference's own src/*.cc
(unmodified, stand-in GEMM for MKL — exact for the integer path). The reference
ships no golden vectors of its own, so these files are the pin
for the oracle (tests/test_oracle_golden.py) and for the CUDA path (-m gpu tests).
Inputs are seeded; files are small (<1 MB total) and committed.

Note: the reference's calibrator is randomised once a layer emits >1000 values
(calibrator.cc:9-22), so per-layer (scale, zero_point) are RECORDED here rather
than recomputed; all cases whose ranges are deterministic (<=1000 samples) also
pin calibrator.cc:24-37 itself.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from int8inferenceengine_b200 import workloads as W  # noqa: E402
from oracle import models, ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
m = ref.module()


def T(a):
    return m.tensor(np.ascontiguousarray(a, np.float32))


def npy(t):
    return np.array(t.numpy(), copy=True)


def save(name, **kw):
    p = os.path.join(OUT, name + ".npz")
    np.savez_compressed(p, **kw)
    print(f"{name}.npz  {os.path.getsize(p)} B")


def elementwise():
    rng = np.random.default_rng(100)
    x = rng.uniform(-3.1, 3.1, size=(3, 5, 7)).astype(np.float32)
    # wrap cases of the unclamped quantise (quantize_utils.cc:49): SURVEY App. C
    edge = np.array([-4, -3.2, -3.175, 0, 3.2, 3.3, 10, -0.0124, 0.0126, 1e-7], np.float32)
    out = {}
    for tag, arr, s, z in [("a", x, 0.025, 127), ("b", x, 0.0371, 100), ("edge", edge, 0.025, 127),
                           ("c", x * 40, 1.0, 0)]:
        q = m.quantize(T(arr), s, z)
        out[f"q_{tag}_x"] = arr
        out[f"q_{tag}_sz"] = np.array([s, z], np.float64)
        out[f"q_{tag}_q"] = npy(q)
        out[f"q_{tag}_deq"] = npy(m.dequantize(q))
    # relu<u8> / max_pool2d<u8> (functional.cc:15-64) on a quantised tensor
    y = rng.uniform(-3, 3, size=(2, 3, 13, 13)).astype(np.float32)
    q = m.quantize(T(y), 0.025, 127)
    out["fn_x"] = y
    out["fn_q"] = npy(q)
    out["fn_relu"] = npy(m.relu(q))
    out["fn_pool32"] = npy(m.max_pool2d(q, 3, 2))
    out["fn_pool22"] = npy(m.max_pool2d(q, 2, 2))
    out["fn_pool31"] = npy(m.max_pool2d(q, 3, 1))
    save("kat_elementwise", **out)


def get_range():
    """calibrator.cc:24-37 through an identity Linear (fp32 forward reproduces x exactly)."""
    rng = np.random.default_rng(101)
    cases = {
        "full1000": rng.uniform(-3, 5, size=(100, 10)),
        "lt1000": rng.uniform(-3, 5, size=(30, 10)),      # quirk: max read from the zero tail
        "all_pos": rng.uniform(0.5, 4, size=(100, 10)),
        "all_neg": rng.uniform(-4, -0.5, size=(100, 10)),
        "all_zero": np.zeros((100, 10)),
        "tiny": rng.uniform(-1e-3, 1e-3, size=(100, 10)),
        "wide": rng.uniform(-300, 900, size=(100, 10)),
        "lt1000_neg": rng.uniform(-4, -0.5, size=(7, 10)),
    }
    out = {}
    for tag, arr in cases.items():
        arr = arr.astype(np.float32)
        L = m.Linear(10, 10)
        L.load_weight(np.eye(10, dtype=np.float32))
        L.load_bias(np.zeros(10, np.float32))
        L.prepare()
        y = npy(L(T(arr)))
        assert np.array_equal(y, arr)
        L.convert()
        o = L(m.quantize(T(arr[:1]), 0.025, 127))
        out[f"{tag}_samples"] = arr
        out[f"{tag}_sz"] = np.array([o.scale(), o.zero_point()], np.float64)
    save("kat_get_range", **out)


def linear():
    rng = np.random.default_rng(102)
    out = {}
    for tag, (M, K, N, mcal) in {"a": (7, 300, 24, 40), "b": (16, 784, 10, 100), "c": (3, 65, 130, 7)}.items():
        w = rng.uniform(-0.2, 0.2, size=(N, K)).astype(np.float32)
        b = rng.uniform(-0.05, 0.05, size=(N,)).astype(np.float32)
        xcal = rng.uniform(-2, 2, size=(mcal, K)).astype(np.float32)
        x = rng.uniform(-2, 2, size=(M, K)).astype(np.float32)
        L = m.Linear(K, N)
        L.load_weight(w)
        L.load_bias(b)
        L.prepare()
        ycal = npy(L(T(xcal)))
        L.convert()
        q = m.quantize(T(x), 0.025, 127)
        o = L(q)
        out.update({f"{tag}_w": w, f"{tag}_b": b, f"{tag}_xcal": xcal, f"{tag}_x": x,
                    f"{tag}_ycal": ycal, f"{tag}_qin": npy(q), f"{tag}_out": npy(o),
                    f"{tag}_sz": np.array([o.scale(), o.zero_point()], np.float64),
                    f"{tag}_deq": npy(m.dequantize(o))})
    save("kat_linear", **out)


def conv():
    rng = np.random.default_rng(103)
    geoms = {  # (n, c, h, w, kc, k, stride, pad)   SURVEY §8c probe geometries (shrunk)
        "k5": (2, 3, 12, 12, 8, 5, 1, 0),
        "k11s4p2": (2, 3, 35, 35, 16, 11, 4, 2),
        "k5p2": (2, 20, 9, 9, 12, 5, 1, 2),
        "k3s7p3": (2, 10, 22, 22, 20, 3, 7, 3),
        "k3p1": (3, 32, 6, 6, 48, 3, 1, 1),
        "k3p1_c64": (2, 64, 13, 13, 32, 3, 1, 1),
    }
    out = {}
    for tag, (n, c, h, w_, kc, k, s, p) in geoms.items():
        a = np.sqrt(6.0 / (c * k * k))
        w = rng.uniform(-a, a, size=(kc, c, k, k)).astype(np.float32)
        b = rng.uniform(-0.05, 0.05, size=(kc,)).astype(np.float32)
        xcal = rng.uniform(-2.1, 2.6, size=(4, c, h, w_)).astype(np.float32)
        x = rng.uniform(-2.1, 2.6, size=(n, c, h, w_)).astype(np.float32)
        L = m.Conv2d(c, kc, k, s, p)
        L.load_weight(w)
        L.load_bias(b)
        L.prepare()
        L(T(xcal))
        L.convert()
        q = m.quantize(T(x), 0.025, 127)
        o = L(q)
        # a second pass with a different input zero-point exercises the zp padding/offset
        q2 = m.quantize(T(x), 0.031, 90)
        o2 = L(q2)
        out.update({f"{tag}_geom": np.array([n, c, h, w_, kc, k, s, p]), f"{tag}_w": w, f"{tag}_b": b,
                    f"{tag}_x": x, f"{tag}_qin": npy(q), f"{tag}_out": npy(o),
                    f"{tag}_sz": np.array([o.scale(), o.zero_point()], np.float64),
                    f"{tag}_qin2": npy(q2), f"{tag}_out2": npy(o2)})
    save("kat_conv", **out)


def nets():
    for topo, bcal, b in [("fc_mnist", 100, 16), ("lenet", 100, 4), ("simple_conv", 100, 4), ("mini_alex", 100, 4)]:
        sd = W.make_weights(topo, 0)
        r = models.RefModel(topo, sd)
        r.calibrate(W.make_images(topo, bcal, 1))
        x = W.make_images(topo, b, 2)
        logits, recs = r.forward_int8(x, record=True)
        qp = r.qparams(x)
        out = {"batch": np.array(b), "logits": logits,
               "qp_names": np.array(list(qp.keys())),
               "qp_scale": np.array([v[0] for v in qp.values()], np.float32),
               "qp_zp": np.array([v[1] for v in qp.values()], np.int32)}
        for i, (tag, arr, s, z) in enumerate(recs):
            out[f"op{i:02d}_{tag}"] = arr
        save(f"net_{topo}", **out)


def alexnet():
    """Full AlexNet-224 (BASELINE config 3 shapes): weights seed 0, calibration batch 100
    seed 1, evaluation batch 2 seed 2. Activations are too large to commit, so per-op
    SHA-256 digests + the logits + the recorded (scale, zp) are stored."""
    topo = "alexnet"
    sd = W.make_weights(topo, 0)
    r = models.RefModel(topo, sd)
    r.calibrate(W.make_images(topo, 100, 1))
    x = W.make_images(topo, 2, 2)
    logits, recs = r.forward_int8(x, record=True)
    qp = r.qparams(x)
    out = {"batch": np.array(2), "logits": logits,
           "qp_names": np.array(list(qp.keys())),
           "qp_scale": np.array([v[0] for v in qp.values()], np.float32),
           "qp_zp": np.array([v[1] for v in qp.values()], np.int32),
           "op_tags": np.array([t for t, *_ in recs]),
           "op_sha256": np.array([hashlib.sha256(a.tobytes()).hexdigest() for _, a, _, _ in recs]),
           "op_sum": np.array([int(a.astype(np.int64).sum()) for _, a, _, _ in recs], np.int64),
           "final_u8": recs[-1][1]}
    save("net_alexnet", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["elementwise", "get_range", "linear", "conv", "nets", "alexnet"]
    for w in which:
        globals()[w]()
