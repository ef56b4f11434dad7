"""Live cross-check of the C restatement against the compiled reference
(oracle/_ref). Skipped when that build is absent. CPU only."""
import numpy as np
import pytest

from int8inferenceengine_b200 import workloads as W
from oracle import models, port, ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("topo,bcal,b", [("fc_mnist", 100, 32), ("lenet", 100, 6), ("simple_conv", 100, 5), ("mini_alex", 100, 3)])
def test_port_matches_reference_layer_by_layer(topo, bcal, b):
    sd = W.make_weights(topo, 7)
    r = models.RefModel(topo, sd)
    r.calibrate(W.make_images(topo, bcal, 8))
    x = W.make_images(topo, b, 9)
    lr, rr = r.forward_int8(x, record=True)
    p = models.PortModel(topo, sd)
    p.convert(r.qparams(x))
    lp, rp = p.forward_int8(x, record=True)
    for (ta, a, sa, za), (tb, bb, sb, zb) in zip(rr, rp):
        assert ta == tb and sa == sb and za == zb
        assert np.array_equal(a, bb), ta
    assert np.array_equal(lr, lp)


def test_fp32_forward_close():
    # reference tests use atol 0.1 for fp32 layers (unittest/test_layers.py:10-11)
    sd = W.make_weights("simple_conv", 3)
    x = W.make_images("simple_conv", 3, 4)
    a = models.RefModel("simple_conv", sd).forward_fp32(x)
    b = models.PortModel("simple_conv", sd).forward_fp32(x)
    assert np.allclose(a, b, atol=1e-3)


def test_random_conv_geometries():
    m = ref.module()
    rng = np.random.default_rng(5)
    for _ in range(12):
        c, kc = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        k = int(rng.integers(1, 6))
        s = int(rng.integers(1, 4))
        p = int(rng.integers(0, 3))
        h, w_ = int(rng.integers(k, k + 9)), int(rng.integers(k, k + 9))
        wt = rng.uniform(-0.5, 0.5, size=(kc, c, k, k)).astype(np.float32)
        b = rng.uniform(-0.1, 0.1, size=(kc,)).astype(np.float32)
        x = rng.uniform(-3, 3, size=(2, c, h, w_)).astype(np.float32)
        L = m.Conv2d(c, kc, k, s, p)
        L.load_weight(wt)
        L.load_bias(b)
        L.prepare()
        L(m.tensor(x))
        L.convert()
        zp = int(rng.integers(0, 256))
        q = m.quantize(m.tensor(x), 0.03, zp)
        o = L(q)
        qw, qb, ws = port.quantize_weight(wt, b)
        got = port.conv2d_u8(np.array(q.numpy()), qw, qb, s, p, np.float32(0.03), zp, ws,
                             np.float32(o.scale()), int(o.zero_point()))
        assert np.array_equal(got, np.array(o.numpy())), (c, kc, k, s, p, h, w_)


def test_random_linear_geometries():
    # Linear::forward_prop(Tensor<u8>&&), fully_connected.cc:22-52, incl. 1-row / 1-feature shapes,
    # zero points at both ends of the range and inputs that saturate the requantise clamp
    m = ref.module()
    rng = np.random.default_rng(11)
    for it in range(12):
        rows, k, n = int(rng.integers(1, 40)), int(rng.integers(1, 300)), int(rng.integers(1, 70))
        wt = rng.uniform(-0.4, 0.4, size=(n, k)).astype(np.float32)
        b = rng.uniform(-0.2, 0.2, size=(n,)).astype(np.float32)
        x = rng.uniform(-4, 4, size=(max(rows, 100), k)).astype(np.float32)   # >= 1000 outputs: see SURVEY A9
        L = m.Linear(k, n)
        L.load_weight(wt)
        L.load_bias(b)
        L.prepare()
        L(m.tensor(x))
        L.convert()
        zp = int((0, 255, 127)[it % 3] if it < 6 else rng.integers(0, 256))
        q = m.quantize(m.tensor(x[:rows]), 0.04, zp)
        o = L(q)
        qw, qb, ws = port.quantize_weight(wt, b)
        got = port.linear_u8(np.array(q.numpy()), qw, qb, np.float32(0.04), zp, ws,
                             np.float32(o.scale()), int(o.zero_point()))
        assert np.array_equal(got, np.array(o.numpy())), (rows, k, n, zp)


def test_random_elementwise_and_pool():
    # quantize (unclamped, wrapping: quantize_utils.cc:44-52), dequantize (:54-58), relu<u8>
    # (functional.cc:15-26) and max_pool2d<u8> (:36-64) incl. overlapping / ragged windows
    m = ref.module()
    rng = np.random.default_rng(12)
    for _ in range(10):
        n, c = int(rng.integers(1, 4)), int(rng.integers(1, 7))
        h, w_ = int(rng.integers(3, 20)), int(rng.integers(3, 20))
        x = rng.uniform(-9, 9, size=(n, c, h, w_)).astype(np.float32)   # beyond the 0.025/127 window: wraps
        scale, zp = np.float32(rng.choice([0.025, 0.05, 0.1])), int(rng.integers(0, 256))
        q = m.quantize(m.tensor(x), float(scale), zp)
        qn = np.array(q.numpy())
        assert np.array_equal(port.quantize(x, scale, zp), qn)
        assert np.array_equal(port.dequantize(qn, scale, zp), np.array(m.dequantize(q).numpy()))
        assert np.array_equal(port.relu_u8(qn, zp), np.array(m.relu(q).numpy()))
        k = int(rng.integers(1, min(h, w_, 4) + 1))
        s = int(rng.integers(1, 4))
        assert np.array_equal(port.max_pool2d_u8(qn, k, s), np.array(m.max_pool2d(q, k, s).numpy())), (h, w_, k, s)
