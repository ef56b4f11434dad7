"""-m gpu parity of the FP32 forward kernels (Conv2d / Linear forward_prop(float),
relu<float>, max_pool2d<float>) vs the oracle's restatement of conv2d.cc:63-98,
fully_connected.cc:5-21, functional.cc:5-13,36-64.

Tolerance: the reference's own fp32 tests use atol 0.1 (unittest/test_layers.py:13-71); fp32 GEMMs
only differ by summation order, so the bar here is |err| <= 2e-5 * sum_k |a_k w_k| (a few ulps of
the magnitude that was accumulated) — written per test below. relu / max-pool are bit-exact."""
import numpy as np
import pytest
import torch

import i8ie
from int8inferenceengine_b200 import _lib, backend as B
from oracle import port

pytestmark = pytest.mark.gpu

CONV = [  # n, c, h, w, kc, k, stride, pad
    (2, 3, 32, 32, 20, 5, 1, 0),        # BASELINE config 2, conv1
    (3, 20, 28, 28, 50, 5, 1, 0),
    (2, 3, 67, 67, 96, 11, 4, 2),       # AlexNet stem geometry
    (2, 10, 50, 50, 20, 3, 7, 3),       # unittest/test_layers.py:58-71
    (5, 7, 9, 11, 130, 3, 2, 1),        # ragged: K = 63, N = 130 (3 column tiles), M = 150
    (1, 1, 5, 5, 1, 5, 1, 0),           # one output
    (130, 4, 6, 6, 8, 3, 1, 1),         # many images, small planes
]


def _bound(absx, absw, stride, pad):
    """sum_k |a_k| |w_k| per output, computed by the oracle's own conv on the absolute values."""
    return port.conv2d_f32(absx, absw, np.zeros(absw.shape[0], np.float32), stride, pad)


@pytest.mark.parametrize("geom", CONV)
def test_conv2d_f32_matches_oracle(geom):
    n, c, h, w_, kc, k, s, p = geom
    rng = np.random.default_rng(sum(geom))
    x = rng.uniform(-2.1, 2.6, size=(n, c, h, w_)).astype(np.float32)
    w = rng.uniform(-0.3, 0.3, size=(kc, c, k, k)).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, size=(kc,)).astype(np.float32)
    L = B.Conv2d(c, kc, k, s, p)
    L.load_weight(w)
    L.load_bias(b)
    before = _lib.launch_count()
    got = L(B.tensor(x)).numpy()
    assert _lib.launch_count() > before          # our kernel ran, not a library conv
    exp = port.conv2d_f32(x, w, b, s, p)
    assert got.shape == exp.shape
    tol = 2e-5 * (_bound(np.abs(x), np.abs(w), s, p) + np.abs(b)[None, :, None, None]) + 1e-30
    assert np.all(np.abs(got - exp) <= tol)


@pytest.mark.parametrize("m,k,n", [(100, 784, 10), (7, 7680, 10), (33, 300, 64), (1, 5, 3), (200, 800, 500)])
def test_linear_f32_matches_oracle(m, k, n):
    rng = np.random.default_rng(m + k + n)
    x = rng.uniform(-1, 1, size=(m, k)).astype(np.float32)
    w = rng.uniform(-0.2, 0.2, size=(n, k)).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, size=(n,)).astype(np.float32)
    L = B.Linear(k, n)
    L.load_weight(w)
    L.load_bias(b)
    got = L(B.tensor(x)).numpy()
    exp = port.linear_f32(x, w, b)
    tol = 2e-5 * (np.abs(x) @ np.abs(w).T + np.abs(b)[None, :]) + 1e-30
    assert np.all(np.abs(got - exp) <= tol)


def test_fused_calibrator_range_is_the_range_of_the_output():
    """prepare(): the forward kernel folds min / max of what it writes into the calibrator
    (conv2d.cc:94-96, fully_connected.cc:17-19) — same range as a separate pass over the output,
    accumulated over several calibration batches, negative-only and positive-only outputs included."""
    rng = np.random.default_rng(5)
    for bias_shift in (0.0, 40.0, -40.0):
        w = rng.uniform(-0.3, 0.3, size=(24, 6, 3, 3)).astype(np.float32)
        b = (rng.uniform(-0.5, 0.5, size=(24,)) + bias_shift).astype(np.float32)
        L = B.Conv2d(6, 24, 3, 1, 1)
        L.load_weight(w)
        L.load_bias(b)
        L.prepare()
        lo, hi = np.inf, -np.inf
        for i in range(3):
            x = rng.uniform(-2, 2, size=(9, 6, 20, 20)).astype(np.float32)
            y = L(B.tensor(x)).numpy()
            lo, hi = min(lo, y.min()), max(hi, y.max())
        assert L._cal.out_cnt == 3 * y.size
        assert np.float32(L._cal.mn) == lo and np.float32(L._cal.mx) == hi
    lin = B.Linear(300, 40)
    w = rng.uniform(-0.2, 0.2, size=(40, 300)).astype(np.float32)
    lin.load_weight(w)
    lin.load_bias(np.zeros(40, np.float32))
    lin.prepare()
    y = lin(B.tensor(rng.uniform(-1, 1, size=(50, 300)).astype(np.float32))).numpy()
    assert np.float32(lin._cal.mn) == y.min() and np.float32(lin._cal.mx) == y.max()


def test_relu_and_maxpool_f32_bit_exact():
    rng = np.random.default_rng(11)
    x = rng.uniform(-3, 3, size=(3, 5, 13, 17)).astype(np.float32)
    x.reshape(-1)[::97] = 0.0
    x.reshape(-1)[5::131] = -0.0
    t = B.tensor(x)
    assert np.array_equal(B.relu(t).numpy(), port.relu_f32(x))
    for k, s in [(2, 2), (3, 2), (3, 1), (5, 3), (13, 1)]:
        assert np.array_equal(B.max_pool2d(t, k, s).numpy(), port.max_pool2d_f32(x, k, s)), (k, s)
    big = rng.uniform(-1, 1, size=(1 << 21) + 3).astype(np.float32)       # grid-stride path
    assert np.array_equal(B.relu(B.tensor(big)).numpy(), port.relu_f32(big))


def test_fp32_model_forward_matches_oracle_model():
    """Whole fp32 forward (the "My Engine FP32" column / the calibration pass) of the CIFAR conv net
    and the LeNet-style net vs the oracle model, logits within fp32-GEMM tolerance, same argmax."""
    from int8inferenceengine_b200 import workloads as W
    from int8inferenceengine_b200.runner import TopologyModule
    from oracle import models
    for topo in ("simple_conv", "lenet", "mini_alex"):
        sd = W.make_weights(topo, 0)
        x = W.make_images(topo, 16, 3)
        m = TopologyModule(topo)              # not converted: fp32 path
        m.load(sd)
        got = m(i8ie.tensor(x)).numpy()
        exp = models.PortModel(topo, sd).forward_fp32(x)
        assert np.allclose(got, exp, rtol=1e-4, atol=1e-4 * np.abs(exp).max())
        assert np.array_equal(got.argmax(1), exp.argmax(1))
