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
