"""Pins the C restatement (oracle/i8ie_oracle.c) against the golden vectors the
compiled reference produced (tests/golden/make_golden.py). CPU only."""
import numpy as np
import pytest

from int8inferenceengine_b200 import workloads as W
from oracle import models, port

from conftest import load_golden


def test_quantize_dequantize_kat():
    g = load_golden("kat_elementwise")
    for tag in ["a", "b", "edge", "c"]:
        s, z = g[f"q_{tag}_sz"]
        q = port.quantize(g[f"q_{tag}_x"], np.float32(s), int(z))
        assert np.array_equal(q, g[f"q_{tag}_q"]), tag
        d = port.dequantize(q, np.float32(s), int(z))
        assert np.array_equal(d, g[f"q_{tag}_deq"]), tag
    # the documented wrap cases: quantize(0.025,127) of
    # [-4,-3.2,-3.175,0,3.2,3.3,10] -> [223,255,0,127,255,3,15]
    assert list(g["q_edge_q"][:7]) == [223, 255, 0, 127, 255, 3, 15]


def test_relu_pool_kat():
    g = load_golden("kat_elementwise")
    q = g["fn_q"]
    assert np.array_equal(port.relu_u8(q, 127), g["fn_relu"])
    assert np.array_equal(port.max_pool2d_u8(q, 3, 2), g["fn_pool32"])
    assert np.array_equal(port.max_pool2d_u8(q, 2, 2), g["fn_pool22"])
    assert np.array_equal(port.max_pool2d_u8(q, 3, 1), g["fn_pool31"])


@pytest.mark.parametrize("tag", ["full1000", "lt1000", "all_pos", "all_neg", "all_zero", "tiny", "wide", "lt1000_neg"])
def test_get_range_kat(tag):
    g = load_golden("kat_get_range")
    s, z = port.get_range(g[f"{tag}_samples"])
    es, ez = g[f"{tag}_sz"]
    assert s == np.float32(es) and z == int(ez)
    if tag not in ("lt1000", "lt1000_neg"):
        # with a full buffer get_range(1) is exactly a min/max reduction
        a = g[f"{tag}_samples"]
        s2, z2 = port.get_range_minmax(a.min(), a.max())
        assert s2 == np.float32(es) and z2 == int(ez)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_linear_kat(tag):
    g = load_golden("kat_linear")
    w, b, x = g[f"{tag}_w"], g[f"{tag}_b"], g[f"{tag}_x"]
    qw, qb, ws = port.quantize_weight(w, b)
    os_, oz = g[f"{tag}_sz"]
    # the layer's range comes from <=1000 calibration outputs -> deterministic
    ycal = g[f"{tag}_ycal"]
    if ycal.size <= 1000:
        s, z = port.get_range(ycal)
        assert s == np.float32(os_) and z == int(oz)
    qin = port.quantize(x, 0.025, 127)
    assert np.array_equal(qin, g[f"{tag}_qin"])
    out = port.linear_u8(qin, qw, qb, np.float32(0.025), 127, ws, np.float32(os_), int(oz))
    assert np.array_equal(out, g[f"{tag}_out"])
    assert np.array_equal(port.dequantize(out, np.float32(os_), int(oz)), g[f"{tag}_deq"])


@pytest.mark.parametrize("tag", ["k5", "k11s4p2", "k5p2", "k3s7p3", "k3p1", "k3p1_c64"])
def test_conv_kat(tag):
    g = load_golden("kat_conv")
    n, c, h, w_, kc, k, s, p = [int(v) for v in g[f"{tag}_geom"]]
    qw, qb, ws = port.quantize_weight(g[f"{tag}_w"], g[f"{tag}_b"])
    os_, oz = g[f"{tag}_sz"]
    out = port.conv2d_u8(g[f"{tag}_qin"], qw, qb, s, p, np.float32(0.025), 127, ws, np.float32(os_), int(oz))
    assert np.array_equal(out, g[f"{tag}_out"])
    out2 = port.conv2d_u8(g[f"{tag}_qin2"], qw, qb, s, p, np.float32(0.031), 90, ws, np.float32(os_), int(oz))
    assert np.array_equal(out2, g[f"{tag}_out2"])


@pytest.mark.parametrize("topo", ["fc_mnist", "lenet", "simple_conv", "mini_alex"])
def test_net_golden(topo):
    g = load_golden(f"net_{topo}")
    qp = {str(n): (s, int(z)) for n, s, z in zip(g["qp_names"], g["qp_scale"], g["qp_zp"])}
    pm = models.PortModel(topo, W.make_weights(topo, 0))
    pm.convert(qp)
    logits, recs = pm.forward_int8(W.make_images(topo, int(g["batch"]), 2), record=True)
    keys = sorted(k for k in g.files if k.startswith("op"))
    assert len(keys) == len(recs)
    for k, (tag, arr, _, _) in zip(keys, recs):
        assert k.endswith(tag)
        assert np.array_equal(arr, g[k]), k
    assert np.array_equal(logits, g["logits"])


def test_alexnet_golden():
    """Full AlexNet-224, batch 2: per-op SHA-256 of the u8 activations + logits."""
    import hashlib
    g = load_golden("net_alexnet")
    qp = {str(n): (s, int(z)) for n, s, z in zip(g["qp_names"], g["qp_scale"], g["qp_zp"])}
    pm = models.PortModel("alexnet", W.make_weights("alexnet", 0))
    pm.convert(qp)
    logits, recs = pm.forward_int8(W.make_images("alexnet", 2, 2), record=True)
    assert [t for t, *_ in recs] == [str(t) for t in g["op_tags"]]
    for (tag, arr, _, _), sha in zip(recs, g["op_sha256"]):
        assert hashlib.sha256(arr.tobytes()).hexdigest() == str(sha), tag
    assert np.array_equal(logits, g["logits"])


def test_per_channel_extension_reduces_to_the_reference_when_scales_are_equal():
    """The F4 oracle extension is a composition of the pinned functions: with every channel given the
    per-tensor scale it must reproduce the reference results exactly."""
    from oracle import port
    rng = np.random.default_rng(11)
    w = rng.uniform(-0.2, 0.2, size=(12, 5, 3, 3)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(12,)).astype(np.float32)
    q = rng.integers(0, 256, size=(2, 5, 9, 9), dtype=np.uint8)
    qw, qb, ws = port.quantize_weight(w, b)
    ref = port.conv2d_u8(q, qw, qb, 1, 1, np.float32(0.03), 90, ws, np.float32(0.05), 120)
    got = port.conv2d_u8_pc(q, qw, qb, 1, 1, np.float32(0.03), 90, np.full(12, ws, np.float32), np.float32(0.05), 120)
    assert np.array_equal(ref, got)
    w2 = rng.uniform(-0.2, 0.2, size=(7, 40)).astype(np.float32)
    b2 = rng.uniform(-0.05, 0.05, size=(7,)).astype(np.float32)
    x2 = rng.integers(0, 256, size=(6, 40), dtype=np.uint8)
    qw2, qb2, ws2 = port.quantize_weight(w2, b2)
    assert np.array_equal(port.linear_u8(x2, qw2, qb2, np.float32(0.03), 90, ws2, np.float32(0.05), 120),
                          port.linear_u8_pc(x2, qw2, qb2, np.float32(0.03), 90, np.full(7, ws2, np.float32), np.float32(0.05), 120))
    # per-channel weight quantisation = layer.cc:6-26 applied to each channel on its own
    qw3, qb3, s3 = port.quantize_weight_per_channel(w2, b2)
    for j in range(7):
        a, c, sj = port.quantize_weight(w2[j], b2[j:j + 1])
        assert np.array_equal(a, qw3[j]) and c[0] == qb3[j] and sj == s3[j]
