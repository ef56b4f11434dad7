"""-m gpu parity: conv2d / fully-connected through the C ABI vs the oracle and the golden
vectors of the compiled reference. Bit-exact: u8 outputs AND s32 accumulators."""
import numpy as np
import pytest
import torch

from int8inferenceengine_b200 import backend as B
from oracle import port

from conftest import load_golden
from gpu_utils import make_layer, u8_tensor_from_nchw

pytestmark = pytest.mark.gpu

IMPLS = [1, 0]   # 1 = forced SIMT dp4a kernel, 0 = auto (tcgen05 where eligible)


def run_conv(L, q, in_scale, in_zp, impl, want_acc=True):
    n, c, h, w = q.shape
    kc, _, kh, kw = L._qw_shape
    oh = (h - kh + 2 * L._pad) // L._stride + 1
    ow = (w - kw + 2 * L._pad) // L._stride + 1
    acc = torch.empty(n * oh * ow * kc, dtype=torch.int32, device="cuda") if want_acc else None
    out = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), acc_out=acc, impl=impl)
    return out.numpy(), (acc.cpu().numpy().reshape(n, oh * ow, kc) if want_acc else None)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("tag", ["k5", "k11s4p2", "k5p2", "k3s7p3", "k3p1", "k3p1_c64"])
def test_conv_golden(tag, impl):
    g = load_golden("kat_conv")
    n, c, h, w_, kc, k, s, p = [int(v) for v in g[f"{tag}_geom"]]
    os_, oz = g[f"{tag}_sz"]
    L = make_layer("conv", g[f"{tag}_w"], g[f"{tag}_b"], (os_, int(oz)), s, p)
    out, _ = run_conv(L, g[f"{tag}_qin"], 0.025, 127, impl)
    assert np.array_equal(out, g[f"{tag}_out"])
    out2, _ = run_conv(L, g[f"{tag}_qin2"], 0.031, 90, impl)
    assert np.array_equal(out2, g[f"{tag}_out2"])


GEOMS = [  # n, c, h, w, kc, k, stride, pad
    (1, 1, 5, 5, 1, 1, 1, 0), (2, 3, 32, 32, 20, 5, 1, 0), (2, 20, 28, 28, 50, 5, 1, 0),
    (3, 50, 12, 12, 120, 5, 1, 0), (2, 3, 67, 67, 32, 11, 4, 2), (2, 32, 7, 7, 64, 5, 1, 2),
    (2, 64, 13, 13, 96, 3, 1, 1), (1, 96, 27, 27, 256, 5, 1, 2), (2, 256, 13, 13, 384, 3, 1, 1),
    (5, 16, 9, 8, 24, 3, 2, 1), (2, 17, 6, 10, 33, 4, 3, 2), (1, 128, 8, 8, 128, 3, 1, 1),
    (130, 8, 3, 3, 8, 3, 1, 1),
]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("geom", GEOMS)
def test_conv_vs_oracle(geom, impl):
    n, c, h, w_, kc, k, s, p = geom
    rng = np.random.default_rng(sum(geom))
    a = np.sqrt(6.0 / (c * k * k))
    w = rng.uniform(-a, a, size=(kc, c, k, k)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(kc,)).astype(np.float32)
    q = rng.integers(0, 256, size=(n, c, h, w_), dtype=np.uint8)
    in_scale, in_zp = np.float32(0.0293), int(rng.integers(0, 256))
    out_scale, out_zp = np.float32(0.061), int(rng.integers(60, 190))
    L = make_layer("conv", w, b, (out_scale, out_zp), s, p)
    qw, qb, ws = port.quantize_weight(w, b)
    assert np.array_equal(L.q_weight().numpy(), qw) and np.array_equal(L.q_bias().numpy(), qb)
    exp, exp_acc = port.conv2d_u8(q, qw, qb, s, p, in_scale, in_zp, ws, out_scale, out_zp, want_acc=True)
    out, acc = run_conv(L, q, in_scale, in_zp, impl)
    assert np.array_equal(acc, exp_acc), "s32 accumulators differ"
    assert np.array_equal(out, exp), "requantised u8 differ"
    # fused relu epilogue == separate relu<u8>
    L.fuse_relu = True
    out_r, _ = run_conv(L, q, in_scale, in_zp, impl, want_acc=False)
    assert np.array_equal(out_r, port.relu_u8(exp, out_zp))


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_linear_golden(tag, impl):
    g = load_golden("kat_linear")
    os_, oz = g[f"{tag}_sz"]
    L = make_layer("fc", g[f"{tag}_w"], g[f"{tag}_b"], (os_, int(oz)))
    x = u8_tensor_from_nchw(g[f"{tag}_qin"], 0.025, 127)
    out = L._forward_u8(x, impl=impl)
    assert np.array_equal(out.numpy(), g[f"{tag}_out"])
    assert np.array_equal(B.dequantize(out).numpy(), g[f"{tag}_deq"])


FC_SHAPES = [(1, 1, 1), (100, 784, 10), (7, 300, 24), (3, 65, 130), (100, 7680, 10), (16, 9216, 512),
             (130, 4096, 4096), (257, 512, 300), (5, 4096, 10), (64, 800, 500)]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("shape", FC_SHAPES)
def test_linear_vs_oracle(shape, impl):
    m, k, n = shape
    rng = np.random.default_rng(m + k + n)
    a = np.sqrt(6.0 / k)
    w = rng.uniform(-a, a, size=(n, k)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(n,)).astype(np.float32)
    q = rng.integers(0, 256, size=(m, k), dtype=np.uint8)
    in_scale, in_zp = np.float32(0.0518), int(rng.integers(0, 256))
    out_scale, out_zp = np.float32(0.18), int(rng.integers(60, 190))
    L = make_layer("fc", w, b, (out_scale, out_zp))
    qw, qb, ws = port.quantize_weight(w, b)
    exp, exp_acc = port.linear_u8(q, qw, qb, in_scale, in_zp, ws, out_scale, out_zp, want_acc=True)
    acc = torch.empty(m * n, dtype=torch.int32, device="cuda")
    out = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), acc_out=acc, impl=impl)
    assert np.array_equal(acc.cpu().numpy().reshape(m, n), exp_acc), "s32 accumulators differ"
    assert np.array_equal(out.numpy(), exp)
    L.fuse_relu = True
    out_r = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), impl=impl)
    assert np.array_equal(out_r.numpy(), port.relu_u8(exp, out_zp))


@pytest.mark.parametrize("shape", [(100, 4096, 10), (5, 784, 10), (1, 64, 16), (7, 300, 24), (130, 512, 300)])
@pytest.mark.parametrize("relu", [False, True])
def test_linear_with_fused_dequantize(shape, relu):
    """i8ie_fc_u8_deq: the last fc + Module.__call__'s dequantize (module.py:22-24) through one entry
    point — classifier heads (<= 16 outputs) in one kernel, other shapes fc + dequantise. Both the u8
    result and the fp32 logits against the oracle (fully_connected.cc:22-52, quantize_utils.cc:38-42)."""
    m, k, n = shape
    rng = np.random.default_rng(7 * m + k + n)
    a = np.sqrt(6.0 / k)
    w = rng.uniform(-a, a, size=(n, k)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(n,)).astype(np.float32)
    q = rng.integers(0, 256, size=(m, k), dtype=np.uint8)
    in_scale, in_zp = np.float32(0.0518), int(rng.integers(0, 256))
    out_scale, out_zp = np.float32(0.18), int(rng.integers(60, 190))
    L = make_layer("fc", w, b, (out_scale, out_zp))
    qw, qb, ws = port.quantize_weight(w, b)
    exp = port.linear_u8(q, qw, qb, in_scale, in_zp, ws, out_scale, out_zp)
    if relu:
        exp = port.relu_u8(exp, out_zp)
    exp_f = port.dequantize(exp, out_scale, out_zp)
    y, f32 = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), relu=relu, deq=True)
    assert np.array_equal(y.numpy(), exp)
    assert np.array_equal(f32.cpu().numpy().reshape(m, n), exp_f)
    # the deferred-launch route Module.__call__ takes: layer call -> [relu] -> dequantize
    t = L(u8_tensor_from_nchw(q, in_scale, in_zp))
    if relu:
        t = B.relu(t)
    assert t._pending("layer") is not None
    before = B._lib.launch_count()
    got = B.dequantize(t)
    launches = B._lib.launch_count() - before
    assert np.array_equal(got.numpy(), exp_f)
    assert np.array_equal(t.numpy(), exp)
    if n <= 16 and k % 16 == 0:
        assert launches == 1, "a classifier head and its dequantise are one kernel"


@pytest.mark.parametrize("m", [1, 100, 128, 300])
@pytest.mark.parametrize("nk", [(256, 1152), (384, 4096), (1024, 640)])
def test_linear_tiled_weight_stream(m, nk):
    """fc whose weights have the tiled, pre-swizzled copy attached (i8ie_fc_weight_tiled_attach): the tcgen05
    kernel reads contiguous 16 KB blocks instead of row-strided boxes — with and without split-K, one and
    several M tiles, 128- and 256-wide N tiles — accumulators and bytes against the oracle."""
    n, k = nk
    rng = np.random.default_rng(m + 3 * n + k)
    a = np.sqrt(6.0 / k)
    w = rng.uniform(-a, a, size=(n, k)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(n,)).astype(np.float32)
    q = rng.integers(0, 256, size=(m, k), dtype=np.uint8)
    in_scale, in_zp = np.float32(0.0518), int(rng.integers(0, 256))
    out_scale, out_zp = np.float32(0.21), int(rng.integers(60, 190))
    L = make_layer("fc", w, b, (out_scale, out_zp))
    assert L._w_tiled is not None
    qw, qb, ws = port.quantize_weight(w, b)
    exp, exp_acc = port.linear_u8(q, qw, qb, in_scale, in_zp, ws, out_scale, out_zp, want_acc=True)
    acc = torch.empty(m * n, dtype=torch.int32, device="cuda")
    out = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), acc_out=acc, impl=2)
    assert np.array_equal(acc.cpu().numpy().reshape(m, n), exp_acc), "s32 accumulators differ"
    assert np.array_equal(out.numpy(), exp)
    L._drop_tiled()                      # same layer through the tiled tensor map
    out2 = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), impl=2)
    assert np.array_equal(out2.numpy(), exp)


@pytest.mark.parametrize("shape", [(100, 1024, 4096), (64, 520, 3200), (128, 4096, 4096), (97, 2304, 3210)])
def test_linear_cluster_split_k(shape):
    """Wide small-batch fc (one M tile of >= 64 rows, enough N tiles to fill the chip): K is split over a 4-CTA
    cluster and the partial sums are folded through distributed shared memory (tc_fc_cluster_kernel). K tails,
    N that is not a multiple of the tile, relu, the s32 accumulators and the bytes against the oracle."""
    m, k, n = shape
    rng = np.random.default_rng(m + k + n)
    a = np.sqrt(6.0 / k)
    w = rng.uniform(-a, a, size=(n, k)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(n,)).astype(np.float32)
    q = rng.integers(0, 256, size=(m, k), dtype=np.uint8)
    in_scale, in_zp = np.float32(0.0518), int(rng.integers(0, 256))
    out_scale, out_zp = np.float32(0.25), int(rng.integers(60, 190))
    L = make_layer("fc", w, b, (out_scale, out_zp))
    qw, qb, ws = port.quantize_weight(w, b)
    exp, exp_acc = port.linear_u8(q, qw, qb, in_scale, in_zp, ws, out_scale, out_zp, want_acc=True)
    acc = torch.empty(m * n, dtype=torch.int32, device="cuda")
    out = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), acc_out=acc, impl=2)
    assert np.array_equal(acc.cpu().numpy().reshape(m, n), exp_acc), "s32 accumulators differ"
    assert np.array_equal(out.numpy(), exp)
    out_r = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), impl=2, relu=True)
    assert np.array_equal(out_r.numpy(), port.relu_u8(exp, out_zp))


def test_fc_bias_float_roundtrip_above_2_24():
    """fully_connected.cc:44 adds the bias in fp32 on the s32 accumulator: bits above 2^24 are
    lost exactly as in the reference. Large K with saturated operands reaches that range."""
    m, k, n = 4, 9216, 16
    w = np.full((n, k), 1.0, np.float32)
    w[::2] *= -1
    w[0, :5] = 0.013
    b = np.linspace(-1, 1, n).astype(np.float32)
    q = np.full((m, k), 255, np.uint8)
    q[1] = 200
    L = make_layer("fc", w, b, (np.float32(3000.0), 128))
    qw, qb, ws = port.quantize_weight(w, b)
    exp, exp_acc = port.linear_u8(q, qw, qb, np.float32(0.025), 3, ws, np.float32(3000.0), 128, want_acc=True)
    assert np.abs(exp_acc).max() > 2 ** 24
    acc = torch.empty(m * n, dtype=torch.int32, device="cuda")
    out = L._forward_u8(u8_tensor_from_nchw(q, 0.025, 3), acc_out=acc)
    assert np.array_equal(acc.cpu().numpy().reshape(m, n), exp_acc)
    assert np.array_equal(out.numpy(), exp)


# ---- F4 extension: per-output-channel weight scales (opt-in; the default stays the reference's) ----------
PC_CONV = [  # n, c, h, w, kc, k, stride, pad -> one geometry per kernel family
    (2, 3, 20, 20, 20, 5, 1, 0),          # SIMT dp4a (cp = 16)
    (2, 32, 8, 8, 40, 3, 1, 1),           # single-CTA tcgen05 (BK 32)
    (3, 256, 13, 13, 384, 3, 1, 1),       # CTA-pair tcgen05
    (3, 96, 27, 27, 256, 5, 1, 2),        # row mode
    (2, 3, 67, 67, 32, 11, 4, 2),         # stem
]


@pytest.mark.parametrize("geom", PC_CONV)
def test_conv_per_channel_scales(geom):
    import torch
    from int8inferenceengine_b200 import backend as B
    n, c, h, w_, kc, k, s, p = geom
    rng = np.random.default_rng(sum(geom) + 3)
    a = np.sqrt(6.0 / (c * k * k))
    # channels of very different magnitude: where per-channel scales matter
    w = (rng.uniform(-a, a, size=(kc, c, k, k)) * rng.uniform(0.05, 1.0, size=(kc, 1, 1, 1))).astype(np.float32)
    b = rng.uniform(-0.02, 0.02, size=(kc,)).astype(np.float32)
    q = rng.integers(0, 256, size=(n, c, h, w_), dtype=np.uint8)
    in_scale, in_zp, out_scale, out_zp = np.float32(0.025), 127, np.float32(0.03), 100
    L = B.Conv2d(c, kc, k, s, p)
    L.load_weight(w)
    L.load_bias(b)
    L.set_qparams(out_scale, out_zp)
    L.per_channel = True
    L.convert()
    qw, qb, ws = port.quantize_weight_per_channel(w, b)
    assert np.array_equal(L.q_weight().numpy(), qw) and np.array_equal(L.q_bias().numpy(), qb)
    assert np.array_equal(L.weight_scales(), ws)
    exp, exp_acc = port.conv2d_u8_pc(q, qw, qb, s, p, in_scale, in_zp, ws, out_scale, out_zp, want_acc=True)
    oh, ow = exp.shape[2], exp.shape[3]
    acc = torch.empty(n * oh * ow * kc, dtype=torch.int32, device="cuda")
    out = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), acc_out=acc)
    assert np.array_equal(acc.cpu().numpy().reshape(n, oh * ow, kc), exp_acc)
    assert np.array_equal(out.numpy(), exp)
    out_r = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), relu=True)
    assert np.array_equal(out_r.numpy(), port.relu_u8(exp, out_zp))
    # the scales matter: the per-tensor result differs
    assert not np.array_equal(exp, port.conv2d_u8(q, *port.quantize_weight(w, b)[:2], s, p, in_scale, in_zp,
                                                   port.quantize_weight(w, b)[2], out_scale, out_zp))


@pytest.mark.parametrize("shape", [(5, 4096, 10), (100, 9216, 512), (130, 300, 200), (3, 20, 7)])
def test_linear_per_channel_scales(shape):
    import torch
    from int8inferenceengine_b200 import backend as B
    m, k, n = shape
    rng = np.random.default_rng(m + k + n)
    a = np.sqrt(6.0 / k)
    w = (rng.uniform(-a, a, size=(n, k)) * rng.uniform(0.05, 1.0, size=(n, 1))).astype(np.float32)
    b = rng.uniform(-0.02, 0.02, size=(n,)).astype(np.float32)
    q = rng.integers(0, 256, size=(m, k), dtype=np.uint8)
    in_scale, in_zp, out_scale, out_zp = np.float32(0.0518), 116, np.float32(0.09), 127
    L = B.Linear(k, n)
    L.load_weight(w)
    L.load_bias(b)
    L.set_qparams(out_scale, out_zp)
    L.per_channel = True
    L.convert()
    qw, qb, ws = port.quantize_weight_per_channel(w, b)
    exp, exp_acc = port.linear_u8_pc(q, qw, qb, in_scale, in_zp, ws, out_scale, out_zp, want_acc=True)
    acc = torch.empty(m * n, dtype=torch.int32, device="cuda")
    out = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), acc_out=acc)
    assert np.array_equal(acc.cpu().numpy().reshape(m, n), exp_acc)
    assert np.array_equal(out.numpy(), exp)


def test_per_channel_net_runs_and_default_is_untouched():
    """A whole topology with per-channel scales agrees with the composed oracle; the same topology without
    the option still equals the reference golden logits (covered by test_gpu_nets)."""
    import i8ie
    from int8inferenceengine_b200 import workloads as W
    from int8inferenceengine_b200.runner import build_module
    topo = "mini_alex"
    sd = W.make_weights(topo, 0)
    m0 = build_module(topo, sd, calib=W.make_images(topo, 100, 1))
    qp = {nm: (np.float32(L.layer._scale), int(L.layer._zp)) for nm, L in m0.layers().items()}
    m = build_module(topo, sd, qparams=qp, per_channel=True)
    x = W.make_images(topo, 9, 2)
    got = m(i8ie.tensor(x)).numpy()
    # composed oracle forward
    qx = port.quantize(x, W.INPUT_SCALE, W.INPUT_ZP)
    scale, zp = np.float32(W.INPUT_SCALE), W.INPUT_ZP
    for op in W.TOPOLOGIES[topo]["ops"]:
        if op[0] == "conv":
            qw, qb, ws = port.quantize_weight_per_channel(sd[op[1] + ".weight"], sd[op[1] + ".bias"])
            qx = port.conv2d_u8_pc(qx, qw, qb, op[5], op[6], scale, zp, ws, *qp[op[1]])
            scale, zp = qp[op[1]]
        elif op[0] == "fc":
            qw, qb, ws = port.quantize_weight_per_channel(sd[op[1] + ".weight"], sd[op[1] + ".bias"])
            qx = port.linear_u8_pc(qx, qw, qb, scale, zp, ws, *qp[op[1]])
            scale, zp = qp[op[1]]
        elif op[0] == "relu":
            qx = port.relu_u8(qx, zp)
        elif op[0] == "pool":
            qx = port.max_pool2d_u8(qx, op[1], op[2])
        else:
            qx = qx.reshape(-1, op[1])
    assert np.array_equal(got, port.dequantize(qx, scale, zp))
