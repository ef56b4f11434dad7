"""-m gpu: the tcgen05 (kind::i8) kernels specifically — forced impl=2 — against the oracle:
s32 accumulators and u8 outputs bit-exact, no protocol timeouts, on TMA-aligned shapes
incl. ragged M / N / K tails, zero-point border correction, stride, and every K-block width."""
import numpy as np
import pytest
import torch

from int8inferenceengine_b200 import _lib
from oracle import port

from gpu_utils import make_layer, u8_tensor_from_nchw

pytestmark = pytest.mark.gpu


def _no_tc_error():
    assert _lib.load().i8ie_debug_tc_error(1) == 0, "a tensor-core kernel hit a barrier timeout"


FC = [(128, 32, 32), (128, 128, 32), (1, 64, 16), (100, 784, 10), (100, 7680, 10), (128, 9216, 256),
      (300, 1000, 200), (130, 4096, 4096), (1000, 4096, 10), (257, 515, 129), (17, 33, 700),
      # AlexNet fc1 / fc2 at the bench batch and at small batch: split-K across CTAs + fold kernel
      (100, 9216, 4096), (100, 4096, 4096), (125, 9216, 4096), (7, 4096, 4096), (1, 9216, 4096), (33, 2000, 1000)]


@pytest.mark.parametrize("shape", FC)
def test_tc_fc(shape):
    m, k, n = shape
    rng = np.random.default_rng(m * 7 + k + n)
    a = np.sqrt(6.0 / k)
    w = rng.uniform(-a, a, size=(n, k)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(n,)).astype(np.float32)
    q = rng.integers(0, 256, size=(m, k), dtype=np.uint8)
    in_scale, in_zp, out_scale, out_zp = np.float32(0.0518), 116, np.float32(0.18), 127
    L = make_layer("fc", w, b, (out_scale, out_zp))
    qw, qb, ws = port.quantize_weight(w, b)
    exp, exp_acc = port.linear_u8(q, qw, qb, in_scale, in_zp, ws, out_scale, out_zp, want_acc=True)
    acc = torch.empty(m * n, dtype=torch.int32, device="cuda")
    out = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), acc_out=acc, impl=2)
    _no_tc_error()
    assert np.array_equal(acc.cpu().numpy().reshape(m, n), exp_acc)
    assert np.array_equal(out.numpy(), exp)
    for _ in range(3):   # back-to-back launches reuse the split-K scratch
        out = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), impl=2)
    _no_tc_error()
    assert np.array_equal(out.numpy(), exp)
    # pad lanes of the padded output rows carry the zero point
    ldy = (n + 15) // 16 * 16
    raw = out.buf.cpu().numpy().reshape(m, ldy)
    assert np.all(raw[:, n:] == out_zp)


CONV = [  # n, c, h, w, kc, k, stride, pad
    (2, 32, 8, 8, 32, 1, 1, 0), (2, 32, 8, 8, 32, 3, 1, 0), (2, 32, 8, 8, 32, 3, 1, 1),
    (3, 64, 13, 13, 96, 3, 1, 1), (2, 96, 27, 27, 256, 5, 1, 2), (2, 256, 13, 13, 384, 3, 1, 1),
    (2, 384, 13, 13, 256, 3, 1, 1), (2, 128, 9, 11, 64, 3, 2, 1), (1, 32, 20, 20, 40, 5, 3, 2),
    (5, 50, 12, 12, 120, 5, 1, 0), (4, 20, 28, 28, 50, 5, 1, 0), (1, 160, 5, 5, 10, 5, 1, 2),
    (7, 64, 6, 6, 300, 3, 1, 3),
]


@pytest.mark.parametrize("geom", CONV)
def test_tc_conv(geom):
    n, c, h, w_, kc, k, s, p = geom
    rng = np.random.default_rng(sum(geom))
    a = np.sqrt(6.0 / (c * k * k))
    w = rng.uniform(-a, a, size=(kc, c, k, k)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(kc,)).astype(np.float32)
    q = rng.integers(0, 256, size=(n, c, h, w_), dtype=np.uint8)
    in_scale, in_zp = np.float32(0.0293), int(rng.integers(1, 256))
    out_scale, out_zp = np.float32(0.061), int(rng.integers(60, 190))
    L = make_layer("conv", w, b, (out_scale, out_zp), s, p)
    qw, qb, ws = port.quantize_weight(w, b)
    exp, exp_acc = port.conv2d_u8(q, qw, qb, s, p, in_scale, in_zp, ws, out_scale, out_zp, want_acc=True)
    oh, ow = exp.shape[2], exp.shape[3]
    acc = torch.empty(n * oh * ow * kc, dtype=torch.int32, device="cuda")
    out = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), acc_out=acc, impl=2)
    _no_tc_error()
    assert np.array_equal(acc.cpu().numpy().reshape(n, oh * ow, kc), exp_acc)
    assert np.array_equal(out.numpy(), exp)
    L.fuse_relu = True
    out_r = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), impl=2)
    assert np.array_equal(out_r.numpy(), port.relu_u8(exp, out_zp))


PAIR = [  # shapes served by the CTA-pair kernel (cp % 128 == 0, N tile 128 / 192 / 256): n, c, h, w, kc, k, stride, pad
    (1, 128, 5, 5, 128, 3, 1, 1),        # one M tile: the odd CTA of the pair only lends its weight half
    (3, 128, 13, 13, 128, 3, 2, 0),      # stride 2, no padding, N tile 128
    (5, 256, 13, 13, 384, 3, 1, 1),      # 7 M tiles (odd) x 2 N tiles of 192
    (9, 96, 27, 27, 256, 5, 1, 2),       # 52 M tiles, pad 2, pitch-128 input with 96 real channels
    (2, 128, 16, 16, 500, 1, 1, 0),      # 1x1, N = 500 -> two N tiles of 256 with a ragged tail
    (4, 384, 7, 9, 192, 3, 1, 2),        # pad 2 with a 3x3 filter, 3 channel blocks
]


def _run_conv_parity(geom, impl=2):
    n, c, h, w_, kc, k, s, p = geom
    rng = np.random.default_rng(sum(geom) + 1)
    a = np.sqrt(6.0 / (c * k * k))
    w = rng.uniform(-a, a, size=(kc, c, k, k)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(kc,)).astype(np.float32)
    q = rng.integers(0, 256, size=(n, c, h, w_), dtype=np.uint8)
    in_scale, in_zp = np.float32(0.0293), int(rng.integers(1, 256))
    out_scale, out_zp = np.float32(0.061), int(rng.integers(60, 190))
    L = make_layer("conv", w, b, (out_scale, out_zp), s, p)
    qw, qb, ws = port.quantize_weight(w, b)
    exp, exp_acc = port.conv2d_u8(q, qw, qb, s, p, in_scale, in_zp, ws, out_scale, out_zp, want_acc=True)
    oh, ow = exp.shape[2], exp.shape[3]
    acc = torch.empty(n * oh * ow * kc, dtype=torch.int32, device="cuda")
    out = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), acc_out=acc, impl=impl)
    _no_tc_error()
    assert L._last_impl == impl
    assert np.array_equal(acc.cpu().numpy().reshape(n, oh * ow, kc), exp_acc)
    assert np.array_equal(out.numpy(), exp)
    raw = out.buf.cpu().numpy().reshape(n * oh * ow, -1)
    assert np.all(raw[:, kc:] == out_zp)          # pad lanes of the output pitch carry the zero point
    L.fuse_relu = True
    out_r = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), impl=impl)
    assert np.array_equal(out_r.numpy(), port.relu_u8(exp, out_zp))


SMALL_M = [  # AlexNet conv2..conv5 at batch 1 .. 16 (BASELINE config 5's small end): one to a dozen pair tiles
    (1, 256, 13, 13, 384, 3, 1, 1), (1, 384, 13, 13, 384, 3, 1, 1), (4, 384, 13, 13, 256, 3, 1, 1),
    (16, 384, 13, 13, 256, 3, 1, 1), (8, 256, 13, 13, 384, 3, 1, 1),
]


@pytest.mark.parametrize("geom", SMALL_M)
def test_tc_conv_small_m(geom):
    """Small-M convolutions (a handful of tiles, most of the chip idle): accumulators, bytes, pad lanes and relu
    against the oracle."""
    _run_conv_parity(geom)


@pytest.mark.parametrize("geom", [(1, 96, 27, 27, 256, 5, 1, 2), (3, 96, 27, 27, 256, 5, 1, 2)])
def test_tc_conv_small_m_row_mode(geom):
    """The same for a row-mode plan (AlexNet conv2 at batch 1 / 3: physically padded input, no border table)."""
    _run_conv_parity(geom, impl=4)


@pytest.mark.parametrize("geom", PAIR)
def test_tc_conv_cta_pair(geom):
    """tcgen05.mma.cta_group::2 kernel (two SMs per 256-row tile)."""
    _run_conv_parity(geom)


@pytest.mark.parametrize("geom", [PAIR[2], PAIR[3], PAIR[5]])
def test_tc_conv_single_cta_variant(geom, monkeypatch):
    """The same layers through the single-CTA kernel (I8IE_NO_CLUSTER=1) stay bit-exact."""
    monkeypatch.setenv("I8IE_NO_CLUSTER", "1")
    _run_conv_parity(geom)


PAIR_WRAP = [  # more pair tiles than the 74 clusters of the persistent grid, odd tails: every cluster runs several
    # tiles, so the TMEM accumulator ring wraps, tmem_full / tmem_empty flip phase and the TMA ring runs ahead
    # across tile boundaries — all checked against the oracle (s32 accumulators and u8)
    (40, 96, 27, 27, 256, 5, 1, 2),      # AlexNet conv2: 228 M tiles -> 114 pair tiles, 1 N tile
    (130, 256, 13, 13, 384, 3, 1, 1),    # AlexNet conv3: 172 M tiles -> 86 pair tiles x 2 N tiles of 192
    (125, 384, 13, 13, 256, 3, 1, 1),    # AlexNet conv5 at the 8-GPU shard size: 166 M tiles (83 pairs, last one ragged)
    (61, 384, 13, 13, 384, 3, 1, 1),     # AlexNet conv4: 81 M tiles (odd) -> 41 pair tiles x 2, tail pair half empty
    (140, 128, 17, 19, 320, 3, 2, 1),    # stride 2, N = 320 at pitch 384 -> 50 pair tiles x 2 N tiles of 192, ragged N tail
]


@pytest.mark.parametrize("geom", PAIR_WRAP)
def test_tc_conv_cta_pair_many_tiles(geom):
    _run_conv_parity(geom)


@pytest.mark.parametrize("bn", ["128", "192", "256"])
@pytest.mark.parametrize("geom", [(100, 256, 13, 13, 384, 3, 1, 1),     # AlexNet conv3 at the bench batch (67 pair tiles)
                                  (130, 384, 13, 13, 384, 3, 1, 1)])    # conv4, 86 pair tiles
def test_tc_conv_pair_every_n_tile_width(geom, bn, monkeypatch):
    """Every N tile width of the pair kernel stays bit-exact (I8IE_TC_BN overrides the plan's pick)."""
    monkeypatch.setenv("I8IE_TC_BN", bn)
    _run_conv_parity(geom)


ROW = [  # stride-1 convs whose channel count is not a multiple of 128: row mode (physically padded input,
    # K = filter row x contiguous kw * cp run): n, c, h, w, kc, k, stride, pad
    (9, 96, 27, 27, 256, 5, 1, 2),       # AlexNet conv2: 5 x 512 bytes of K instead of 25 x 128
    (40, 96, 27, 27, 256, 5, 1, 2),      # ... with more pair tiles than clusters
    (3, 48, 14, 14, 128, 3, 1, 1),       # run of 144 bytes -> KR 256
    (2, 160, 9, 9, 256, 3, 1, 0),        # no padding, run 480 -> 512
    (5, 20, 28, 28, 130, 5, 1, 3),       # pad 3, C = 20 -> pitch 32, run 160 -> 256, ragged N
    (2, 200, 11, 7, 384, 2, 1, 1),       # even filter, non-square image, N = 384
    (7, 3, 32, 32, 20, 5, 1, 0),         # narrow layers on the single-CTA kernel: simple-conv conv1 (pitch 16, no padding:
    (7, 20, 28, 28, 50, 5, 1, 0),        #   the NHWC tensor itself is the operand), conv2 (pitch 32), conv3 (pitch 64)
    (7, 50, 12, 12, 120, 5, 1, 0),
    (3, 1, 28, 28, 20, 5, 1, 2),         # one channel, padded
]


@pytest.mark.parametrize("geom", ROW)
def test_tc_conv_row_mode(geom):
    _run_conv_parity(geom, impl=4)


def test_row_mode_is_auto_selected_and_fed_by_the_pool():
    """Auto dispatch takes row mode for AlexNet conv2, and a pending max-pool writes the physically padded
    operand directly (no NHWC round trip): pool -> conv equals the oracle's pool -> conv."""
    from int8inferenceengine_b200 import backend as B
    rng = np.random.default_rng(5)
    n, c, h, w_ = 6, 96, 55, 55
    q = rng.integers(0, 256, size=(n, c, h, w_), dtype=np.uint8)
    a = np.sqrt(6.0 / (c * 25))
    w = rng.uniform(-a, a, size=(256, c, 5, 5)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(256,)).astype(np.float32)
    in_scale, in_zp, out_scale, out_zp = np.float32(0.0518), 116, np.float32(0.0819), 105
    L = make_layer("conv", w, b, (out_scale, out_zp), 1, 2)
    x = u8_tensor_from_nchw(q, in_scale, in_zp)
    pooled = B.max_pool2d(x, 3, 2)
    y = L(pooled)
    qw, qb, ws = port.quantize_weight(w, b)
    exp = port.conv2d_u8(port.max_pool2d_u8(q, 3, 2), qw, qb, 1, 2, in_scale, in_zp, ws, out_scale, out_zp)
    assert np.array_equal(y.numpy(), exp)
    _no_tc_error()
    assert L._last_impl == 4
    assert pooled._pending("pool") is not None      # the pool itself was never launched in its NHWC form
    assert np.array_equal(pooled.numpy(), port.max_pool2d_u8(q, 3, 2))


def test_tc_ineligible_shapes_are_refused_when_forced():
    rng = np.random.default_rng(0)
    w = rng.uniform(-0.3, 0.3, size=(8, 3, 3, 3)).astype(np.float32)   # cp = 16: not a 32-byte K block
    L = make_layer("conv", w, np.zeros(8, np.float32), (np.float32(0.05), 128), 1, 1)
    q = rng.integers(0, 256, size=(1, 3, 8, 8), dtype=np.uint8)
    with pytest.raises(RuntimeError):
        L._forward_u8(u8_tensor_from_nchw(q, 0.03, 77), impl=2)
    # ... and taken by the SIMT kernel under auto dispatch
    L._forward_u8(u8_tensor_from_nchw(q, 0.03, 77), impl=0)


STEM = [  # n, c, h, w, kc, k, stride, pad  (c <= 4, stride 4 or 8)
    (2, 3, 67, 67, 32, 11, 4, 2), (2, 3, 224, 224, 96, 11, 4, 2), (2, 1, 40, 36, 16, 7, 4, 3),
    (1, 4, 50, 50, 24, 8, 8, 0), (3, 2, 33, 31, 10, 5, 4, 1), (130, 3, 19, 19, 8, 11, 4, 2),
]


@pytest.mark.parametrize("geom", STEM)
def test_tc_stem_conv(geom):
    """Small-C strided first-layer conv (AlexNet conv1) on the tcgen05 stem path, plus the
    fused input-quantise entry point (i8ie_conv2d_f32_u8)."""
    from int8inferenceengine_b200 import backend as B
    n, c, h, w_, kc, k, s, p = geom
    rng = np.random.default_rng(sum(geom))
    a = np.sqrt(6.0 / (c * k * k))
    w = rng.uniform(-a, a, size=(kc, c, k, k)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(kc,)).astype(np.float32)
    x = rng.uniform(-3.1, 3.1, size=(n, c, h, w_)).astype(np.float32)
    in_scale, in_zp = np.float32(0.025), 127
    out_scale, out_zp = np.float32(0.0518), 116
    q = port.quantize(x, in_scale, in_zp)
    L = make_layer("conv", w, b, (out_scale, out_zp), s, p)
    qw, qb, ws = port.quantize_weight(w, b)
    exp, exp_acc = port.conv2d_u8(q, qw, qb, s, p, in_scale, in_zp, ws, out_scale, out_zp, want_acc=True)
    oh, ow = exp.shape[2], exp.shape[3]
    acc = torch.empty(n * oh * ow * kc, dtype=torch.int32, device="cuda")
    out = L._forward_u8(u8_tensor_from_nchw(q, in_scale, in_zp), acc_out=acc, impl=2)
    _no_tc_error()
    assert L._last_impl == 3
    assert np.array_equal(acc.cpu().numpy().reshape(n, oh * ow, kc), exp_acc)
    assert np.array_equal(out.numpy(), exp)
    # fused quantise + conv straight from the fp32 image
    acc.zero_()
    out2 = L.forward_quantize_fused(B.tensor(x), in_scale, in_zp, acc_out=acc)
    _no_tc_error()
    assert out2 is not None
    assert np.array_equal(acc.cpu().numpy().reshape(n, oh * ow, kc), exp_acc)
    assert np.array_equal(out2.numpy(), exp)
    # a different input zero point exercises the physical zp border
    q3 = port.quantize(x, np.float32(0.031), 90)
    exp3 = port.conv2d_u8(q3, qw, qb, s, p, np.float32(0.031), 90, ws, out_scale, out_zp)
    out3 = L._forward_u8(u8_tensor_from_nchw(q3, 0.031, 90), impl=0)
    assert np.array_equal(out3.numpy(), exp3)


STEM_FQ = [  # RGB stride-4 stems whose rows are whole 16-byte units: the stem kernel quantises the fp32 image itself
    (3, 3, 64, 64, 64, 11, 4, 2), (5, 3, 100, 96, 96, 11, 4, 2), (2, 3, 52, 48, 32, 7, 4, 2),
    (150, 3, 32, 32, 16, 11, 4, 2),     # 600 tiles: both shared-memory rings wrap several times per SM
    (2, 3, 224, 224, 96, 11, 4, 0),     # no padding
]


@pytest.mark.parametrize("geom", STEM_FQ)
@pytest.mark.parametrize("wild", [False, True])
def test_tc_stem_fused_quantize(geom, wild):
    """i8ie_conv2d_f32_u8 with the quantise fused into the stem kernel (fp32 rows bulk-copied to shared
    memory, quantised by converter warps): same bits as quantize() followed by the u8 conv, including
    out-of-range / non-finite pixels (x86 cast semantics of quantize_utils.cc:44-52)."""
    from int8inferenceengine_b200 import backend as B
    n, c, h, w_, kc, k, s, p = geom
    rng = np.random.default_rng(sum(geom) + 7)
    a = np.sqrt(6.0 / (c * k * k))
    w = rng.uniform(-a, a, size=(kc, c, k, k)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(kc,)).astype(np.float32)
    x = rng.uniform(-3.1, 3.1, size=(n, c, h, w_)).astype(np.float32)
    if wild:
        flat = x.reshape(-1)
        idx = rng.choice(flat.size, size=min(4000, flat.size // 8), replace=False)
        flat[idx] = rng.choice(np.array([np.nan, np.inf, -np.inf, 1e30, -1e30, 7.5, -9.25, 3.0e9, -4.0e9, 1e-30, -0.0],
                                        np.float32), size=idx.size)
    in_scale, in_zp = np.float32(0.025), 127
    out_scale, out_zp = np.float32(0.0518), 116
    q = port.quantize(x, in_scale, in_zp)
    L = make_layer("conv", w, b, (out_scale, out_zp), s, p)
    qw, qb, ws = port.quantize_weight(w, b)
    exp, exp_acc = port.conv2d_u8(q, qw, qb, s, p, in_scale, in_zp, ws, out_scale, out_zp, want_acc=True)
    oh, ow = exp.shape[2], exp.shape[3]
    acc = torch.zeros(n * oh * ow * kc, dtype=torch.int32, device="cuda")
    out = L.forward_quantize_fused(B.tensor(x), in_scale, in_zp, acc_out=acc)
    _no_tc_error()
    assert out is not None
    assert np.array_equal(acc.cpu().numpy().reshape(n, oh * ow, kc), exp_acc)
    assert np.array_equal(out.numpy(), exp)
    raw = out.buf.cpu().numpy().reshape(n * oh * ow, -1)
    assert np.all(raw[:, kc:] == out_zp)
