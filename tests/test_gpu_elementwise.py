"""-m gpu parity: HBM-bound kernels vs the oracle / golden vectors, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import port

from conftest import load_golden
from gpu_utils import dev, lib, stream

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 15, 16, 17, 1000, 4096 + 3, 1 << 20, (1 << 20) + 5]


def _check(rc):
    assert rc == 0, lib().i8ie_last_error()


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("scale,zp", [(0.025, 127), (0.0371, 100), (1.0, 0)])
def test_quantize_flat(n, scale, zp):
    rng = np.random.default_rng(n + zp)
    x = rng.uniform(-4, 4, size=n).astype(np.float32)   # includes wrap-around values
    q = torch.empty(n, dtype=torch.uint8, device="cuda")
    _check(lib().i8ie_quantize_f32_u8(dev(x).data_ptr(), q.data_ptr(), n, scale, zp, stream()))
    assert np.array_equal(q.cpu().numpy(), port.quantize(x, np.float32(scale), zp))


def test_quantize_extreme_values_match_x86_cast():
    """inf / NaN / beyond-int32 / denormal inputs: the reference's cast is cvttss2si + low byte."""
    specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 3e38, -3e38, 1e30, -1e30, 1e18, -1e18, 9.9e17,
                         2.0**31 * 0.025, 2.0**31 * 0.025 * 1.0001, -2.0**31 * 0.025, 5.3e7, -5.3e7, 5.4e7,
                         1e-45, -1e-45, 1e-38, -1e-38, 1e-30, 6.374, 6.375, 6.376, -3.175, -3.2, 3.2],
                        dtype=np.float32)
    rng = np.random.default_rng(7)
    big = (rng.standard_normal(4096) * 10.0 ** rng.uniform(-40, 38, size=4096)).astype(np.float32)
    x = np.concatenate([specials, big])
    q = torch.empty(x.size, dtype=torch.uint8, device="cuda")
    for scale, zp in [(0.025, 127), (1.0, 0), (3.7e-5, 255), (812.5, 3), (1e-19, 10), (2e18, 10)]:
        _check(lib().i8ie_quantize_f32_u8(dev(x).data_ptr(), q.data_ptr(), x.size, scale, zp, stream()))
        assert np.array_equal(q.cpu().numpy(), port.quantize(x, np.float32(scale), zp)), (scale, zp)


def test_quantize_rounding_boundaries():
    """x just below / at / above k*scale for every code k: the truncation boundaries of x/scale + zp,
    where a quotient that is off by one ulp would change the result."""
    k = np.arange(-300, 600, dtype=np.float64)
    for scale, zp in [(0.025, 127), (0.0371, 100), (0.1, 0), (1.0 / 3.0, 17), (0.0078125, 128), (1.7, 250)]:
        s32 = np.float32(scale)
        centre = (k * float(s32)).astype(np.float32)
        xs = [centre]
        for _ in range(3):
            xs.append(np.nextafter(xs[-1], np.float32(np.inf)))
        lo = centre
        for _ in range(3):
            lo = np.nextafter(lo, np.float32(-np.inf))
            xs.append(lo)
        x = np.concatenate(xs).astype(np.float32)
        q = torch.empty(x.size, dtype=torch.uint8, device="cuda")
        _check(lib().i8ie_quantize_f32_u8(dev(x).data_ptr(), q.data_ptr(), x.size, float(s32), zp, stream()))
        assert np.array_equal(q.cpu().numpy(), port.quantize(x, s32, zp)), (scale, zp)


def test_quantize_dense_float_sweep():
    """Every float32 in a band of consecutive bit patterns (2^22 of them) for the default input qparams."""
    bits = np.arange(0x3F000000, 0x3F000000 + (1 << 22), dtype=np.uint32)   # [0.5, 1.0)
    for sign in (0, 0x80000000):
        x = (bits | np.uint32(sign)).view(np.float32)
        q = torch.empty(x.size, dtype=torch.uint8, device="cuda")
        _check(lib().i8ie_quantize_f32_u8(dev(x).data_ptr(), q.data_ptr(), x.size, 0.025, 127, stream()))
        assert np.array_equal(q.cpu().numpy(), port.quantize(x, np.float32(0.025), 127))


def test_quantize_misaligned_and_golden():
    g = load_golden("kat_elementwise")
    for tag in ["a", "b", "edge", "c"]:
        s, z = g[f"q_{tag}_sz"]
        x = g[f"q_{tag}_x"].ravel()
        buf = torch.zeros(x.size + 3, dtype=torch.float32, device="cuda")
        buf[1:1 + x.size] = dev(x)                     # 4-byte aligned only
        q = torch.empty(x.size + 16, dtype=torch.uint8, device="cuda")
        _check(lib().i8ie_quantize_f32_u8(buf.data_ptr() + 4, q.data_ptr() + 1, x.size, float(s), int(z), stream()))
        assert np.array_equal(q.cpu().numpy()[1:1 + x.size], g[f"q_{tag}_q"].ravel()), tag


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (2, 3, 5, 7), (3, 1, 28, 28), (2, 20, 9, 11), (1, 96, 4, 4), (2, 17, 3, 3)])
def test_quantize_nchw_to_nhwc(shape):
    n, c, h, w = shape
    rng = np.random.default_rng(c)
    x = rng.uniform(-3.1, 3.1, size=shape).astype(np.float32)
    cp = (c + 15) // 16 * 16
    q = torch.empty(n * h * w * cp, dtype=torch.uint8, device="cuda")
    _check(lib().i8ie_quantize_nchw_f32_nhwc_u8(dev(x).data_ptr(), q.data_ptr(), n, c, h, w, cp, 0.025, 127, stream()))
    got = q.cpu().numpy().reshape(n, h, w, cp)
    exp = port.quantize(x, np.float32(0.025), 127).transpose(0, 2, 3, 1)
    assert np.array_equal(got[..., :c], exp)
    assert np.all(got[..., c:] == 127)


@pytest.mark.parametrize("n", SIZES)
def test_dequantize_flat(n):
    rng = np.random.default_rng(n)
    q = rng.integers(0, 256, size=n, dtype=np.uint8)
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    _check(lib().i8ie_dequantize_u8_f32(dev(q).data_ptr(), x.data_ptr(), n, 0.0417, 93, stream()))
    assert np.array_equal(x.cpu().numpy(), port.dequantize(q, np.float32(0.0417), 93))


def test_dequantize_all_codes_all_zero_points():
    q = np.tile(np.arange(256, dtype=np.uint8), 5)[: 256 * 5 - 3]
    x = torch.empty(q.size, dtype=torch.float32, device="cuda")
    for zp in [0, 1, 93, 127, 128, 254, 255]:
        for scale in [0.0417, 1.0, 3.3e-7, 1234.5]:
            _check(lib().i8ie_dequantize_u8_f32(dev(q).data_ptr(), x.data_ptr(), q.size, scale, zp, stream()))
            assert np.array_equal(x.cpu().numpy(), port.dequantize(q, np.float32(scale), zp)), (zp, scale)


def test_dequantize_misaligned():
    rng = np.random.default_rng(11)
    q = rng.integers(0, 256, size=1000 + 1, dtype=np.uint8)
    qd = dev(q)
    x = torch.empty(1000 + 1, dtype=torch.float32, device="cuda")
    _check(lib().i8ie_dequantize_u8_f32(qd.data_ptr() + 1, x.data_ptr() + 4, 1000, 0.2, 127, stream()))
    assert np.array_equal(x.cpu().numpy()[1:], port.dequantize(q[1:], np.float32(0.2), 127))


def test_dequantize_rows():
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, size=(37, 48), dtype=np.uint8)
    x = torch.empty(37 * 10, dtype=torch.float32, device="cuda")
    _check(lib().i8ie_dequantize_rows_u8_f32(dev(q).data_ptr(), x.data_ptr(), 37, 10, 48, 0.2, 127, stream()))
    assert np.array_equal(x.cpu().numpy().reshape(37, 10), port.dequantize(q[:, :10], np.float32(0.2), 127))


@pytest.mark.parametrize("n", SIZES)
def test_downscale(n):
    rng = np.random.default_rng(n)
    acc = rng.integers(-200000, 200000, size=n, dtype=np.int32)
    if n > 8:
        acc[:8] = [0, 1, -1, 2**31 - 1, -2**31, 2**24 + 1, -2**24 - 1, 123456789]
    y = torch.empty(n, dtype=torch.uint8, device="cuda")
    sa, sb, sc, zp = np.float32(0.025), np.float32(0.0031), np.float32(0.0518), 116
    _check(lib().i8ie_downscale_s32_u8(dev(acc).data_ptr(), y.data_ptr(), n, sa, sb, sc, zp, stream()))
    assert np.array_equal(y.cpu().numpy(), port.down_scale(acc, sa, sb, sc, zp))


def test_downscale_exhaustive_rounding_boundaries():
    """Every accumulator in a dense window around the clamp / truncation boundaries."""
    acc = np.arange(-70000, 70000, dtype=np.int32)
    y = torch.empty(acc.size, dtype=torch.uint8, device="cuda")
    for sa, sb, sc, zp in [(0.025, 0.00312, 0.0518, 116), (0.0819, 0.0021, 0.148, 120), (0.2, 0.001, 0.136, 107),
                           (0.025, 0.0123, 1.0, 0)]:
        sa, sb, sc = np.float32(sa), np.float32(sb), np.float32(sc)
        _check(lib().i8ie_downscale_s32_u8(dev(acc).data_ptr(), y.data_ptr(), acc.size, sa, sb, sc, zp, stream()))
        assert np.array_equal(y.cpu().numpy(), port.down_scale(acc, sa, sb, sc, zp))


@pytest.mark.parametrize("n", [1, 15, 16, 17, 1000, 4099, 1 << 20, (1 << 22) + 7])
def test_minmax(n):
    rng = np.random.default_rng(n)
    x = rng.normal(0, 3, size=n).astype(np.float32)
    ws = torch.zeros(int(lib().i8ie_minmax_workspace_bytes()), dtype=torch.uint8, device="cuda")
    out = torch.empty(2, dtype=torch.float32, device="cuda")
    xd = dev(x)
    for _ in range(2):  # the workspace must be reusable without re-zeroing
        _check(lib().i8ie_minmax_f32(xd.data_ptr(), n, out.data_ptr(), ws.data_ptr(), stream()))
        mn, mx = out.cpu().numpy()
        assert mn == x.min() and mx == x.max()


@pytest.mark.parametrize("n", SIZES)
def test_relu(n):
    rng = np.random.default_rng(n)
    q = rng.integers(0, 256, size=n, dtype=np.uint8)
    y = torch.empty(n, dtype=torch.uint8, device="cuda")
    _check(lib().i8ie_relu_u8(dev(q).data_ptr(), y.data_ptr(), n, 131, stream()))
    assert np.array_equal(y.cpu().numpy(), port.relu_u8(q, 131))


@pytest.mark.parametrize("shape,k,s", [((2, 3, 13, 13), 3, 2), ((2, 3, 13, 13), 2, 2), ((2, 3, 13, 13), 3, 1),
                                       ((1, 96, 55, 55), 3, 2), ((3, 50, 24, 24), 2, 2), ((2, 20, 7, 9), 7, 1)])
@pytest.mark.parametrize("out_nchw", [0, 1])
def test_maxpool(shape, k, s, out_nchw):
    n, c, h, w = shape
    rng = np.random.default_rng(c + k)
    q = rng.integers(0, 256, size=shape, dtype=np.uint8)
    cp = (c + 15) // 16 * 16
    xin = np.full((n, h, w, cp), 7, np.uint8)
    xin[..., :c] = q.transpose(0, 2, 3, 1)
    oh, ow = (h - k) // s + 1, (w - k) // s + 1
    y = torch.zeros(n * oh * ow * (c if out_nchw else cp), dtype=torch.uint8, device="cuda")
    _check(lib().i8ie_maxpool_u8_nhwc(dev(xin).data_ptr(), y.data_ptr(), n, h, w, c, cp, k, s, out_nchw, stream()))
    exp = port.max_pool2d_u8(q, k, s)
    if out_nchw:
        assert np.array_equal(y.cpu().numpy().reshape(n, c, oh, ow), exp)
    else:
        got = y.cpu().numpy().reshape(n, oh, ow, cp)
        assert np.array_equal(got[..., :c], exp.transpose(0, 2, 3, 1))
        assert np.all(got[..., c:] == 7)      # pad lanes keep the producer's pad value (its zero point)


def test_relu_pool_golden():
    g = load_golden("kat_elementwise")
    import i8ie
    t = i8ie.quantize(i8ie.tensor(g["fn_x"]), 0.025, 127)
    assert np.array_equal(t.numpy(), g["fn_q"])
    assert np.array_equal(i8ie.relu(t).numpy(), g["fn_relu"])
    assert np.array_equal(i8ie.max_pool2d(t, 3, 2).numpy(), g["fn_pool32"])
    assert np.array_equal(i8ie.max_pool2d(t, 2, 2).numpy(), g["fn_pool22"])
    assert np.array_equal(i8ie.max_pool2d(t, 3, 1).numpy(), g["fn_pool31"])


def test_layout_roundtrip():
    rng = np.random.default_rng(9)
    for shape in [(2, 3, 5, 7), (1, 20, 4, 4), (3, 33, 2, 9)]:
        n, c, h, w = shape
        cp = (c + 15) // 16 * 16
        q = rng.integers(0, 256, size=shape, dtype=np.uint8)
        a = torch.empty(n * h * w * cp, dtype=torch.uint8, device="cuda")
        _check(lib().i8ie_u8_nchw_to_nhwc(dev(q).data_ptr(), a.data_ptr(), n, c, h, w, cp, 55, stream()))
        got = a.cpu().numpy().reshape(n, h, w, cp)
        assert np.array_equal(got[..., :c], q.transpose(0, 2, 3, 1)) and np.all(got[..., c:] == 55)
        b = torch.empty(q.size, dtype=torch.uint8, device="cuda")
        _check(lib().i8ie_u8_nhwc_to_nchw(a.data_ptr(), b.data_ptr(), n, c, h, w, cp, stream()))
        assert np.array_equal(b.cpu().numpy().reshape(shape), q)


@pytest.mark.parametrize("n,k,zp", [(10, 784, 127), (37, 75, 0), (4096, 9216, 255), (5, 1, 200), (64, 4096, 131)])
def test_zp_offsets(n, k, zp):
    rng = np.random.default_rng(n + k)
    qw = rng.integers(-128, 128, size=(n, k), dtype=np.int8)
    if n == 4096:
        qw[:7] = 127      # zp * sum|w| >= 2^24: exercises the sequential-fp32 fallback
        qw[7:9] = -128
    qb = rng.integers(-128, 128, size=(n,), dtype=np.int8)
    oc = torch.empty(n, dtype=torch.int32, device="cuda")
    bf = torch.empty(n, dtype=torch.float32, device="cuda")
    qw_d, qb_d = dev(qw), dev(qb)   # keep both alive: temporaries would alias in the caching allocator
    for is_conv in (1, 0):
        _check(lib().i8ie_zp_offsets(qw_d.data_ptr(), qb_d.data_ptr(), n, k, zp, 0.025, is_conv,
                                     oc.data_ptr(), bf.data_ptr(), stream()))
        if is_conv:
            exp = port.conv_offsets(qw, qb, zp, np.float32(0.025))
            assert np.all(bf.cpu().numpy() == 0)
        else:
            exp = port.fc_offsets(qw, zp)
            assert np.array_equal(bf.cpu().numpy(), qb.astype(np.float32) / np.float32(0.025))
        assert np.array_equal(oc.cpu().numpy(), exp)
