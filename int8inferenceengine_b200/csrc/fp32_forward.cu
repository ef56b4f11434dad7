// FP32 forward kernels — the calibration side of the path.
//   Conv2d::forward_prop(Tensor<float>&&)  conv2d.cc:63-98   (im2col + sgemm + bias, NCHW out)
//   Linear::forward_prop(Tensor<float>&&)  fully_connected.cc:5-21
//   relu<float>                             functional.cc:5-13
//   max_pool2d<float>                       functional.cc:36-64
// These only feed the calibrator (conv2d.cc:94-96, fully_connected.cc:17-19) and the "FP32" column
// of the reference's tables, and they run once per model, so they are plain fp32 FMA kernels: the
// tensor cores have no fp32 input mode (kind::tf32 drops 13 mantissa bits, which would move the
// calibration ranges away from the reference's). One implicit-GEMM kernel serves both layer types:
// no im2col buffer, the bias add and the calibrator's min/max are fused into the epilogue, the
// output is written straight in the reference's NCHW / [M, N] order.
#include <cfloat>

#include "common.cuh"

namespace i8ie {
namespace {

constexpr int FBM = 128, FBN = 64, FBK = 16, FTHREADS = 256;
constexpr int FAP = FBM + 4, FBP = FBN + 4;   // padded smem pitches (rows stay 16-byte aligned)

struct F32Geom {
  int n, c, h, w, kc, kh, kw, stride, pad, oh, ow;
  int M, K;   // M = n*oh*ow output pixels, K = c*kh*kw
};

// float atomic min / max through the integer units (sign-split ordering trick)
__device__ __forceinline__ void atomic_min_f32(float* a, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f32(float* a, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}

struct KTap {
  int off;   // ci*h*w + ky*w + kx
  int ky, kx;
  int ok;    // k < K
};

__device__ __forceinline__ KTap make_tap(const F32Geom& g, int k) {
  KTap t;
  t.ok = k < g.K;
  const int kk = t.ok ? k : 0;
  const int khw = g.kh * g.kw;
  const int ci = kk / khw, r = kk - ci * khw;
  t.ky = r / g.kw;
  t.kx = r - t.ky * g.kw;
  t.off = (ci * g.h + t.ky) * g.w + t.kx;
  return t;
}

// C[m][n] = sum_k A[m][k] * W[n][k] + bias[n];  A gathered from the NCHW input on the fly
// (k = (ci, ky, kx) in the reference's im2col order, conv2d.cc:11,25,27; out-of-range taps are 0).
// 128 x 64 output tile per block, 8 x 4 per thread, K tiles of 16 double-buffered through registers.
__global__ void __launch_bounds__(FTHREADS) f32_igemm_kernel(const F32Geom g, const float* __restrict__ x,
                                                             const float* __restrict__ wt,
                                                             const float* __restrict__ bias, float* __restrict__ y,
                                                             float* __restrict__ minmax) {
  __shared__ __align__(16) float As[2][FBK][FAP];
  __shared__ __align__(16) float Bs[2][FBK][FBP];
  __shared__ KTap ktab[2][FBK];
  __shared__ float red[2][FTHREADS / 32];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * FBM, n0 = blockIdx.y * FBN;
  const int ohw = g.oh * g.ow;
  const int nk = (g.K + FBK - 1) / FBK;

  // loader roles. A: this thread always fetches output pixel m0 + am for k = ak, ak + 2, ...
  const int am = tid & (FBM - 1), ak = tid >> 7;
  const int mrow = m0 + am;
  const bool m_ok = mrow < g.M;
  int iy0 = 0, ix0 = 0;
  const float* xb = x;
  if (m_ok) {
    const int img = mrow / ohw, pix = mrow - img * ohw;
    const int oy = pix / g.ow, ox = pix - oy * g.ow;
    iy0 = oy * g.stride - g.pad;
    ix0 = ox * g.stride - g.pad;
    xb = x + (size_t)img * g.c * g.h * g.w + (ptrdiff_t)iy0 * g.w + ix0;
  }
  // B: weight row n0 + bn + 16*j, column k0 + bk
  const int bk = tid & (FBK - 1), bn = tid >> 4;

  if (tid < 2 * FBK) ktab[tid >> 4][tid & 15] = make_tap(g, tid);   // taps of K tiles 0 and 1
  __syncthreads();

  float areg[8], breg[4];
  auto gload = [&](int kt) {
    const int k0 = kt * FBK;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const KTap t = ktab[kt & 1][ak + 2 * j];
      const int iy = iy0 + t.ky, ix = ix0 + t.kx;
      const bool ok = m_ok && t.ok && iy >= 0 && iy < g.h && ix >= 0 && ix < g.w;
      areg[j] = ok ? __ldg(xb + t.off) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + bn + 16 * j;
      breg[j] = (n < g.kc && k0 + bk < g.K) ? __ldg(wt + (size_t)n * g.K + k0 + bk) : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[buf][ak + 2 * j][am] = areg[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[buf][bk][bn + 16 * j] = breg[j];
  };

  // compute roles: rows tm*4 .. +3 and 64 + tm*4 .. +3, columns tn*4 .. +3 (tm fastest across
  // lanes: the NCHW stores of a warp walk consecutive pixels of one channel)
  const int tm = tid & 15, tn = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
    if (tid < FBK) ktab[buf][tid] = make_tap(g, (kt + 2) * FBK + tid);   // tile kt's taps were consumed last round
#pragma unroll
    for (int k = 0; k < FBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + tm * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

  // epilogue: + bias (conv2d.cc:85-91, fully_connected.cc:12-16), NCHW / [M, N] store, min / max
  float mn = FLT_MAX, mx = -FLT_MAX;
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = n0 + tn * 4 + j;
    bv[j] = col < g.kc ? __ldg(bias + col) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i < 4 ? tm * 4 + i : 64 + tm * 4 + (i - 4));
    if (row >= g.M) continue;
    const int img = row / ohw, pix = row - img * ohw;
    float* yr = y + (size_t)img * g.kc * ohw + pix;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tn * 4 + j;
      if (col >= g.kc) continue;
      const float v = acc[i][j] + bv[j];
      yr[(size_t)col * ohw] = v;
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
  }
  if (minmax != nullptr) {   // calibrator.cc:6-27 with quantile 1: the range of everything the layer emitted
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = mn; red[1][tid >> 5] = mx; }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int i = 1; i < FTHREADS / 32; ++i) { mn = fminf(mn, red[0][i]); mx = fmaxf(mx, red[1][i]); }
      if (mn <= mx) {
        atomic_min_f32(minmax, mn);
        atomic_max_f32(minmax + 1, mx);
      }
    }
  }
}

__global__ void relu_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    y[i] = (v > 0) ? v : 0.f;   // functional.cc:10 (NaN -> 0 like the reference's comparison)
  }
}

__global__ void maxpool_f32_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int h,
                                        int w, int oh, int ow, int ks, int stride) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int p = (int)(idx % ow);
  const int o = (int)((idx / ow) % oh);
  const int64_t plane = idx / ((int64_t)ow * oh);   // n*c + channel
  const float* src = x + (plane * h + (int64_t)o * stride) * w + (int64_t)p * stride;
  float m = -FLT_MAX;   // functional.cc:28-31
  for (int a = 0; a < ks; ++a)
    for (int b = 0; b < ks; ++b) {
      const float v = src[a * w + b];
      m = (m >= v) ? m : v;   // functional.cc:55-57, same comparison (NaN handling included)
    }
  y[idx] = m;
}

int launch_f32_igemm(const F32Geom& g, const float* x, const float* w, const float* bias, float* y, float* minmax,
                     cudaStream_t stream) {
  const dim3 grid((unsigned)((g.M + FBM - 1) / FBM), (unsigned)((g.kc + FBN - 1) / FBN));
  I8IE_REQUIRE(grid.y <= 65535u, "f32 forward: too many output channels (%d)", g.kc);
  f32_igemm_kernel<<<grid, FTHREADS, 0, stream>>>(g, x, w, bias, y, minmax);
  return check_launch("f32_igemm_kernel");
}

}  // namespace
}  // namespace i8ie

using namespace i8ie;

extern "C" {

int i8ie_conv2d_f32(const float* x, const float* w, const float* bias, float* y, int n, int c, int h, int wd,
                    int kc, int kh, int kw, int stride, int pad, float* minmax2, void* stream) {
  I8IE_REQUIRE(n >= 0 && c >= 1 && h >= 1 && wd >= 1 && kc >= 1 && kh >= 1 && kw >= 1 && stride >= 1 && pad >= 0,
               "conv2d_f32: bad geometry");
  I8IE_REQUIRE(h + 2 * pad >= kh && wd + 2 * pad >= kw, "conv2d_f32: kernel larger than the padded input");
  F32Geom g{n, c, h, wd, kc, kh, kw, stride, pad, (h - kh + 2 * pad) / stride + 1, (wd - kw + 2 * pad) / stride + 1, 0, 0};
  const int64_t M = (int64_t)n * g.oh * g.ow, K = (int64_t)c * kh * kw;
  I8IE_REQUIRE(M < (1ll << 31) - FBM && K < (1ll << 31) - 3 * FBK && (int64_t)c * h * wd < (1ll << 31),
               "conv2d_f32: problem too large for 32-bit indexing");
  g.M = (int)M; g.K = (int)K;
  if (M == 0) return I8IE_OK;
  return launch_f32_igemm(g, x, w, bias, y, minmax2, (cudaStream_t)stream);
}

int i8ie_linear_f32(const float* x, const float* w, const float* bias, float* y, int m, int n, int k,
                    float* minmax2, void* stream) {
  I8IE_REQUIRE(m >= 0 && n >= 1 && k >= 1, "linear_f32: bad shape");
  I8IE_REQUIRE(m < (1ll << 31) - FBM && k < (1ll << 31) - 3 * FBK, "linear_f32: problem too large");
  if (m == 0) return I8IE_OK;
  // a linear layer is the 1x1 "convolution" of m one-pixel images with k channels
  F32Geom g{m, k, 1, 1, n, 1, 1, 1, 0, 1, 1, m, k};
  return launch_f32_igemm(g, x, w, bias, y, minmax2, (cudaStream_t)stream);
}

int i8ie_relu_f32(const float* x, float* y, int64_t n, void* stream) {
  I8IE_REQUIRE(n >= 0, "relu_f32: bad n");
  if (n == 0) return I8IE_OK;
  const int64_t want = (n + 255) / 256;
  const int grid = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
  relu_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, n);
  return check_launch("relu_f32_kernel");
}

int i8ie_maxpool_f32_nchw(const float* x, float* y, int n, int c, int h, int w, int ksize, int stride,
                          void* stream) {
  I8IE_REQUIRE(n >= 0 && c >= 1 && ksize >= 1 && stride >= 1 && h >= ksize && w >= ksize, "maxpool_f32: bad geometry");
  const int oh = (h - ksize) / stride + 1, ow = (w - ksize) / stride + 1;
  const int64_t total = (int64_t)n * c * oh * ow;
  if (total == 0) return I8IE_OK;
  I8IE_REQUIRE((total + 255) / 256 < (1ll << 31), "maxpool_f32: too many outputs");
  maxpool_f32_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, total, h, w, oh, ow,
                                                                                           ksize, stride);
  return check_launch("maxpool_f32_nchw_kernel");
}

}  // extern "C"
