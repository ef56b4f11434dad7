/* i8ie parity oracle — a plain-C CPU restatement of the reference's INT8 hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (int8inferenceengine_b200) never does and fails loudly without its CUDA library.
 *
 * Parity status: PINNED. Every function here is checked (tests/test_oracle_*.py)
 *   (1) against golden vectors produced by the reference's own src/*.cc compiled
 *       unmodified in this container (oracle/_ref, recipe: oracle/Makefile;
 *       generator: tests/golden/make_golden.py), committed under tests/golden/;
 *   (2) live against oracle/_ref whenever that build is present.
 * The reference itself ships no golden vectors / known-answer tests
 * (SURVEY.md §4, §8c), so (1)+(2) are the pin.
 *
 * Arithmetic contract: IEEE binary32, round-to-nearest-even, no FMA contraction
 * (built with -ffp-contract=off, no -ffast-math — the reference is built with
 * plain -O3, CMakeLists.txt:15), IEEE division, truncation toward zero on every
 * float->integer conversion. Out-of-range float->u8/s8 casts follow the observed
 * x86-64/gcc behaviour of the reference build: (T)(int32_t)trunc(v) (wraps mod 256).
 *
 * Citations are file:line in /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* (u8)(float) / (s8)(float) as the reference's x86-64 build performs them:
 * cvttss2si to a 32-bit int, then the low byte. */
static inline uint8_t f2u8_wrap(float v) { return (uint8_t)(int32_t)v; }
static inline int8_t f2s8_wrap(float v) { return (int8_t)(int32_t)v; }

/* A1 — quantize(Tensor<float>&, scale, zp): quantize_utils.cc:44-52.
 * out[i] = in[i] / scale + zp, stored to u8 with NO clamp (line 49). */
ORC_API void orc_quantize_f32_u8(const float* x, uint8_t* q, int64_t n, float scale, int zp) {
  const uint8_t zpb = (uint8_t)zp;
  for (int64_t i = 0; i < n; ++i) q[i] = f2u8_wrap(x[i] / scale + (float)zpb);
}

/* A7 — clamped element-wise variants, quantize_utils.cc:3-10 and :12-19
 * (no caller in the reference; restated for completeness). */
ORC_API void orc_quantize_f32_u8_clamped(const float* x, uint8_t* q, int64_t n, float scale, int zp) {
  const uint8_t zpb = (uint8_t)zp;
  for (int64_t i = 0; i < n; ++i) {
    float t = x[i] / scale + (float)zpb;
    q[i] = (t >= 255) ? 255 : (t < 0) ? 0 : f2u8_wrap(t);
  }
}
ORC_API void orc_quantize_f32_s8_clamped(const float* x, int8_t* q, int64_t n, float scale) {
  for (int64_t i = 0; i < n; ++i) {
    float t = x[i] / scale;
    q[i] = (t <= -127) ? -127 : (t >= 127) ? 127 : f2s8_wrap(t);
  }
}

/* A5 — dequantize(float*, u8*, size, scale, zp): quantize_utils.cc:38-42.
 * (Q[i] - zp) is an int subtraction, then one fp32 multiply. */
ORC_API void orc_dequantize_u8_f32(const uint8_t* q, float* x, int64_t n, float scale, int zp) {
  const uint8_t zpb = (uint8_t)zp;
  for (int64_t i = 0; i < n; ++i) x[i] = (float)((int)q[i] - (int)zpb) * scale;
}

/* A6 — dequantize(float*, int*, size, sa, sb): quantize_utils.cc:21-25 (no caller). */
ORC_API void orc_dequantize_s32_f32(const int32_t* acc, float* x, int64_t n, float sa, float sb) {
  for (int64_t i = 0; i < n; ++i) x[i] = ((float)acc[i] * sa) * sb;
}

/* A4 — down_scale: quantize_utils.cc:27-36. The requantise every layer ends with. */
static inline uint8_t down_scale_one(int32_t acc, float sa, float sb, float sc, uint8_t zp_c) {
  float dequant = ((float)acc * sa) * sb;      /* line 30: Q[i] * sa * sb, left to right */
  float quant = dequant / sc + (float)zp_c;    /* line 31 */
  return (quant >= 255) ? 255 : (quant < 0) ? 0 : f2u8_wrap(quant); /* lines 32-33 */
}
ORC_API void orc_down_scale(const int32_t* acc, uint8_t* out, int64_t n, float sa, float sb,
                            float sc, int zp_c) {
  for (int64_t i = 0; i < n; ++i) out[i] = down_scale_one(acc[i], sa, sb, sc, (uint8_t)zp_c);
}

/* A8 — quantize_weight: layer.cc:6-26. One shared min/max over weight U bias,
 * scale = (max-min)/127 for both, truncating UNclamped casts. Returns the scale. */
ORC_API float orc_quantize_weight(const float* w, int64_t nw, const float* b, int64_t nb,
                                  int8_t* qw, int8_t* qb) {
  float mx = -3.402823466e+38f, mn = 3.402823466e+38f;
  for (int64_t i = 0; i < nw; ++i) { mn = (w[i] < mn) ? w[i] : mn; mx = (mx < w[i]) ? w[i] : mx; }
  for (int64_t i = 0; i < nb; ++i) { mn = (b[i] < mn) ? b[i] : mn; mx = (mx < b[i]) ? b[i] : mx; }
  const float scale = (mx - mn) / 127;          /* lines 18-19 */
  for (int64_t i = 0; i < nw; ++i) qw[i] = f2s8_wrap(w[i] / scale);  /* line 21 */
  for (int64_t i = 0; i < nb; ++i) qb[i] = f2s8_wrap(b[i] / scale);  /* line 24 */
  return scale;
}

/* A2b — conv zero-point/bias offsets: conv2d.cc:117-124.
 * t accumulates (float)(zp * w) sequentially in the weight row's memory order
 * (c, kh, kw); oc[j] = (int)(q_bias[j] / in_scale - t). */
ORC_API void orc_conv_offsets(const int8_t* qw, const int8_t* qb, int kc, int K, int in_zp,
                              float in_scale, int32_t* oc) {
  const int zp = (uint8_t)in_zp;
  for (int j = 0; j < kc; ++j) {
    float t = 0;
    for (int k = 0; k < K; ++k) t += (float)(zp * (int)qw[(size_t)j * K + k]);
    oc[j] = (int32_t)((float)(int)qb[j] / in_scale - t);
  }
}

/* A3 (first part) — FC offsets: fully_connected.cc:30-38. oc[i] = (int)(-t). */
ORC_API void orc_fc_offsets(const int8_t* qw, int n, int K, int in_zp, int32_t* oc) {
  const int zp = (uint8_t)in_zp;
  for (int i = 0; i < n; ++i) {
    float t = 0;
    for (int j = 0; j < K; ++j) t += (float)(zp * (int)qw[(size_t)i * K + j]);
    oc[i] = (int32_t)(-t);
  }
}

/* A2a — im2col with zero-point padding: conv2d.cc:17-49.
 * Row (ti*ow+tj) of M holds the K = c*kh*kw taps of output pixel (ti,tj) in
 * (c, kh, kw) order; taps outside the image are written as zero_point (:24-25).
 * (The reference takes the unpadded branch when padding==0 (:41-46); with no
 * out-of-range tap the two branches write the same bytes.) */
static void im2col_u8(uint8_t* M, const uint8_t* I, int c, int h, int w, int kh, int kw,
                      int stride, int padding, int oh, int ow, uint8_t zero_point) {
  const int K = c * kh * kw;
  for (int ti = 0; ti < oh; ++ti) {
    const int i = ti * stride - padding;
    for (int tj = 0; tj < ow; ++tj) {
      const int j = tj * stride - padding;
      uint8_t* row = M + (size_t)(ti * ow + tj) * K;
      for (int k = 0; k < c; ++k)
        for (int l = 0; l < kh; ++l)
          for (int m = 0; m < kw; ++m) {
            const int y = i + l, x = j + m;
            row[k * kh * kw + l * kw + m] =
                (y < 0 || x < 0 || y >= h || x >= w) ? zero_point
                                                     : I[(size_t)(h * w) * k + (size_t)y * w + x];
          }
    }
  }
}

/* The one third-party call on the path — cblas_gemm_s8u8s32 as the reference
 * issues it (conv2d.cc:131-133, fully_connected.cc:39-41): RowMajor, A u8 [m,k]
 * NoTrans, B s8 [n,k] Trans, alpha=1, beta=0, ao=bo=0, CblasRowOffset with
 * co[n]. Intel MKL 2019.5 (CMakeLists.txt:25) is not under /root/reference; its
 * published semantics for this call are C = A*B^T + co broadcast over rows in
 * non-saturating s32, which this loop restates exactly. */
static void gemm_u8s8s32(const uint8_t* A, const int8_t* B, int32_t* C, int m, int n, int k,
                         const int32_t* co) {
  for (int i = 0; i < m; ++i) {
    const uint8_t* a = A + (size_t)i * k;
    for (int j = 0; j < n; ++j) {
      const int8_t* b = B + (size_t)j * k;
      int32_t s = 0;
      for (int p = 0; p < k; ++p) s += (int32_t)a[p] * (int32_t)b[p];
      C[(size_t)i * n + j] = s + co[j];
    }
  }
}

/* A2 — Conv2d::forward_prop(Tensor<u8>&&): conv2d.cc:100-142.
 * in  u8 NCHW [n,c,h,w]; qw s8 OIHW [kc,c,kh,kw]; qb s8 [kc];
 * out u8 NCHW [n,kc,oh,ow]; acc_out (optional, may be NULL) receives the s32
 * GEMM result incl. oc as [n, oh*ow, kc] (the reference never exposes it). */
ORC_API void orc_conv2d_u8(const uint8_t* in, int n, int c, int h, int w, const int8_t* qw,
                           const int8_t* qb, int kc, int kh, int kw, int stride, int padding,
                           float in_scale, int in_zp, float w_scale, float out_scale,
                           int out_zp, uint8_t* out, int32_t* acc_out) {
  const int oh = (h - kh + 2 * padding) / stride + 1;   /* :108 */
  const int ow = (w - kw + 2 * padding) / stride + 1;   /* :109 */
  const int mat_m = oh * ow, mat_n = kc, mat_k = c * kh * kw;  /* :114-116 */
  int32_t* oc = (int32_t*)malloc(sizeof(int32_t) * (size_t)mat_n);
  orc_conv_offsets(qw, qb, kc, mat_k, in_zp, in_scale, oc);   /* :117-124 */
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < n; ++i) {                                /* :125-139 */
    uint8_t* matricize = (uint8_t*)malloc((size_t)mat_m * mat_k);
    int32_t* C = (int32_t*)malloc(sizeof(int32_t) * (size_t)mat_m * mat_n);
    im2col_u8(matricize, in + (size_t)i * c * h * w, c, h, w, kh, kw, stride, padding, oh, ow,
              (uint8_t)in_zp);                                 /* :129-130 */
    gemm_u8s8s32(matricize, qw, C, mat_m, mat_n, mat_k, oc);  /* :131-133 */
    if (acc_out) memcpy(acc_out + (size_t)i * mat_m * mat_n, C, sizeof(int32_t) * (size_t)mat_m * mat_n);
    uint8_t* o = out + (size_t)i * kc * oh * ow;
    for (int p = 0; p < mat_m; ++p)                            /* :134-136 down_scale + transpose */
      for (int j = 0; j < mat_n; ++j)
        o[(size_t)j * mat_m + p] =
            down_scale_one(C[(size_t)p * mat_n + j], in_scale, w_scale, out_scale, (uint8_t)out_zp);
    free(C);
    free(matricize);
  }
  free(oc);
}

/* A3 — Linear::forward_prop(Tensor<u8>&&): fully_connected.cc:22-52.
 * in u8 [m,k]; qw s8 [n,k]; qb s8 [n]; out u8 [m,n]; acc_out (optional) gets
 * the s32 values AFTER the float bias add of lines 42-46. */
ORC_API void orc_linear_u8(const uint8_t* in, int m, int k, const int8_t* qw, const int8_t* qb,
                           int n, float in_scale, int in_zp, float w_scale, float out_scale,
                           int out_zp, uint8_t* out, int32_t* acc_out) {
  int32_t* C = (int32_t*)malloc(sizeof(int32_t) * (size_t)m * n);
  int32_t* oc = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  orc_fc_offsets(qw, n, k, in_zp, oc);                         /* :30-38 */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < m; ++i)                                  /* :39-41, row-parallel (exact) */
    gemm_u8s8s32(in + (size_t)i * k, qw, C + (size_t)i * n, 1, n, k, oc);
  for (int i = 0; i < m; ++i)                                  /* :42-46  int += float */
    for (int j = 0; j < n; ++j)
      C[(size_t)i * n + j] = (int32_t)((float)C[(size_t)i * n + j] + (float)(int)qb[j] / in_scale);
  if (acc_out) memcpy(acc_out, C, sizeof(int32_t) * (size_t)m * n);
  orc_down_scale(C, out, (int64_t)m * n, in_scale, w_scale, out_scale, out_zp);  /* :47-48 */
  free(C);
  free(oc);
}

/* A10 — relu<u8_t>: functional.cc:15-26. max(x, zero_point). */
ORC_API void orc_relu_u8(const uint8_t* in, uint8_t* out, int64_t n, int zp) {
  const uint8_t z = (uint8_t)zp;
  for (int64_t i = 0; i < n; ++i) out[i] = (in[i] > z) ? in[i] : z;
}

/* A11 — max_pool2d<u8_t>: functional.cc:36-64. No padding, floor output size. */
ORC_API void orc_max_pool2d_u8(const uint8_t* in, int n, int c, int h, int w, int ksize, int stride,
                               uint8_t* out) {
  const int oh = (h - ksize) / stride + 1, ow = (w - ksize) / stride + 1;  /* :40-41 */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < c; ++j) {
      const uint8_t* p = in + ((size_t)i * c + j) * h * w;
      uint8_t* o = out + ((size_t)i * c + j) * oh * ow;
      for (int y = 0; y < oh; ++y)
        for (int x = 0; x < ow; ++x) {
          uint8_t mx = 0;                                       /* min<u8_t>() :33-35 */
          for (int a = 0; a < ksize; ++a)
            for (int b = 0; b < ksize; ++b) {
              const uint8_t v = p[(size_t)(y * stride + a) * w + (x * stride + b)];
              mx = (mx >= v) ? mx : v;                          /* :53-54 */
            }
          o[(size_t)y * ow + x] = mx;
        }
    }
}

/* A9 — Calibrator::get_range(quantile = 1): calibrator.cc:24-37, from the state
 * Calibrator::sample (calibrator.cc:6-23) leaves behind: a 1000-slot buffer that
 * make_unique<Calibrator>() zero-initialises (layer.cc:33), whose first `cnt`
 * slots hold samples. `samples`/`cnt` here are that filled prefix, cnt in
 * [1,1000]. (Beyond 1000 values the reference replaces random slots from an
 * unseeded mt19937 — not reproducible, see orc_get_range_minmax.) */
static int cmp_float(const void* a, const void* b) {
  const float x = *(const float*)a, y = *(const float*)b;
  return (x < y) ? -1 : (x > y) ? 1 : 0;
}
static void range_from_min_max(float out_min, float out_max, float* scale, uint8_t* zero_point) {
  out_min = fminf(out_min, 0.f);                                /* :28 */
  out_max = fmaxf(out_max, 0.f);                                /* :29 */
  /* :30 — 255 * (0 - min) is float; (max - min) is float, + 1e-09 promotes to double */
  const double q = (double)(255 * (0 - out_min)) / ((double)(out_max - out_min) + 1e-09);
  const uint8_t zp = (uint8_t)(int32_t)q;
  float s = (zp == 0) ? (out_max - out_min) / 255 : (0 - out_min) / zp;  /* :31-32 */
  if (s == 0) s = 1;                                            /* :33-35 */
  *scale = s;
  *zero_point = zp;
}
ORC_API int orc_get_range(const float* samples, int64_t cnt, float* scale, uint8_t* zero_point) {
  if (cnt < 1 || cnt > 1000) return -1;
  float buf[1000];
  memset(buf, 0, sizeof(buf));
  memcpy(buf, samples, sizeof(float) * (size_t)cnt);
  qsort(buf, 1000, sizeof(float), cmp_float);                   /* :25 sorts ALL 1000 slots */
  range_from_min_max(buf[0], buf[cnt - 1], scale, zero_point);  /* :26-27 with quantile 1 */
  return 0;
}
/* The scalar arithmetic of calibrator.cc:28-35 with (min, max) supplied — what
 * the B200 min/max calibrator feeds. */
ORC_API void orc_get_range_minmax(float mn, float mx, float* scale, uint8_t* zero_point) {
  range_from_min_max(mn, mx, scale, zero_point);
}

/* ---- FP32 side (tolerance-only, feeds calibration) ---------- */

/* Conv2d::forward_prop(Tensor<float>&&): conv2d.cc:63-98 (im2col + sgemm + bias). */
ORC_API void orc_conv2d_f32(const float* in, int n, int c, int h, int w, const float* wt,
                            const float* bias, int kc, int kh, int kw, int stride, int padding,
                            float* out) {
  const int oh = (h - kh + 2 * padding) / stride + 1, ow = (w - kw + 2 * padding) / stride + 1;
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < kc; ++j)
      for (int y = 0; y < oh; ++y)
        for (int x = 0; x < ow; ++x) {
          float s = 0.f;
          for (int k = 0; k < c; ++k)
            for (int l = 0; l < kh; ++l)
              for (int m = 0; m < kw; ++m) {
                const int yy = y * stride - padding + l, xx = x * stride - padding + m;
                const float v = (yy < 0 || xx < 0 || yy >= h || xx >= w)
                                    ? 0.f
                                    : in[(((size_t)i * c + k) * h + yy) * w + xx];
                s += v * wt[(((size_t)j * c + k) * kh + l) * kw + m];
              }
          out[(((size_t)i * kc + j) * oh + y) * ow + x] = s + bias[j];
        }
}

/* Linear::forward_prop(Tensor<float>&&): fully_connected.cc:5-21. */
ORC_API void orc_linear_f32(const float* in, int m, int k, const float* wt, const float* bias,
                            int n, float* out) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      float s = 0.f;
      for (int p = 0; p < k; ++p) s += in[(size_t)i * k + p] * wt[(size_t)j * k + p];
      out[(size_t)i * n + j] = s + bias[j];
    }
}

/* relu<float>: functional.cc:5-13. */
ORC_API void orc_relu_f32(const float* in, float* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = (in[i] > 0) ? in[i] : 0;
}

/* max_pool2d<float>: functional.cc:36-64 (init value -FLT_MAX, :28-31). */
ORC_API void orc_max_pool2d_f32(const float* in, int n, int c, int h, int w, int ksize, int stride,
                                float* out) {
  const int oh = (h - ksize) / stride + 1, ow = (w - ksize) / stride + 1;
  for (int i = 0; i < n * c; ++i) {
    const float* p = in + (size_t)i * h * w;
    float* o = out + (size_t)i * oh * ow;
    for (int y = 0; y < oh; ++y)
      for (int x = 0; x < ow; ++x) {
        float mx = -3.402823466e+38f;
        for (int a = 0; a < ksize; ++a)
          for (int b = 0; b < ksize; ++b) {
            const float v = p[(size_t)(y * stride + a) * w + (x * stride + b)];
            mx = (mx >= v) ? mx : v;
          }
        o[(size_t)y * ow + x] = mx;
      }
  }
}
