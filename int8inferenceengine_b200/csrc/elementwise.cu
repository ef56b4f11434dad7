// HBM-bound kernels of the i8ie INT8 path: quantize / dequantize / down_scale /
// min-max / relu / max-pool / layout glue / weight preparation.
// All of them are streaming kernels: 128-bit coalesced accesses, several
// independent loads in flight per thread, grids sized in multiples of the SM count.
#include <cfloat>
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace i8ie { int tc_error_sink_init(); }   // tc_gemm.cu

namespace i8ie {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

constexpr int kThreads = 256;

inline int stream_grid(int64_t work_items, int per_block) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)num_sms() * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// Persistent grid-stride kernels: exactly as many blocks as can be resident at once (occupancy of
// THIS kernel x SM count), so the grid-stride loop has no partial last wave.
template <typename K>
inline int resident_grid(K kernel, int64_t work_items, int per_block) {
  // tiny pointer-keyed cache (kernels with identical signatures share this instantiation)
  static std::atomic<const void*> keys[16];
  static std::atomic<int> vals[16];
  const void* key = reinterpret_cast<const void*>(kernel);
  int per_sm = 0;
  for (int i = 0; i < 16; ++i) {
    const void* k = keys[i].load(std::memory_order_acquire);
    if (k == key) { per_sm = vals[i].load(std::memory_order_relaxed); break; }
    if (k == nullptr) {
      int v = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kThreads, 0) != cudaSuccess || v < 1) v = 4;
      per_sm = v;
      const void* expect = nullptr;
      vals[i].store(v, std::memory_order_relaxed);
      if (keys[i].compare_exchange_strong(expect, key, std::memory_order_release)) break;
      if (expect == key) break;
    }
  }
  if (per_sm == 0) per_sm = 4;
  int64_t blocks = (work_items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3) == 0; }

// ---- A1 quantize f32 -> u8 (flat) -------------------------------------------
// Warp-coalesced streaming: each thread issues kU independent 128-bit loads, the j-th one of
// a warp covering one contiguous 512-byte run, and writes one 32-bit word per load (128 B
// per warp-store). FAST: the IEEE division x / scale is the FMA-corrected product with the
// hoisted reciprocal (exactly RN(x / scale), see quant_u8_wrap_fast) — no MUFU per element.
constexpr int kU = 4;

template <bool FAST>
__global__ void __launch_bounds__(kThreads) quantize_flat_kernel(const float* __restrict__ x,
                                                                 uint8_t* __restrict__ q, int64_t n,
                                                                 float scale, float zpf, int vec_ok,
                                                                 const float* const* __restrict__ xslot,
                                                                 float fast_lim) {
  if (xslot) x = *xslot;   // source address read at run time (CUDA-graph replay on a new input buffer)
  const int64_t nvec = vec_ok ? (n >> 2) : 0;   // float4 groups -> one output word each
  const float rcp = __frcp_rn(scale);
  const QuantFast2 qc = make_quant_fast2(scale, rcp, zpf);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  uint32_t* q32 = reinterpret_cast<uint32_t*>(q);
  for (int64_t base = (int64_t)blockIdx.x * (kThreads * kU); base < nvec; base += (int64_t)gridDim.x * (kThreads * kU)) {
    float4 f[kU];
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const int64_t v = base + j * kThreads + threadIdx.x;
      f[j] = (v < nvec) ? ld_stream_f4(x4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t w[kU];
    bool done = false;
    if (FAST) {
      // one magnitude test for the whole group, then the unguarded packed path
      float amax = 0.f;
#pragma unroll
      for (int j = 0; j < kU; ++j)
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(f[j].x), fabsf(f[j].y))), fmaxf(fabsf(f[j].z), fabsf(f[j].w)));
      if (amax < fast_lim) {
#pragma unroll
        for (int j = 0; j < kU; ++j) w[j] = quant4_fast(f[j].x, f[j].y, f[j].z, f[j].w, qc);
        done = true;
      }
    }
    if (!done) {
#pragma unroll
      for (int j = 0; j < kU; ++j)
        w[j] = quant_u8_wrap(f[j].x, scale, zpf) | (quant_u8_wrap(f[j].y, scale, zpf) << 8) |
               (quant_u8_wrap(f[j].z, scale, zpf) << 16) | (quant_u8_wrap(f[j].w, scale, zpf) << 24);
    }
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const int64_t v = base + j * kThreads + threadIdx.x;
      if (v < nvec) st_stream_u32(q32 + v, w[j]);
    }
  }
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (nvec << 2) + tid; i < n; i += nthreads) q[i] = (uint8_t)quant_u8_wrap(x[i], scale, zpf);
}

// ---- A1 fused with NCHW(f32) -> NHWC(u8, pitch cp) ---------------------------
// One thread per (pixel, 16-channel group): adjacent threads read adjacent pixels of
// the same plane (coalesced) and write one 16-byte NHWC segment each.
__global__ void __launch_bounds__(kThreads) quantize_nchw_nhwc_kernel(
    const float* __restrict__ x, uint8_t* __restrict__ q, int n, int c, int hw, int cp, float scale,
    float zpf, uint32_t zpb, const float* const* __restrict__ xslot) {
  pdl_launch_dependents();
  pdl_wait();
  if (xslot) x = *xslot;
  const int groups = cp >> 4;
  const int64_t total = (int64_t)n * hw * groups;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = t % ((int64_t)n * hw);   // pixel fastest -> coalesced plane reads
    const int g = (int)(t / ((int64_t)n * hw));
    const int img = (int)(pix / hw);
    const int p = (int)(pix % hw);
    const float* src = x + ((int64_t)img * c) * hw + p;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t word = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int ch = g * 16 + j * 4 + b;
        const uint32_t v = (ch < c) ? quant_u8_wrap(__ldg(src + (int64_t)ch * hw), scale, zpf) : zpb;
        word |= v << (8 * b);
      }
      w[j] = word;
    }
    *reinterpret_cast<uint4*>(q + pix * cp + g * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---- A5 dequantize u8 -> f32 (flat) -----------------------------------------
// One 32-bit load (4 codes) and one 128-bit store per thread and step, both warp-contiguous;
// kUD steps in flight. (float)(q - zp) is built without a convert: 0x4B0000qq is the float
// 2^23 + q, and (2^23 + q) - (2^23 + zp) is exact.
constexpr int kUD = 8;

__global__ void __launch_bounds__(kThreads) dequantize_flat_kernel(const uint8_t* __restrict__ q,
                                                                   float* __restrict__ x, int64_t n,
                                                                   float scale, int zp, int vec_ok) {
  const int64_t nvec = vec_ok ? (n >> 2) : 0;
  const uint32_t* q32 = reinterpret_cast<const uint32_t*>(q);
  float4* x4 = reinterpret_cast<float4*>(x);
  const float bias = 8388608.f + (float)zp;
  for (int64_t base = (int64_t)blockIdx.x * (kThreads * kUD); base < nvec; base += (int64_t)gridDim.x * (kThreads * kUD)) {
    uint32_t w[kUD];
#pragma unroll
    for (int j = 0; j < kUD; ++j) {
      const int64_t v = base + j * kThreads + threadIdx.x;
      w[j] = (v < nvec) ? ld_stream_u32(q32 + v) : 0u;
    }
#pragma unroll
    for (int j = 0; j < kUD; ++j) {
      const int64_t v = base + j * kThreads + threadIdx.x;
      float4 f;
      f.x = __fmul_rn(__fsub_rn(__uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7650)), bias), scale);
      f.y = __fmul_rn(__fsub_rn(__uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7651)), bias), scale);
      f.z = __fmul_rn(__fsub_rn(__uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7652)), bias), scale);
      f.w = __fmul_rn(__fsub_rn(__uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7653)), bias), scale);
      if (v < nvec) st_stream_f4(x4 + v, f);
    }
  }
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (nvec << 2) + tid; i < n; i += nthreads) x[i] = dequant_f32(q[i], zp, scale);
}

__global__ void dequantize_rows_kernel(const uint8_t* __restrict__ q, float* __restrict__ x, int rows,
                                       int cols, int pitch, float scale, int zp) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = (int64_t)rows * cols;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(t / cols), c = (int)(t % cols);
    x[t] = dequant_f32(q[(int64_t)r * pitch + c], zp, scale);
  }
}

// ---- A4 down_scale s32 -> u8 (flat) -------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(kThreads) downscale_flat_kernel(const int32_t* __restrict__ acc,
                                                                  uint8_t* __restrict__ y, int64_t n,
                                                                  float sa, float sb, float sc,
                                                                  float zpf, int vec_ok) {
  const int64_t nvec = vec_ok ? (n >> 2) : 0;
  const RequantFast2 rq = make_requant_fast2(sa, sb, sc, __frcp_rn(sc), zpf);
  const uint4* a4 = reinterpret_cast<const uint4*>(acc);
  uint32_t* y32 = reinterpret_cast<uint32_t*>(y);
  for (int64_t base = (int64_t)blockIdx.x * (kThreads * kU); base < nvec; base += (int64_t)gridDim.x * (kThreads * kU)) {
    uint4 a[kU];
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const int64_t v = base + j * kThreads + threadIdx.x;
      a[j] = (v < nvec) ? ld_stream_u4(a4 + v) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const int64_t v = base + j * kThreads + threadIdx.x;
      uint32_t w;
      if (FAST) {
        uint32_t y0, y1, y2, y3;
        requant2_u8_fast<false>((int)a[j].x, (int)a[j].y, rq, y0, y1);
        requant2_u8_fast<false>((int)a[j].z, (int)a[j].w, rq, y2, y3);
        w = y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
      } else {
        w = requant_u8((int)a[j].x, sa, sb, sc, zpf) | (requant_u8((int)a[j].y, sa, sb, sc, zpf) << 8) |
            (requant_u8((int)a[j].z, sa, sb, sc, zpf) << 16) | (requant_u8((int)a[j].w, sa, sb, sc, zpf) << 24);
      }
      if (v < nvec) st_stream_u32(y32 + v, w);
    }
  }
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (nvec << 2) + tid; i < n; i += nthreads) y[i] = (uint8_t)requant_u8(acc[i], sa, sb, sc, zpf);
}

// ---- A9 min/max reduction ------------------------------------------------------
// warp shuffle -> block smem -> per-block partial -> the last block to finish folds
// the partials (single launch, deterministic: min/max are order-independent).
constexpr int kMinMaxMaxBlocks = 148 * 8;
struct MinMaxWs {
  unsigned int done;
  unsigned int pad[3];
  float mn[kMinMaxMaxBlocks];
  float mx[kMinMaxMaxBlocks];
};

__device__ __forceinline__ void warp_minmax(float& mn, float& mx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
}

__global__ void __launch_bounds__(kThreads) minmax_kernel(const float* __restrict__ x, int64_t n,
                                                          float* __restrict__ out2, MinMaxWs* ws,
                                                          int vec_ok) {
  __shared__ float smn[kThreads / 32], smx[kThreads / 32];
  __shared__ bool is_last;
  float mn = FLT_MAX, mx = -FLT_MAX;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t nvec = vec_ok ? (n >> 4) : 0;  // 16 floats = 4 x float4 per thread-iteration
  for (int64_t v = tid; v < nvec; v += nthreads) {
    const float4* p = reinterpret_cast<const float4*>(x) + v * 4;
    float4 f0 = ld_stream_f4(p), f1 = ld_stream_f4(p + 1), f2 = ld_stream_f4(p + 2), f3 = ld_stream_f4(p + 3);
    mn = fminf(mn, fminf(fminf(fminf(f0.x, f0.y), fminf(f0.z, f0.w)), fminf(fminf(f1.x, f1.y), fminf(f1.z, f1.w))));
    mn = fminf(mn, fminf(fminf(fminf(f2.x, f2.y), fminf(f2.z, f2.w)), fminf(fminf(f3.x, f3.y), fminf(f3.z, f3.w))));
    mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(f0.x, f0.y), fmaxf(f0.z, f0.w)), fmaxf(fmaxf(f1.x, f1.y), fmaxf(f1.z, f1.w))));
    mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(f2.x, f2.y), fmaxf(f2.z, f2.w)), fmaxf(fmaxf(f3.x, f3.y), fmaxf(f3.z, f3.w))));
  }
  for (int64_t i = (nvec << 4) + tid; i < n; i += nthreads) {
    mn = fminf(mn, x[i]);
    mx = fmaxf(mx, x[i]);
  }
  warp_minmax(mn, mx);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { smn[wid] = mn; smx[wid] = mx; }
  __syncthreads();
  if (wid == 0) {
    mn = (lane < kThreads / 32) ? smn[lane] : FLT_MAX;
    mx = (lane < kThreads / 32) ? smx[lane] : -FLT_MAX;
    warp_minmax(mn, mx);
    if (lane == 0) {
      ws->mn[blockIdx.x] = mn;
      ws->mx[blockIdx.x] = mx;
      __threadfence();
      const unsigned int prev = atomicAdd(&ws->done, 1u);
      is_last = (prev == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    mn = FLT_MAX; mx = -FLT_MAX;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
      mn = fminf(mn, ((volatile float*)ws->mn)[b]);
      mx = fmaxf(mx, ((volatile float*)ws->mx)[b]);
    }
    warp_minmax(mn, mx);
    __syncthreads();
    if (lane == 0) { smn[wid] = mn; smx[wid] = mx; }
    __syncthreads();
    if (wid == 0) {
      mn = (lane < kThreads / 32) ? smn[lane] : FLT_MAX;
      mx = (lane < kThreads / 32) ? smx[lane] : -FLT_MAX;
      warp_minmax(mn, mx);
      if (lane == 0) {
        out2[0] = mn;
        out2[1] = mx;
        ws->done = 0;  // workspace is reusable without re-zeroing
      }
    }
  }
}

// ---- A10 relu<u8> --------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) relu_flat_kernel(const uint8_t* __restrict__ x,
                                                             uint8_t* __restrict__ y, int64_t n,
                                                             uint32_t zp4, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t nvec = vec_ok ? (n >> 4) : 0;
  for (int64_t v = tid; v < nvec; v += nthreads) {
    uint4 a = ld_stream_u4(reinterpret_cast<const uint4*>(x) + v);
    a.x = __vmaxu4(a.x, zp4); a.y = __vmaxu4(a.y, zp4); a.z = __vmaxu4(a.z, zp4); a.w = __vmaxu4(a.w, zp4);
    st_stream_u4(reinterpret_cast<uint4*>(y) + v, a);
  }
  const uint8_t z = (uint8_t)(zp4 & 0xff);
  for (int64_t i = (nvec << 4) + tid; i < n; i += nthreads) y[i] = x[i] > z ? x[i] : z;
}

// ---- indirect copy: materialises a run-time-addressed source into a fixed buffer --------------
__global__ void __launch_bounds__(kThreads) copy_indirect_kernel(const uint4* const* __restrict__ slot,
                                                                 uint4* __restrict__ dst, int64_t nvec) {
  const uint4* __restrict__ src = *slot;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x)
    st_stream_u4(dst + v, ld_stream_u4(src + v));
}

// ---- multi-GPU result exchange: pack [agreement count | logits] for ONE all-gather -----------
// chunk layout (bytes): int64 agree | int32 rows | int32 cols | rows_cap*cols fp32 logits
struct ExchangeHeader {
  long long agree;
  int rows, cols;
};

// One block. Row r: argmax (first maximum, like numpy/torch) compared with ref_argmax[r]; the
// logits are copied behind the header.
__global__ void __launch_bounds__(1024) top1_pack_kernel(const float* __restrict__ logits,
                                                         const long long* __restrict__ ref_argmax, int rows,
                                                         int cols, uint8_t* __restrict__ packed) {
  __shared__ int s_cnt[32];
  int cnt = 0;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float* row = logits + (size_t)r * cols;
    int best = 0;
    float bv = row[0];
    for (int j = 1; j < cols; ++j) {
      const float v = row[j];
      if (v > bv) { bv = v; best = j; }
    }
    if (ref_argmax && (long long)best == ref_argmax[r]) ++cnt;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x < 32) {
    cnt = (threadIdx.x < (blockDim.x >> 5)) ? s_cnt[threadIdx.x] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (threadIdx.x == 0) {
      ExchangeHeader* h = reinterpret_cast<ExchangeHeader*>(packed);
      h->agree = cnt; h->rows = rows; h->cols = cols;
    }
  }
  float* dst = reinterpret_cast<float*>(packed + sizeof(ExchangeHeader));
  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) dst[i] = logits[i];
}

// One block. Gathered chunks (rank order, `chunk_bytes` apart) -> dense logits of all ranks in
// rank order (only each rank's valid rows) + the summed agreement count.
__global__ void __launch_bounds__(1024) top1_unpack_kernel(const uint8_t* __restrict__ gathered, int world,
                                                           long long chunk_bytes, float* __restrict__ logits_all,
                                                           long long* __restrict__ agree_total) {
  long long total = 0;
  int row0 = 0;
  for (int k = 0; k < world; ++k) {
    const uint8_t* chunk = gathered + (size_t)k * chunk_bytes;
    const ExchangeHeader h = *reinterpret_cast<const ExchangeHeader*>(chunk);
    const float* src = reinterpret_cast<const float*>(chunk + sizeof(ExchangeHeader));
    float* dst = logits_all + (size_t)row0 * h.cols;
    for (int i = threadIdx.x; i < h.rows * h.cols; i += blockDim.x) dst[i] = src[i];
    total += h.agree;
    row0 += h.rows;
  }
  if (threadIdx.x == 0) *agree_total = total;
}

// ---- A11 max_pool2d<u8>, NHWC --------------------------------------------------
// Byte-wise unsigned max of 16 channels, kept as even/odd bytes in u16x2 lanes: one tap word
// costs 2 PRMT + 2 VIMNMX.U16x2 (the __vmaxu4 intrinsic is a 7-instruction emulation on sm_100).
struct U8x16Max {
  uint32_t e[4], o[4];
  __device__ __forceinline__ void init() {   // min<u8_t>() == 0 (functional.cc:33-35)
#pragma unroll
    for (int i = 0; i < 4; ++i) { e[i] = 0; o[i] = 0; }
  }
  __device__ __forceinline__ void one(int i, uint32_t w) {
    const uint32_t we = __byte_perm(w, 0, 0x4240), wo = __byte_perm(w, 0, 0x4341);
    asm("max.u16x2 %0, %0, %1;" : "+r"(e[i]) : "r"(we));
    asm("max.u16x2 %0, %0, %1;" : "+r"(o[i]) : "r"(wo));
  }
  __device__ __forceinline__ void tap(const uint4& v) { one(0, v.x); one(1, v.y); one(2, v.z); one(3, v.w); }
  __device__ __forceinline__ uint4 result() const {
    return make_uint4(__byte_perm(e[0], o[0], 0x6240), __byte_perm(e[1], o[1], 0x6240),
                      __byte_perm(e[2], o[2], 0x6240), __byte_perm(e[3], o[3], 0x6240));
  }
};

// window max of one (output pixel, 16-channel group); `base` points at the window's first tap.
// Pad groups (all lanes hold the zero point everywhere) need a single tap.
template <int KS>
__device__ __forceinline__ uint4 pool_window(const uint8_t* __restrict__ base, int w, int cp, int ks_rt, bool pad_group) {
  const int ks = KS > 0 ? KS : ks_rt;
  if (pad_group) return __ldg(reinterpret_cast<const uint4*>(base));
  U8x16Max m;
  m.init();
  if (KS > 0) {
    uint4 v[KS > 0 ? KS * KS : 1];   // all k*k loads in flight before the first max
#pragma unroll
    for (int a = 0; a < KS; ++a)
#pragma unroll
      for (int b = 0; b < KS; ++b)
        v[a * KS + b] = __ldg(reinterpret_cast<const uint4*>(base + ((int64_t)a * w + b) * cp));
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) m.tap(v[i]);
  } else {
    for (int a = 0; a < ks; ++a)
      for (int b = 0; b < ks; ++b) m.tap(__ldg(reinterpret_cast<const uint4*>(base + ((int64_t)a * w + b) * cp)));
  }
  return m.result();
}

// One block per output row (img, oy): thread = (ox, 16-channel group) — no per-thread division
// chain when the group count is a power of two. NHWC output.
template <int KS>
__global__ void __launch_bounds__(kThreads) maxpool_nhwc_rows_kernel(const uint8_t* __restrict__ x,
                                                                     uint8_t* __restrict__ y, int h, int w,
                                                                     int c, int cp, int ks_rt, int st, int oh,
                                                                     int ow, int gshift) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = cp >> 4;
  const int img = blockIdx.x / oh, oy = blockIdx.x - img * oh;
  const uint8_t* xrow = x + ((int64_t)img * h + (int64_t)oy * st) * w * cp;
  uint8_t* yrow = y + ((int64_t)img * oh + oy) * ow * cp;
  for (int t = threadIdx.x; t < ow * groups; t += blockDim.x) {
    const int ox = gshift >= 0 ? (t >> gshift) : (t / groups);
    const int g = t - ox * groups;
    const uint4 m = pool_window<KS>(xrow + (int64_t)ox * st * cp + g * 16, w, cp, ks_rt, g * 16 >= c);
    *reinterpret_cast<uint4*>(yrow + (int64_t)ox * cp + g * 16) = m;
  }
}

// Same, into a PHYSICALLY PADDED NHWC tensor for a following "row mode" convolution (tc_gemm.cu):
//   y[img][oh + 2*opad][ow + 2*opad][ocp], border pixels = zp (conv2d.cc:24-25 pads with the input
// zero point), channel pitch ocp <= cp (only the groups that hold real channels are kept).
// One block per padded output row; ks = st = 1 makes it a plain pad-copy.
template <int KS>
__global__ void __launch_bounds__(kThreads) maxpool_nhwc_rows_padded_kernel(const uint8_t* __restrict__ x,
                                                                            uint8_t* __restrict__ y, int h, int w,
                                                                            int c, int cp, int ks_rt, int st, int oh,
                                                                            int ow, int ocp, int opad, uint32_t zp4) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = ocp >> 4;
  const int ohp = oh + 2 * opad, owp = ow + 2 * opad;
  const int img = blockIdx.x / ohp, oyp = blockIdx.x - img * ohp;
  const int oy = oyp - opad;
  uint8_t* yrow = y + ((int64_t)img * ohp + oyp) * owp * ocp;
  const uint4 z = make_uint4(zp4, zp4, zp4, zp4);
  if (oy < 0 || oy >= oh) {
    for (int t = threadIdx.x; t < owp * groups; t += blockDim.x) reinterpret_cast<uint4*>(yrow)[t] = z;
    return;
  }
  const uint8_t* xrow = x + ((int64_t)img * h + (int64_t)oy * st) * w * cp;
  for (int t = threadIdx.x; t < owp * groups; t += blockDim.x) {
    const int oxp = t / groups, g = t - oxp * groups;
    const int ox = oxp - opad;
    uint4 m = z;
    if (ox >= 0 && ox < ow) m = pool_window<KS>(xrow + (int64_t)ox * st * cp + g * 16, w, cp, ks_rt, g * 16 >= c);
    reinterpret_cast<uint4*>(yrow)[t] = m;
  }
}

// Pool + flatten: one block per image; the NCHW-ordered result [c][oh*ow] is assembled in
// shared memory and written out with coalesced 128-bit stores (it is contiguous per image).
template <int KS>
__global__ void __launch_bounds__(kThreads) maxpool_nhwc_to_nchw_image_kernel(const uint8_t* __restrict__ x,
                                                                              uint8_t* __restrict__ y, int h,
                                                                              int w, int c, int cp, int ks_rt,
                                                                              int st, int oh, int ow) {
  extern __shared__ __align__(16) uint8_t s_out[];
  pdl_launch_dependents();
  pdl_wait();
  const int groups = (c + 15) >> 4;   // only groups holding real channels
  const int img = blockIdx.x, plane = oh * ow;
  const uint8_t* ximg = x + (int64_t)img * h * w * cp;
  for (int t = threadIdx.x; t < plane * groups; t += blockDim.x) {
    const int g = t % groups, pix = t / groups;
    const int oy = pix / ow, ox = pix - oy * ow;
    const uint4 m = pool_window<KS>(ximg + (((int64_t)oy * st) * w + (int64_t)ox * st) * cp + g * 16, w, cp, ks_rt, false);
    const uint32_t wds[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int ch = g * 16 + j;
      if (ch < c) s_out[ch * plane + pix] = (uint8_t)(wds[j >> 2] >> (8 * (j & 3)));
    }
  }
  __syncthreads();
  const int total = c * plane;   // multiple of 16 (checked by the host)
  uint4* dst = reinterpret_cast<uint4*>(y + (int64_t)img * total);
  for (int v = threadIdx.x; v < (total >> 4); v += blockDim.x) dst[v] = reinterpret_cast<const uint4*>(s_out)[v];
}

// Generic fallback (any size, NCHW scatter): one thread per (output pixel, 16-channel group).
template <int KS>
__global__ void __launch_bounds__(kThreads) maxpool_nhwc_kernel(const uint8_t* __restrict__ x,
                                                                uint8_t* __restrict__ y, int n, int h,
                                                                int w, int c, int cp, int ks_rt, int st,
                                                                int oh, int ow, int out_nchw) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = cp >> 4;
  const int64_t total = (int64_t)n * oh * ow * groups;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(t % groups);
    int64_t r = t / groups;
    const int ox = (int)(r % ow); r /= ow;
    const int oy = (int)(r % oh);
    const int img = (int)(r / oh);
    const uint8_t* base = x + (((int64_t)img * h + oy * st) * w + ox * st) * cp + g * 16;
    const uint4 m = pool_window<KS>(base, w, cp, ks_rt, g * 16 >= c);
    if (!out_nchw) {
      *reinterpret_cast<uint4*>(y + (((int64_t)img * oh + oy) * ow + ox) * cp + g * 16) = m;
    } else {
      const uint32_t wds[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int ch = g * 16 + j;
        if (ch < c) y[(((int64_t)img * c + ch) * oh + oy) * ow + ox] = (uint8_t)(wds[j >> 2] >> (8 * (j & 3)));
      }
    }
  }
}

// ---- layout glue ---------------------------------------------------------------
__global__ void u8_nchw_to_nhwc_kernel(const uint8_t* __restrict__ x, uint8_t* __restrict__ y, int n,
                                       int c, int hw, int cp, uint32_t padv) {
  const int groups = cp >> 4;
  const int64_t npix = (int64_t)n * hw;
  const int64_t total = npix * groups;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = t % npix;
    const int g = (int)(t / npix);
    const int img = (int)(pix / hw), p = (int)(pix % hw);
    const uint8_t* src = x + ((int64_t)img * c) * hw + p;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t word = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int ch = g * 16 + j * 4 + b;
        const uint32_t v = (ch < c) ? (uint32_t)__ldg(src + (int64_t)ch * hw) : padv;
        word |= v << (8 * b);
      }
      w[j] = word;
    }
    *reinterpret_cast<uint4*>(y + pix * cp + g * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void u8_nhwc_to_nchw_kernel(const uint8_t* __restrict__ x, uint8_t* __restrict__ y, int n,
                                       int c, int hw, int cp) {
  const int64_t total = (int64_t)n * c * hw;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(t % hw);
    int64_t r = t / hw;
    const int ch = (int)(r % c);
    const int img = (int)(r / c);
    y[t] = x[((int64_t)img * hw + p) * cp + ch];
  }
}

// ---- A2b / A3 offsets ----------------------------------------------------------
// One warp per output row. t is a sequential fp32 sum of the INTEGERS zp*w[k]; while
// zp * sum|w| < 2^24 every partial sum is an exactly representable integer, so the
// sequential fp32 result equals the exact integer sum and the warp computes it in
// parallel. Otherwise lane 0 replays the reference's sequential fp32 loop.
__global__ void __launch_bounds__(kThreads) zp_offsets_kernel(const int8_t* __restrict__ qw,
                                                              const int8_t* __restrict__ qb, int n,
                                                              int k, int zp, float in_scale,
                                                              int is_conv, int32_t* __restrict__ oc,
                                                              float* __restrict__ bias_f) {
  const int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const int8_t* wr = qw + (int64_t)row * k;
  long long s = 0, sabs = 0;
  int i0 = 0;
  if ((k & 15) == 0 && ((reinterpret_cast<uintptr_t>(wr) & 15) == 0)) {   // 128-bit loads, byte sums via dp4a
    int ps = 0;
    unsigned pa = 0;
    for (int i = lane * 16; i < k; i += 32 * 16) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(wr + i));
      ps = __dp4a((int)v.x, 0x01010101, ps); ps = __dp4a((int)v.y, 0x01010101, ps);
      ps = __dp4a((int)v.z, 0x01010101, ps); ps = __dp4a((int)v.w, 0x01010101, ps);
      pa = __dp4a(__vabs4(v.x), 0x01010101u, pa); pa = __dp4a(__vabs4(v.y), 0x01010101u, pa);
      pa = __dp4a(__vabs4(v.z), 0x01010101u, pa); pa = __dp4a(__vabs4(v.w), 0x01010101u, pa);
    }
    s = ps; sabs = pa;
    i0 = k;
  }
  for (int i = i0 + lane; i < k; i += 32) {
    const int v = wr[i];
    s += v;
    sabs += (v < 0) ? -v : v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sabs += __shfl_xor_sync(0xffffffffu, sabs, o);
  }
  if (lane != 0) return;
  float t;
  if ((long long)zp * sabs < (1ll << 24)) {
    t = (float)((long long)zp * s);
  } else {
    t = 0.f;
    for (int i = 0; i < k; ++i) t = __fadd_rn(t, __int2float_rn(zp * (int)wr[i]));
  }
  const float b = __fdiv_rn(__int2float_rn((int)qb[row]), in_scale);
  if (is_conv) {
    oc[row] = __float2int_rz(__fsub_rn(b, t));  // conv2d.cc:123
    bias_f[row] = 0.f;
  } else {
    oc[row] = __float2int_rz(-t);               // fully_connected.cc:37
    bias_f[row] = b;                            // fully_connected.cc:44
  }
}

__global__ void pack_conv_weight_kernel(const int8_t* __restrict__ qw, int8_t* __restrict__ wp, int kc,
                                        int c, int kh, int kw, int kc_pad, int cp) {
  const int64_t total = (int64_t)kc_pad * kh * kw * cp;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(t % cp);
    int64_t r = t / cp;
    const int x = (int)(r % kw); r /= kw;
    const int y = (int)(r % kh);
    const int o = (int)(r / kh);
    wp[t] = (o < kc && ch < c) ? qw[(((int64_t)o * c + ch) * kh + y) * kw + x] : (int8_t)0;
  }
}

}  // namespace
}  // namespace i8ie

using namespace i8ie;

extern "C" {

const char* i8ie_last_error(void) { return g_err; }
const char* i8ie_version(void) { return "i8ie_sm100 0.1 (sm_100a)"; }
int64_t i8ie_launch_count(void) { return g_launches.load(); }

int i8ie_device_check(void) {
  int dev = 0;
  I8IE_CUDA_OK(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  I8IE_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  I8IE_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10 || minor != 0) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
    return I8IE_ENOTSM100;
  }
  return i8ie::tc_error_sink_init();   // host-mapped mirror of the tensor-core kernels' protocol-error flag
}

static int quantize_flat(const float* x, const float* const* xslot, uint8_t* q, int64_t n, float scale, int zp,
                         void* stream) {
  I8IE_REQUIRE(n >= 0 && zp >= 0 && zp <= 255, "quantize: bad n/zp");
  if (n == 0) return I8IE_OK;
  const int vec = (xslot || aligned16(x)) && aligned4(q);
  const float lim = quant_fast_limit(scale);
  if (lim > 0.f)
    quantize_flat_kernel<true><<<resident_grid(quantize_flat_kernel<true>, (n + 3) / 4, kThreads * kU), kThreads, 0,
                                 (cudaStream_t)stream>>>(x, q, n, scale, (float)zp, vec, xslot, lim);
  else
    quantize_flat_kernel<false><<<resident_grid(quantize_flat_kernel<false>, (n + 3) / 4, kThreads * kU), kThreads, 0,
                                  (cudaStream_t)stream>>>(x, q, n, scale, (float)zp, vec, xslot, 0.f);
  return check_launch("quantize_flat_kernel");
}

int i8ie_quantize_f32_u8(const float* x, uint8_t* q, int64_t n, float scale, int zp, void* stream) {
  return quantize_flat(x, nullptr, q, n, scale, zp, stream);
}

int i8ie_quantize_f32_u8_indirect(const float* const* x_slot, uint8_t* q, int64_t n, float scale, int zp,
                                  void* stream) {
  I8IE_REQUIRE(x_slot != nullptr, "quantize_indirect: null slot");
  return quantize_flat(nullptr, x_slot, q, n, scale, zp, stream);
}

int i8ie_copy_indirect(const void* const* src_slot, void* dst, int64_t nbytes, void* stream) {
  I8IE_REQUIRE(src_slot && dst && nbytes >= 0 && nbytes % 16 == 0 && aligned16(dst), "copy_indirect: bad arguments");
  if (nbytes == 0) return I8IE_OK;
  copy_indirect_kernel<<<stream_grid(nbytes / 16, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4* const*>(src_slot), reinterpret_cast<uint4*>(dst), nbytes / 16);
  return check_launch("copy_indirect_kernel");
}

static int quantize_nhwc(const float* x, const float* const* xslot, uint8_t* q, int n, int c, int h, int w, int cp,
                         float scale, int zp, void* stream) {
  I8IE_REQUIRE(cp % 16 == 0 && cp >= c && zp >= 0 && zp <= 255 && aligned16(q), "quantize_nhwc: bad cp/zp/alignment");
  const int64_t items = (int64_t)n * h * w * (cp / 16);
  if (items == 0) return I8IE_OK;
  launch_pdl(quantize_nchw_nhwc_kernel, dim3(stream_grid(items, kThreads)), dim3(kThreads), 0, (cudaStream_t)stream,
             x, q, n, c, h * w, cp, scale, (float)zp, (uint32_t)zp, xslot);
  return check_launch("quantize_nchw_nhwc_kernel");
}

int i8ie_quantize_nchw_f32_nhwc_u8(const float* x, uint8_t* q, int n, int c, int h, int w, int cp,
                                   float scale, int zp, void* stream) {
  return quantize_nhwc(x, nullptr, q, n, c, h, w, cp, scale, zp, stream);
}

int i8ie_quantize_nchw_f32_nhwc_u8_indirect(const float* const* x_slot, uint8_t* q, int n, int c, int h, int w,
                                            int cp, float scale, int zp, void* stream) {
  I8IE_REQUIRE(x_slot != nullptr, "quantize_nhwc_indirect: null slot");
  return quantize_nhwc(nullptr, x_slot, q, n, c, h, w, cp, scale, zp, stream);
}

int i8ie_dequantize_u8_f32(const uint8_t* q, float* x, int64_t n, float scale, int zp, void* stream) {
  I8IE_REQUIRE(n >= 0, "dequantize: bad n");
  if (n == 0) return I8IE_OK;
  const int vec = aligned16(x) && aligned4(q) && zp >= 0 && zp <= 255;
  dequantize_flat_kernel<<<resident_grid(dequantize_flat_kernel, (n + 3) / 4, kThreads * kUD), kThreads, 0,
                           (cudaStream_t)stream>>>(q, x, n, scale, zp, vec);
  return check_launch("dequantize_flat_kernel");
}

int i8ie_dequantize_rows_u8_f32(const uint8_t* q, float* x, int rows, int cols, int pitch, float scale,
                                int zp, void* stream) {
  I8IE_REQUIRE(rows >= 0 && cols >= 0 && pitch >= cols, "dequantize_rows: bad shape");
  if ((int64_t)rows * cols == 0) return I8IE_OK;
  launch_pdl(dequantize_rows_kernel, dim3(stream_grid((int64_t)rows * cols, kThreads)), dim3(kThreads), 0,
             (cudaStream_t)stream, q, x, rows, cols, pitch, scale, zp);
  return check_launch("dequantize_rows_kernel");
}

int i8ie_downscale_s32_u8(const int32_t* acc, uint8_t* y, int64_t n, float sa, float sb, float sc,
                          int zp_c, void* stream) {
  I8IE_REQUIRE(n >= 0, "downscale: bad n");
  if (n == 0) return I8IE_OK;
  const int vec = aligned16(acc) && aligned4(y);
  if (requant_fast_ok(sa, sb, sc))
    downscale_flat_kernel<true><<<resident_grid(downscale_flat_kernel<true>, (n + 3) / 4, kThreads * kU), kThreads, 0,
                                  (cudaStream_t)stream>>>(acc, y, n, sa, sb, sc, (float)zp_c, vec);
  else
    downscale_flat_kernel<false><<<resident_grid(downscale_flat_kernel<false>, (n + 3) / 4, kThreads * kU), kThreads, 0,
                                   (cudaStream_t)stream>>>(acc, y, n, sa, sb, sc, (float)zp_c, vec);
  return check_launch("downscale_flat_kernel");
}

int64_t i8ie_top1_chunk_bytes(int rows_cap, int cols) {
  return (int64_t)sizeof(ExchangeHeader) + (int64_t)rows_cap * cols * 4;
}

int i8ie_top1_pack(const float* logits, const int64_t* ref_argmax, int rows, int cols, void* packed, void* stream) {
  I8IE_REQUIRE(logits && packed && rows >= 0 && cols > 0 && aligned16(packed), "top1_pack: bad arguments");
  top1_pack_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, reinterpret_cast<const long long*>(ref_argmax), rows,
                                                       cols, reinterpret_cast<uint8_t*>(packed));
  return check_launch("top1_pack_kernel");
}

int i8ie_top1_unpack(const void* gathered, int world, int64_t chunk_bytes, float* logits_all, int64_t* agree_total,
                     void* stream) {
  I8IE_REQUIRE(gathered && logits_all && agree_total && world > 0 && chunk_bytes >= (int64_t)sizeof(ExchangeHeader) &&
                   chunk_bytes % 16 == 0 && aligned16(gathered),
               "top1_unpack: bad arguments");
  top1_unpack_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(gathered), world,
                                                         (long long)chunk_bytes, logits_all,
                                                         reinterpret_cast<long long*>(agree_total));
  return check_launch("top1_unpack_kernel");
}

int64_t i8ie_minmax_workspace_bytes(void) { return (int64_t)sizeof(MinMaxWs); }

int i8ie_minmax_f32(const float* x, int64_t n, float* out2, void* workspace, void* stream) {
  I8IE_REQUIRE(n > 0 && workspace != nullptr, "minmax: n must be > 0 and workspace non-null");
  int grid = resident_grid(minmax_kernel, (n + 15) / 16, kThreads);
  if (grid > kMinMaxMaxBlocks) grid = kMinMaxMaxBlocks;
  minmax_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, n, out2, (MinMaxWs*)workspace, aligned16(x));
  return check_launch("minmax_kernel");
}

// calibrator.cc:28-35
int i8ie_range_from_minmax_host(float mn, float mx, float* scale_host, uint8_t* zp_host) {
  float out_min = fminf(mn, 0.f), out_max = fmaxf(mx, 0.f);
  const double q = (double)(255 * (0 - out_min)) / ((double)(out_max - out_min) + 1e-09);
  const uint8_t zp = (uint8_t)(int32_t)q;
  float s = (zp == 0) ? (out_max - out_min) / 255 : (0 - out_min) / zp;
  if (s == 0) s = 1;
  *scale_host = s;
  *zp_host = zp;
  return I8IE_OK;
}

int i8ie_relu_u8(const uint8_t* x, uint8_t* y, int64_t n, int zp, void* stream) {
  I8IE_REQUIRE(n >= 0 && zp >= 0 && zp <= 255, "relu: bad n/zp");
  if (n == 0) return I8IE_OK;
  const uint32_t z = (uint32_t)zp;
  relu_flat_kernel<<<resident_grid(relu_flat_kernel, (n + 15) / 16, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      x, y, n, z | (z << 8) | (z << 16) | (z << 24), aligned16(x) && aligned16(y));
  return check_launch("relu_flat_kernel");
}

int i8ie_maxpool_u8_nhwc(const uint8_t* x, uint8_t* y, int n, int h, int w, int c, int cp, int ksize,
                         int stride, int out_nchw, void* stream) {
  I8IE_REQUIRE(cp % 16 == 0 && cp >= c && ksize >= 1 && stride >= 1 && h >= ksize && w >= ksize &&
                   aligned16(x) && (out_nchw || aligned16(y)),
               "maxpool: bad shape/alignment");
  const int oh = (h - ksize) / stride + 1, ow = (w - ksize) / stride + 1;
  const int64_t items = (int64_t)n * oh * ow * (cp / 16);
  if (items == 0) return I8IE_OK;
  cudaStream_t s = (cudaStream_t)stream;
#define I8IE_POOL_KS(KERNEL, GRID, SMEM, ...)                                                     \
  do {                                                                                            \
    PdlFamily fam_(kPdlPool);                                                                      \
    if (ksize == 3) launch_pdl(KERNEL<3>, dim3(GRID), dim3(kThreads), SMEM, s, __VA_ARGS__);       \
    else if (ksize == 2) launch_pdl(KERNEL<2>, dim3(GRID), dim3(kThreads), SMEM, s, __VA_ARGS__);  \
    else launch_pdl(KERNEL<0>, dim3(GRID), dim3(kThreads), SMEM, s, __VA_ARGS__);                  \
  } while (0)
  const int64_t img_bytes = (int64_t)c * oh * ow;
  if (!out_nchw && (int64_t)n * oh < (1ll << 30)) {
    const int groups = cp / 16;
    int gshift = -1;
    for (int b = 0; b < 12; ++b) if ((1 << b) == groups) gshift = b;
    I8IE_POOL_KS(maxpool_nhwc_rows_kernel, n * oh, 0, x, y, h, w, c, cp, ksize, stride, oh, ow, gshift);
  } else if (out_nchw && img_bytes <= 48 * 1024 && img_bytes % 16 == 0 && aligned16(y)) {
    I8IE_POOL_KS(maxpool_nhwc_to_nchw_image_kernel, n, (size_t)img_bytes, x, y, h, w, c, cp, ksize, stride, oh, ow);
  } else {
    const int grid = stream_grid(items, kThreads);
    I8IE_POOL_KS(maxpool_nhwc_kernel, grid, 0, x, y, n, h, w, c, cp, ksize, stride, oh, ow, out_nchw);
  }
#undef I8IE_POOL_KS
  return check_launch("maxpool_nhwc_kernel");
}

int i8ie_maxpool_u8_nhwc_padded(const uint8_t* x, uint8_t* y, int n, int h, int w, int c, int cp, int ksize,
                                int stride, int out_cp, int out_pad, int pad_value, void* stream) {
  I8IE_REQUIRE(cp % 16 == 0 && cp >= c && out_cp % 16 == 0 && out_cp >= c && out_cp <= cp && ksize >= 1 &&
                   stride >= 1 && h >= ksize && w >= ksize && out_pad >= 0 && pad_value >= 0 && pad_value <= 255 &&
                   aligned16(x) && aligned16(y),
               "maxpool_padded: bad shape/alignment");
  const int oh = (h - ksize) / stride + 1, ow = (w - ksize) / stride + 1;
  const int64_t rows = (int64_t)n * (oh + 2 * out_pad);
  if (rows == 0) return I8IE_OK;
  I8IE_REQUIRE(rows < (1ll << 30), "maxpool_padded: too many rows");
  cudaStream_t s = (cudaStream_t)stream;
  const uint32_t zp4 = (uint32_t)pad_value * 0x01010101u;
  PdlFamily fam_(kPdlPool);
  if (ksize == 3)
    launch_pdl(maxpool_nhwc_rows_padded_kernel<3>, dim3((unsigned)rows), dim3(kThreads), 0, s, x, y, h, w, c, cp, ksize,
               stride, oh, ow, out_cp, out_pad, zp4);
  else if (ksize == 2)
    launch_pdl(maxpool_nhwc_rows_padded_kernel<2>, dim3((unsigned)rows), dim3(kThreads), 0, s, x, y, h, w, c, cp, ksize,
               stride, oh, ow, out_cp, out_pad, zp4);
  else
    launch_pdl(maxpool_nhwc_rows_padded_kernel<0>, dim3((unsigned)rows), dim3(kThreads), 0, s, x, y, h, w, c, cp, ksize,
               stride, oh, ow, out_cp, out_pad, zp4);
  return check_launch("maxpool_nhwc_rows_padded_kernel");
}

int i8ie_u8_nchw_to_nhwc(const uint8_t* x, uint8_t* y, int n, int c, int h, int w, int cp, int pad_value,
                         void* stream) {
  I8IE_REQUIRE(cp % 16 == 0 && cp >= c && aligned16(y), "nchw_to_nhwc: bad cp/alignment");
  const int64_t items = (int64_t)n * h * w * (cp / 16);
  if (items == 0) return I8IE_OK;
  u8_nchw_to_nhwc_kernel<<<stream_grid(items, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      x, y, n, c, h * w, cp, (uint32_t)(pad_value & 0xff));
  return check_launch("u8_nchw_to_nhwc_kernel");
}

int i8ie_u8_nhwc_to_nchw(const uint8_t* x, uint8_t* y, int n, int c, int h, int w, int cp, void* stream) {
  I8IE_REQUIRE(cp >= c, "nhwc_to_nchw: bad cp");
  const int64_t items = (int64_t)n * c * h * w;
  if (items == 0) return I8IE_OK;
  u8_nhwc_to_nchw_kernel<<<stream_grid(items, kThreads), kThreads, 0, (cudaStream_t)stream>>>(x, y, n, c, h * w, cp);
  return check_launch("u8_nhwc_to_nchw_kernel");
}

// layer.cc:6-26 (host, once per model)
int i8ie_quantize_weight_host(const float* w, int64_t nw, const float* b, int64_t nb, int8_t* qw,
                              int8_t* qb, float* scale_host) {
  float mx = -FLT_MAX, mn = FLT_MAX;
  for (int64_t i = 0; i < nw; ++i) { mn = (w[i] < mn) ? w[i] : mn; mx = (mx < w[i]) ? w[i] : mx; }
  for (int64_t i = 0; i < nb; ++i) { mn = (b[i] < mn) ? b[i] : mn; mx = (mx < b[i]) ? b[i] : mx; }
  const float scale = (mx - mn) / 127;
  for (int64_t i = 0; i < nw; ++i) qw[i] = (int8_t)(int32_t)(w[i] / scale);
  for (int64_t i = 0; i < nb; ++i) qb[i] = (int8_t)(int32_t)(b[i] / scale);
  *scale_host = scale;
  return I8IE_OK;
}

int i8ie_zp_offsets(const int8_t* qw, const int8_t* qb, int n, int k, int in_zp, float in_scale,
                    int is_conv, int32_t* oc, float* bias_f, void* stream) {
  I8IE_REQUIRE(n > 0 && k > 0 && in_zp >= 0 && in_zp <= 255, "zp_offsets: bad shape/zp");
  const int rows_per_block = kThreads / 32;
  zp_offsets_kernel<<<(n + rows_per_block - 1) / rows_per_block, kThreads, 0, (cudaStream_t)stream>>>(
      qw, qb, n, k, in_zp, in_scale, is_conv, oc, bias_f);
  return check_launch("zp_offsets_kernel");
}

int i8ie_pack_conv_weight(const int8_t* qw_oihw, int8_t* w_packed, int kc, int c, int kh, int kw,
                          int kc_pad, int cp, void* stream) {
  I8IE_REQUIRE(kc_pad >= kc && cp >= c && cp % 16 == 0, "pack_conv_weight: bad padding");
  const int64_t items = (int64_t)kc_pad * kh * kw * cp;
  pack_conv_weight_kernel<<<stream_grid(items, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      qw_oihw, w_packed, kc, c, kh, kw, kc_pad, cp);
  return check_launch("pack_conv_weight_kernel");
}

}  // extern "C"
