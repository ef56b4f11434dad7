// Shared device/host helpers for libi8ie_sm100.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/i8ie_sm100.h"

namespace i8ie {

// ---- error plumbing (no exceptions cross the C ABI) -------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return I8IE_ECUDA;
  }
  return I8IE_OK;
}

#define I8IE_CUDA_OK(expr)                                                        \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::i8ie::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));          \
      return I8IE_ECUDA;                                                          \
    }                                                                             \
  } while (0)

#define I8IE_REQUIRE(cond, ...)             \
  do {                                      \
    if (!(cond)) {                          \
      ::i8ie::set_error(__VA_ARGS__);       \
      return I8IE_EINVAL;                   \
    }                                       \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// Kernels of the forward chain are launched with programmatic stream serialisation: a kernel may
// become resident while its predecessor drains, run its prologue (barrier init, TMEM allocation,
// tensor-map prefetch), and must call pdl_wait() before its first global-memory access — that
// returns once the predecessor grid has completed and its writes are visible. Every kernel
// launched through launch_pdl() calls pdl_launch_dependents() first thing and pdl_wait() before
// touching global memory, so dependencies stay transitive. Opt-in with I8IE_PDL=1.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Kernel families for the per-family PDL switch (I8IE_PDL_MASK = sum of the bits; I8IE_PDL=1 = all).
// A launch site tags itself with PdlFamily f(kPdl...) before calling launch_pdl / launch_cluster_pdl.
enum : int { kPdlStem = 1, kPdlPairConv = 2, kPdlPool = 4, kPdlFcCluster = 8, kPdlFcHead = 16, kPdlOther = 32, kPdlTc = 64,
       kPdlExchange = 128 };
inline int& pdl_family_slot() {
  static thread_local int f = kPdlOther;
  return f;
}
struct PdlFamily {
  int saved;
  explicit PdlFamily(int f) : saved(pdl_family_slot()) { pdl_family_slot() = f; }
  ~PdlFamily() { pdl_family_slot() = saved; }
};
inline int pdl_enabled() {   // measured on B200 (profiles/r02_stem_probes.md): which families gain, which lose
  // default: the tensor-core GEMM kernels and the classifier head (step 0.1931 -> 0.1828 ms at batch 100); pools lose
  // 3 us when they become resident early next to a conv CTA, the stem gains nothing (its predecessor is tiny)
  static const int mask = std::getenv("I8IE_PDL") != nullptr ? 0x7fffffff
                          : (std::getenv("I8IE_PDL_MASK") != nullptr ? std::atoi(std::getenv("I8IE_PDL_MASK"))
                                                                     : (kPdlPairConv | kPdlFcCluster | kPdlFcHead | kPdlTc | kPdlExchange));
  return (mask & pdl_family_slot()) != 0 ? 1 : 0;
}

template <typename... KArgs, typename... Args>
inline void launch_cluster_pdl(int cluster_x, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                               cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
  // cluster_x = -1: an EXPLICIT cluster of one CTA (kernels that use shared::cluster bulk copies within their own CTA)
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = (unsigned)(cluster_x < 0 ? 1 : cluster_x); at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = (cluster_x > 1 || cluster_x < 0) ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through check_launch()
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  launch_cluster_pdl(1, kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

// ---- the reference's arithmetic, restated with explicitly-rounded intrinsics ---
// (never contracted into FMAs, IEEE division, no FTZ) — SURVEY.md Appendix A.

// A4 down_scale (quantize_utils.cc:27-36).
__device__ __forceinline__ uint32_t requant_u8(int32_t acc, float sa, float sb, float sc, float zp_c) {
  const float dequant = __fmul_rn(__fmul_rn(__int2float_rn(acc), sa), sb);
  const float quant = __fadd_rn(__fdiv_rn(dequant, sc), zp_c);
  // (quant >= 255) ? 255 : (quant < 0) ? 0 : (u8)trunc(quant); NaN falls to the cast
  // (x86 cvttss2si gives 0x80000000 -> low byte 0; __float2int_rz(NaN) = 0).
  return (quant >= 255.f) ? 255u : (quant < 0.f) ? 0u : (uint32_t)__float2int_rz(quant);
}

// (r >= 255) ? 255 : (r < 0) ? 0 : (u8)trunc(r), NaN -> 0: one saturating convert (F2IP.U8.F32.TRUNC)
__device__ __forceinline__ uint32_t f32_to_u8_sat_rz(float r) {
  uint32_t y;
  asm("cvt.rzi.sat.u8.f32 %0, %1;" : "=r"(y) : "f"(r));
  return y;
}

// Same result as requant_u8 with the IEEE division d / sc replaced by the classic
// FMA-corrected quotient (rcp = RN(1/sc) hoisted out of the loop):
//   q0 = RN(d*rcp); e = RN(d - sc*q0) (exact); q = RN(q0 + e*rcp) == RN(d / sc)
// (Markstein's theorem: correctly rounded when rcp is the correctly rounded reciprocal, q0
// is faithful and nothing over/underflows — the host only selects this path when sc, sa*sb
// and the reachable |d| sit well inside the normal range, see requant_fast_ok()). Clamp and
// truncation are one saturating convert: identical to the reference's
// (q >= 255) ? 255 : (q < 0) ? 0 : (u8)q for every input incl. NaN -> 0.
__device__ __forceinline__ uint32_t requant_u8_fast(int32_t acc, float sa, float sb, float sc, float rcp,
                                                    float zp_c) {
  const float d = __fmul_rn(__fmul_rn(__int2float_rn(acc), sa), sb);
  const float q0 = __fmul_rn(d, rcp);
  const float e = __fmaf_rn(-sc, q0, d);
  const float q = __fmaf_rn(e, rcp, q0);
  return f32_to_u8_sat_rz(__fadd_rn(q, zp_c));
}

// ---- packed fp32x2 arithmetic (sm_100a FMUL2 / FFMA2 / FADD2) -------------------------------
// Two IEEE round-to-nearest fp32 operations per issued instruction; each half rounds exactly
// like the scalar __fmul_rn / __fmaf_rn / __fadd_rn, so results are bit-identical.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// Constants of the fast requantise, broadcast into both halves once per thread.
struct RequantFast2 {
  f32x2_t sa, sb, rcp, nsc, zp;
  float zpf;
};
__device__ __forceinline__ RequantFast2 make_requant_fast2(float sa, float sb, float sc, float rcp, float zpf) {
  RequantFast2 c;
  c.sa = f2_pack(sa, sa); c.sb = f2_pack(sb, sb); c.rcp = f2_pack(rcp, rcp); c.nsc = f2_pack(-sc, -sc);
  c.zp = f2_pack(zpf, zpf);
  c.zpf = zpf;
  return c;
}
// Two accumulators -> two u8 codes (same value as requant_u8_fast, then max(y, zp) if relu: the
// relu<u8> of functional.cc:22-23 is max(r, zp) before the monotone truncating convert).
template <bool RELU>
__device__ __forceinline__ void requant2_u8_fast(int32_t a0, int32_t a1, const RequantFast2& c, uint32_t& y0,
                                                 uint32_t& y1) {
  f32x2_t d = f2_pack(__int2float_rn(a0), __int2float_rn(a1));
  d = f2_mul(f2_mul(d, c.sa), c.sb);
  const f32x2_t q0 = f2_mul(d, c.rcp);
  const f32x2_t e = f2_fma(c.nsc, q0, d);
  const f32x2_t q = f2_fma(e, c.rcp, q0);
  float r0, r1;
  f2_unpack(f2_add(q, c.zp), r0, r1);
  if (RELU) { r0 = fmaxf(r0, c.zpf); r1 = fmaxf(r1, c.zpf); }
  y0 = f32_to_u8_sat_rz(r0);
  y1 = f32_to_u8_sat_rz(r1);
}

// Same with the weight scale given per element pair (per-output-channel scales, F4 extension).
template <bool RELU>
__device__ __forceinline__ void requant2_u8_fast_sb(int32_t a0, int32_t a1, const RequantFast2& c, f32x2_t sb2,
                                                    uint32_t& y0, uint32_t& y1) {
  f32x2_t d = f2_pack(__int2float_rn(a0), __int2float_rn(a1));
  d = f2_mul(f2_mul(d, c.sa), sb2);
  const f32x2_t q0 = f2_mul(d, c.rcp);
  const f32x2_t e = f2_fma(c.nsc, q0, d);
  const f32x2_t q = f2_fma(e, c.rcp, q0);
  float r0, r1;
  f2_unpack(f2_add(q, c.zp), r0, r1);
  if (RELU) { r0 = fmaxf(r0, c.zpf); r1 = fmaxf(r1, c.zpf); }
  y0 = f32_to_u8_sat_rz(r0);
  y1 = f32_to_u8_sat_rz(r1);
}

// FC's `C[i*n+j] += q_bias[j] / in.scale()` (fully_connected.cc:44): int += float.
__device__ __forceinline__ int32_t fc_bias_add(int32_t acc, float bias_f) {
  return __float2int_rz(__fadd_rn(__int2float_rn(acc), bias_f));
}

// A1 quantize element (quantize_utils.cc:49): (u8)(x/scale + zp), unclamped.
// The cast is x86's cvttss2si to int32, then the low byte: anything outside the int32 range
// (and NaN) converts to 0x80000000 -> byte 0. CUDA's convert saturates positive overflow to
// 0x7fffffff instead, so that case is mapped explicitly.
__device__ __forceinline__ uint32_t f2u8_wrap_x86(float v) {
  return (v >= 2147483648.f) ? 0u : ((uint32_t)__float2int_rz(v) & 0xffu);
}
__device__ __forceinline__ uint32_t quant_u8_wrap(float x, float scale, float zpf) {
  return f2u8_wrap_x86(__fadd_rn(__fdiv_rn(x, scale), zpf));
}
// Same result with x / scale computed as the FMA-corrected product with rcp = RN(1/scale)
// (Markstein, as in requant_u8_fast; the host selects this path only for a mid-range scale,
// see quant_fast_ok). Elements too large for the intermediates (|x| >= 1e18, inf, NaN) take
// the IEEE division; tiny ones give |quotient| < 2^-30 either way, which cannot change
// trunc(quotient + zp).
__device__ __forceinline__ uint32_t quant_u8_wrap_fast(float x, float scale, float rcp, float zpf) {
  const float q0 = __fmul_rn(x, rcp);
  const float e = __fmaf_rn(-scale, q0, x);
  float q = __fmaf_rn(e, rcp, q0);
  if (!(fabsf(x) < 1e18f)) q = __fdiv_rn(x, scale);
  return f2u8_wrap_x86(__fadd_rn(q, zpf));
}

// A5 dequantize element (quantize_utils.cc:40).
__device__ __forceinline__ float dequant_f32(uint32_t q, int zp, float scale) {
  return __fmul_rn(__int2float_rn((int)q - zp), scale);
}

__device__ __forceinline__ int32_t dp4a_u8s8(uint32_t a, uint32_t b, int32_t c) {
  int32_t d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Four elements at once, NO guards: valid when every |x| < quant_fast_limit(scale) (then all
// intermediates are normal and |x/scale + zp| < 2^31, so the x86 cast is a plain truncation);
// NaN falls through to byte 0 like the reference. Two packed fp32x2 chains + four converts.
struct QuantFast2 {
  f32x2_t rcp, nscale, zp;
};
__device__ __forceinline__ QuantFast2 make_quant_fast2(float scale, float rcp, float zpf) {
  QuantFast2 c;
  c.rcp = f2_pack(rcp, rcp); c.nscale = f2_pack(-scale, -scale); c.zp = f2_pack(zpf, zpf);
  return c;
}
__device__ __forceinline__ void quant2_fast(float x0, float x1, const QuantFast2& c, int& i0, int& i1) {
  const f32x2_t x = f2_pack(x0, x1);
  const f32x2_t q0 = f2_mul(x, c.rcp);
  const f32x2_t e = f2_fma(c.nscale, q0, x);
  const f32x2_t q = f2_fma(e, c.rcp, q0);
  float v0, v1;
  f2_unpack(f2_add(q, c.zp), v0, v1);
  i0 = __float2int_rz(v0);
  i1 = __float2int_rz(v1);
}
// low bytes of four ints -> one word
__device__ __forceinline__ uint32_t pack_low_bytes(int i0, int i1, int i2, int i3) {
  return __byte_perm(__byte_perm((uint32_t)i0, (uint32_t)i1, 0x0040), __byte_perm((uint32_t)i2, (uint32_t)i3, 0x0040), 0x5410);
}
__device__ __forceinline__ uint32_t quant4_fast(float a, float b, float c, float d, const QuantFast2& k) {
  int i0, i1, i2, i3;
  quant2_fast(a, b, k, i0, i1);
  quant2_fast(c, d, k, i2, i3);
  return pack_low_bytes(i0, i1, i2, i3);
}

// streaming (read-once) 128-bit loads / stores
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// read-once input that should not displace the L2-resident working set (weights, activations)
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float2 ld_evict_first_f2(const void* p, uint64_t pol) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;"
               : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u32(void* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream_f4(void* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Host-side guard for requant_u8_fast: all scales positive, finite and comfortably normal, and
// the largest reachable |d| = 2^31*sa*sb as well as d/sc far from overflow/underflow.
inline bool requant_fast_ok(float sa, float sb, float sc) {
  auto mid = [](float v) { return v > 1e-18f && v < 1e18f; };
  if (!(mid(sa) && mid(sb) && mid(sc))) return false;
  uint32_t bits;
  memcpy(&bits, &sc, sizeof(bits));
  if ((bits & 0x7fffffu) == 0x7fffffu) return false;   // Markstein's excluded divisor (significand all ones)
  const double dmax = 2147483648.0 * (double)sa * (double)sb;
  const double dmin = (double)sa * (double)sb;
  return dmax < 1e30 && dmin > 1e-30 && dmax / sc < 1e30 && dmin / sc > 1e-30;
}

// Host-side guard for quant_u8_wrap_fast.
inline bool quant_fast_ok(float scale) {
  if (!(scale > 1e-18f && scale < 1e18f)) return false;
  uint32_t bits;
  memcpy(&bits, &scale, sizeof(bits));
  return (bits & 0x7fffffu) != 0x7fffffu;
}

// Largest |x| for which the unguarded packed quantise (quant4_fast) is exact; 0 = never use it.
inline float quant_fast_limit(float scale) {
  if (!quant_fast_ok(scale)) return 0.f;
  const float lim = 1073741824.f * scale;   // |x / scale| < 2^30, so |x/scale + zp| < 2^31
  return lim < 1e18f ? lim : 1e18f;
}

// epilogue parameters shared by the SIMT and tcgen05 GEMM-shaped kernels
struct EpiParams {
  const int32_t* oc;      // [n] zero-point(+bias for conv) offsets
  const float* bias_f;    // [n] fc float bias term, or nullptr (conv)
  float sa, sb, sc;       // in.scale, weight scale, out scale
  int zp_out;             // out zero point
  int relu;               // fuse relu<u8>
  int32_t* acc_out;       // optional [m, n] s32 dump (parity tests)
  // F4 extension (opt-in, not in the reference): per-output-channel weight scales. sb_vec = device [n]
  // (then `sb` is ignored and [sb_min, sb_max] bounds its entries for the fast-path guard); nullptr = the
  // reference's single per-tensor scale (layer.cc:18-19).
  const float* sb_vec = nullptr;
  float sb_min = 0.f, sb_max = 0.f;
  // fc + dequantize in one launch (i8ie_fc_u8_deq): dense [m, n] fp32 = dequantize(y), i.e.
  // ((float)y - zp_out) * sc (quantize_utils.cc:38-42). Honoured by fc_head_kernel; the dispatcher
  // runs the standalone dequantise after any other kernel.
  float* deq_out = nullptr;
};

// requant_fast_ok for an epilogue with either kind of weight scale
inline bool requant_fast_ok(const EpiParams& ep) {
  if (ep.sb_vec == nullptr) return requant_fast_ok(ep.sa, ep.sb, ep.sc);
  return requant_fast_ok(ep.sa, ep.sb_min, ep.sc) && requant_fast_ok(ep.sa, ep.sb_max, ep.sc);
}

}  // namespace i8ie
