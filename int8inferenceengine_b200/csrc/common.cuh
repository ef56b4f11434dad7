// Shared device/host helpers for libi8ie_sm100.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
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

// Same result as requant_u8 with the IEEE division d / sc replaced by the classic
// FMA-corrected quotient (rcp = RN(1/sc) hoisted out of the loop):
//   q0 = RN(d*rcp); e = RN(d - sc*q0) (exact); q = RN(q0 + e*rcp) == RN(d / sc)
// (Markstein's theorem: correctly rounded when rcp is the correctly rounded reciprocal, q0
// is faithful and nothing over/underflows — the host only selects this path when sc, sa*sb
// and the reachable |d| sit well inside the normal range, see requant_fast_ok()). The clamp
// is done in fp32 before the truncating convert: identical to the reference's
// (q >= 255) ? 255 : (q < 0) ? 0 : (u8)q for every input incl. NaN -> 0.
__device__ __forceinline__ uint32_t requant_u8_fast(int32_t acc, float sa, float sb, float sc, float rcp,
                                                    float zp_c) {
  const float d = __fmul_rn(__fmul_rn(__int2float_rn(acc), sa), sb);
  const float q0 = __fmul_rn(d, rcp);
  const float e = __fmaf_rn(-sc, q0, d);
  const float q = __fmaf_rn(e, rcp, q0);
  const float r = fminf(fmaxf(__fadd_rn(q, zp_c), 0.f), 255.f);
  return __float2uint_rz(r);
}

// FC's `C[i*n+j] += q_bias[j] / in.scale()` (fully_connected.cc:44): int += float.
__device__ __forceinline__ int32_t fc_bias_add(int32_t acc, float bias_f) {
  return __float2int_rz(__fadd_rn(__int2float_rn(acc), bias_f));
}

// A1 quantize element (quantize_utils.cc:49): (u8)(x/scale + zp), unclamped.
__device__ __forceinline__ uint32_t quant_u8_wrap(float x, float scale, float zpf) {
  return (uint32_t)__float2int_rz(__fadd_rn(__fdiv_rn(x, scale), zpf)) & 0xffu;
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

// epilogue parameters shared by the SIMT and tcgen05 GEMM-shaped kernels
struct EpiParams {
  const int32_t* oc;      // [n] zero-point(+bias for conv) offsets
  const float* bias_f;    // [n] fc float bias term, or nullptr (conv)
  float sa, sb, sc;       // in.scale, weight scale, out scale
  int zp_out;             // out zero point
  int relu;               // fuse relu<u8>
  int32_t* acc_out;       // optional [m, n] s32 dump (parity tests)
};

}  // namespace i8ie
