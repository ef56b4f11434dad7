// Stand-in implementation of the two MKL CBLAS calls the reference uses.
// TEST INFRASTRUCTURE ONLY (see oracle/shim/mkl.h). Only the call pattern the
// reference actually issues is supported (conv2d.cc:83,131;
// fully_connected.cc:10,39): RowMajor, NoTrans x Trans, alpha=1, beta=0,
// ao=bo=0, CblasRowOffset. Anything else aborts loudly.
//
// The s8u8s32 GEMM is exact integer arithmetic (MKL's is non-saturating s32 as
// well), so results are bit-identical to MKL by construction. It is vectorised
// with AVX512-VNNI when the host has it (vpdpbusd has the same u8 x s8
// signedness and does not saturate) so that the CPU baseline timing is not an
// artificially slow scalar loop; the scalar loop is the fallback.
// Threading mirrors MKL's default: serial when called from inside an OpenMP
// parallel region (conv2d.cc:125 calls it once per image per thread), threaded
// when called from the main thread (fully_connected.cc:39).
#include "mkl.h"

#include <immintrin.h>
#include <omp.h>

#include <cstdio>
#include <cstdlib>

namespace {

[[noreturn]] void unsupported(const char* what) {
  std::fprintf(stderr, "mkl_shim: unsupported call pattern: %s\n", what);
  std::abort();
}

bool have_vnni() {
  static const bool ok = __builtin_cpu_supports("avx512f") &&
                         __builtin_cpu_supports("avx512bw") &&
                         __builtin_cpu_supports("avx512vnni");
  return ok && std::getenv("I8IE_SHIM_SCALAR") == nullptr;
}

bool have_avx512f() {
  static const bool ok = __builtin_cpu_supports("avx512f");
  return ok && std::getenv("I8IE_SHIM_SCALAR") == nullptr;
}

// ---- integer GEMM --------------------------------------------------------

void igemm_block_scalar(const uint8_t* a, int lda, const int8_t* b, int ldb,
                        int32_t* c, int ldc, const int32_t* co, int i0, int i1,
                        int j0, int j1, int k) {
  for (int i = i0; i < i1; ++i) {
    const uint8_t* ar = a + (size_t)i * lda;
    for (int j = j0; j < j1; ++j) {
      const int8_t* br = b + (size_t)j * ldb;
      int32_t s = 0;
      for (int p = 0; p < k; ++p) s += (int32_t)ar[p] * (int32_t)br[p];
      c[(size_t)i * ldc + j] = s + co[j];
    }
  }
}

template <int MI, int NJ>
__attribute__((target("avx512f,avx512bw,avx512vnni"))) inline void igemm_tile_vnni(
    const uint8_t* a, int lda, const int8_t* b, int ldb, int32_t* c, int ldc,
    const int32_t* co, int i, int j, int k) {
  __m512i acc[MI][NJ];
  for (int x = 0; x < MI; ++x)
    for (int y = 0; y < NJ; ++y) acc[x][y] = _mm512_setzero_si512();
  const int kfull = k & ~63;
  for (int p = 0; p < kfull; p += 64) {
    __m512i av[MI], bv[NJ];
    for (int x = 0; x < MI; ++x)
      av[x] = _mm512_loadu_si512((const void*)(a + (size_t)(i + x) * lda + p));
    for (int y = 0; y < NJ; ++y)
      bv[y] = _mm512_loadu_si512((const void*)(b + (size_t)(j + y) * ldb + p));
    for (int x = 0; x < MI; ++x)
      for (int y = 0; y < NJ; ++y)
        acc[x][y] = _mm512_dpbusd_epi32(acc[x][y], av[x], bv[y]);
  }
  if (kfull < k) {
    const __mmask64 msk = (~0ULL) >> (64 - (k - kfull));
    __m512i av[MI], bv[NJ];
    for (int x = 0; x < MI; ++x)
      av[x] = _mm512_maskz_loadu_epi8(msk, a + (size_t)(i + x) * lda + kfull);
    for (int y = 0; y < NJ; ++y)
      bv[y] = _mm512_maskz_loadu_epi8(msk, b + (size_t)(j + y) * ldb + kfull);
    for (int x = 0; x < MI; ++x)
      for (int y = 0; y < NJ; ++y)
        acc[x][y] = _mm512_dpbusd_epi32(acc[x][y], av[x], bv[y]);
  }
  for (int x = 0; x < MI; ++x)
    for (int y = 0; y < NJ; ++y)
      c[(size_t)(i + x) * ldc + (j + y)] =
          _mm512_reduce_add_epi32(acc[x][y]) + co[j + y];
}

__attribute__((target("avx512f,avx512bw,avx512vnni"))) void igemm_block_vnni(
    const uint8_t* a, int lda, const int8_t* b, int ldb, int32_t* c, int ldc,
    const int32_t* co, int i0, int i1, int j0, int j1, int k) {
  int i = i0;
  for (; i + 4 <= i1; i += 4) {
    int j = j0;
    for (; j + 4 <= j1; j += 4) igemm_tile_vnni<4, 4>(a, lda, b, ldb, c, ldc, co, i, j, k);
    for (; j < j1; ++j) igemm_tile_vnni<4, 1>(a, lda, b, ldb, c, ldc, co, i, j, k);
  }
  for (; i < i1; ++i) {
    int j = j0;
    for (; j + 4 <= j1; j += 4) igemm_tile_vnni<1, 4>(a, lda, b, ldb, c, ldc, co, i, j, k);
    for (; j < j1; ++j) igemm_tile_vnni<1, 1>(a, lda, b, ldb, c, ldc, co, i, j, k);
  }
}

// ---- fp32 GEMM (tolerance-only path) ---------------------------------------

void sgemm_block_scalar(const float* a, int lda, const float* b, int ldb,
                        float* c, int ldc, int i0, int i1, int j0, int j1, int k) {
  for (int i = i0; i < i1; ++i) {
    const float* ar = a + (size_t)i * lda;
    for (int j = j0; j < j1; ++j) {
      const float* br = b + (size_t)j * ldb;
      float s = 0.f;
      for (int p = 0; p < k; ++p) s += ar[p] * br[p];
      c[(size_t)i * ldc + j] = s;
    }
  }
}

template <int MI, int NJ>
__attribute__((target("avx512f"))) inline void sgemm_tile_avx512(
    const float* a, int lda, const float* b, int ldb, float* c, int ldc, int i,
    int j, int k) {
  __m512 acc[MI][NJ];
  for (int x = 0; x < MI; ++x)
    for (int y = 0; y < NJ; ++y) acc[x][y] = _mm512_setzero_ps();
  const int kfull = k & ~15;
  for (int p = 0; p < kfull; p += 16) {
    __m512 av[MI], bv[NJ];
    for (int x = 0; x < MI; ++x) av[x] = _mm512_loadu_ps(a + (size_t)(i + x) * lda + p);
    for (int y = 0; y < NJ; ++y) bv[y] = _mm512_loadu_ps(b + (size_t)(j + y) * ldb + p);
    for (int x = 0; x < MI; ++x)
      for (int y = 0; y < NJ; ++y) acc[x][y] = _mm512_fmadd_ps(av[x], bv[y], acc[x][y]);
  }
  if (kfull < k) {
    const __mmask16 msk = (__mmask16)((1u << (k - kfull)) - 1u);
    __m512 av[MI], bv[NJ];
    for (int x = 0; x < MI; ++x)
      av[x] = _mm512_maskz_loadu_ps(msk, a + (size_t)(i + x) * lda + kfull);
    for (int y = 0; y < NJ; ++y)
      bv[y] = _mm512_maskz_loadu_ps(msk, b + (size_t)(j + y) * ldb + kfull);
    for (int x = 0; x < MI; ++x)
      for (int y = 0; y < NJ; ++y) acc[x][y] = _mm512_fmadd_ps(av[x], bv[y], acc[x][y]);
  }
  for (int x = 0; x < MI; ++x)
    for (int y = 0; y < NJ; ++y)
      c[(size_t)(i + x) * ldc + (j + y)] = _mm512_reduce_add_ps(acc[x][y]);
}

__attribute__((target("avx512f"))) void sgemm_block_avx512(
    const float* a, int lda, const float* b, int ldb, float* c, int ldc, int i0,
    int i1, int j0, int j1, int k) {
  int i = i0;
  for (; i + 4 <= i1; i += 4) {
    int j = j0;
    for (; j + 4 <= j1; j += 4) sgemm_tile_avx512<4, 4>(a, lda, b, ldb, c, ldc, i, j, k);
    for (; j < j1; ++j) sgemm_tile_avx512<4, 1>(a, lda, b, ldb, c, ldc, i, j, k);
  }
  for (; i < i1; ++i) {
    int j = j0;
    for (; j + 4 <= j1; j += 4) sgemm_tile_avx512<1, 4>(a, lda, b, ldb, c, ldc, i, j, k);
    for (; j < j1; ++j) sgemm_tile_avx512<1, 1>(a, lda, b, ldb, c, ldc, i, j, k);
  }
}

// Splits [0,m)x[0,n) into blocks and runs fn on each; threaded only when not
// already inside a parallel region.
template <typename Fn>
void for_blocks(int m, int n, Fn fn) {
  const int BI = 32, BJ = 32;
  const int nbi = (m + BI - 1) / BI, nbj = (n + BJ - 1) / BJ;
  const int nb = nbi * nbj;
  if (omp_in_parallel() || nb == 1) {
    for (int t = 0; t < nb; ++t) {
      const int bi = t / nbj, bj = t % nbj;
      fn(bi * BI, (bi * BI + BI < m) ? bi * BI + BI : m, bj * BJ,
         (bj * BJ + BJ < n) ? bj * BJ + BJ : n);
    }
  } else {
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < nb; ++t) {
      const int bi = t / nbj, bj = t % nbj;
      fn(bi * BI, (bi * BI + BI < m) ? bi * BI + BI : m, bj * BJ,
         (bj * BJ + BJ < n) ? bj * BJ + BJ : n);
    }
  }
}

}  // namespace

extern "C" void cblas_gemm_s8u8s32(const CBLAS_LAYOUT layout, const CBLAS_TRANSPOSE transa,
                                   const CBLAS_TRANSPOSE transb, const CBLAS_OFFSET offsetc,
                                   const MKL_INT m, const MKL_INT n, const MKL_INT k,
                                   const float alpha, const void* a, const MKL_INT lda,
                                   const MKL_INT8 ao, const void* b, const MKL_INT ldb,
                                   const MKL_INT8 bo, const float beta, MKL_INT32* c,
                                   const MKL_INT ldc, const MKL_INT32* co) {
  if (layout != CblasRowMajor || transa != CblasNoTrans || transb != CblasTrans ||
      offsetc != CblasRowOffset || alpha != 1.0f || beta != 0.0f || ao != 0 || bo != 0)
    unsupported("cblas_gemm_s8u8s32 (only RowMajor NoTrans x Trans, RowOffset, alpha=1, beta=0, ao=bo=0)");
  const uint8_t* A = static_cast<const uint8_t*>(a);
  const int8_t* B = static_cast<const int8_t*>(b);
  const bool vnni = have_vnni();
  for_blocks(m, n, [&](int i0, int i1, int j0, int j1) {
    if (vnni)
      igemm_block_vnni(A, lda, B, ldb, c, ldc, co, i0, i1, j0, j1, k);
    else
      igemm_block_scalar(A, lda, B, ldb, c, ldc, co, i0, i1, j0, j1, k);
  });
}

extern "C" void cblas_sgemm(const CBLAS_LAYOUT layout, const CBLAS_TRANSPOSE transa,
                            const CBLAS_TRANSPOSE transb, const MKL_INT m, const MKL_INT n,
                            const MKL_INT k, const float alpha, const float* a,
                            const MKL_INT lda, const float* b, const MKL_INT ldb,
                            const float beta, float* c, const MKL_INT ldc) {
  if (layout != CblasRowMajor || transa != CblasNoTrans || transb != CblasTrans ||
      alpha != 1.0f || beta != 0.0f)
    unsupported("cblas_sgemm (only RowMajor NoTrans x Trans, alpha=1, beta=0)");
  const bool avx = have_avx512f();
  for_blocks(m, n, [&](int i0, int i1, int j0, int j1) {
    if (avx)
      sgemm_block_avx512(a, lda, b, ldb, c, ldc, i0, i1, j0, j1, k);
    else
      sgemm_block_scalar(a, lda, b, ldb, c, ldc, i0, i1, j0, j1, k);
  });
}

// Lets the bench report which integer kernel the CPU baseline actually ran.
extern "C" int i8ie_shim_uses_vnni(void) { return have_vnni() ? 1 : 0; }
