/* Stand-in for Intel MKL's <mkl.h>: TEST INFRASTRUCTURE ONLY.
 *
 * The reference (t0037799/INT8InferenceEngine) includes "mkl.h" at
 * include/layer.h:3 and uses exactly two CBLAS entry points from it:
 *   cblas_sgemm          conv2d.cc:83, fully_connected.cc:10
 *   cblas_gemm_s8u8s32   conv2d.cc:131, fully_connected.cc:39
 * Intel MKL (pinned by the reference to 2019.5.281, CMakeLists.txt:25) is not
 * installed in this image and cannot be fetched, so this header declares those
 * two functions (and the CBLAS enums their call sites name) and
 * oracle/shim/mkl_shim.cc implements the one call pattern the reference uses.
 * The integer GEMM is exact s32 arithmetic, so any correct implementation gives
 * MKL's bits; the fp32 GEMM is tolerance-only (not on the INT8 path).
 */
#ifndef I8IE_ORACLE_MKL_SHIM_H
#define I8IE_ORACLE_MKL_SHIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MKL_INT int
#define MKL_INT8 int8_t
#define MKL_INT32 int32_t

typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasRowOffset = 171, CblasColOffset = 172, CblasFixOffset = 173 } CBLAS_OFFSET;

void cblas_sgemm(const CBLAS_LAYOUT layout, const CBLAS_TRANSPOSE transa,
                 const CBLAS_TRANSPOSE transb, const MKL_INT m, const MKL_INT n,
                 const MKL_INT k, const float alpha, const float* a,
                 const MKL_INT lda, const float* b, const MKL_INT ldb,
                 const float beta, float* c, const MKL_INT ldc);

/* Row-major NoTrans x Trans as the reference calls it: A is u8 [m,k],
 * B is s8 [n,k], C s32 [m,n]; with CblasRowOffset co has n entries and
 * co[j] is added to every row (MKL's row-major convention). */
void cblas_gemm_s8u8s32(const CBLAS_LAYOUT layout, const CBLAS_TRANSPOSE transa,
                        const CBLAS_TRANSPOSE transb, const CBLAS_OFFSET offsetc,
                        const MKL_INT m, const MKL_INT n, const MKL_INT k,
                        const float alpha, const void* a, const MKL_INT lda,
                        const MKL_INT8 ao, const void* b, const MKL_INT ldb,
                        const MKL_INT8 bo, const float beta, MKL_INT32* c,
                        const MKL_INT ldc, const MKL_INT32* co);

#ifdef __cplusplus
}
#endif
#endif
