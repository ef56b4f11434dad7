// Internal interface between the C-ABI entry points (gemm_api.cu) and the
// GEMM-shaped kernels (simt_gemm.cu, tc_gemm.cu).
#pragma once
#include "common.cuh"

namespace i8ie {

// Geometry of one implicit GEMM: D[M, N] = A[M, K] * W[N, K]^T with
// M = n*oh*ow output pixels, N = kc output channels, K = kh*kw*cp.
struct GemmGeom {
  int n, h, w, cp;          // input NHWC, channel pitch cp (multiple of 16)
  int kh, kw, stride, pad;
  int oh, ow;
  int M, N;                 // GEMM rows / real output channels
  int n_pad;                // rows present in the packed weight
  int ldw;                  // packed weight row pitch in bytes (= kh*kw*cp for conv)
  int out_cp;               // output channel pitch (multiple of 16)
};

int launch_simt_igemm(const GemmGeom& g, const uint8_t* x, const int8_t* w, uint8_t* y,
                      const EpiParams& ep, int zp_in, cudaStream_t stream);

}  // namespace i8ie
