// Internal interface between the C-ABI entry points (gemm_api.cu) and the
// GEMM-shaped kernels (simt_gemm.cu, tc_gemm.cu).
#pragma once
#include <cuda.h>

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

// classifier-head fc (n_pad == 16): one warp per row, see simt_gemm.cu
bool fc_head_eligible(int n_pad, int ldx, int ldw, int ldy, const void* x, const void* w);
int launch_fc_head(const uint8_t* x, int ldx, const int8_t* w, int ldw, uint8_t* y, int ldy, int m, int n, int k,
                   const EpiParams& ep, cudaStream_t stream);

// ---- tcgen05 path (tc_gemm.cu) ---------------------------------------------------------
bool tc_conv_eligible(const GemmGeom& g);
int tc_conv_bk(const GemmGeom& g);
int tc_pick_bn(int n);
int tc_border_table_size(const GemmGeom& g);   // entries (int32) of the zero-point border table
int tc_build_border_table(const GemmGeom& g, const int8_t* w_packed, int32_t* tab, cudaStream_t stream);
int tc_encode_weight_map(CUtensorMap* tm, const int8_t* w, int rows, int ldw, int bk, int bn);
int tc_encode_act_map_im2col(CUtensorMap* tm, const uint8_t* x, const GemmGeom& g, int bk);
int tc_encode_act_map_rows(CUtensorMap* tm, const uint8_t* x, int m, int k, int ldx);
// row mode: physically padded input, K = (filter row, byte of the kw * cp run) — see tc_gemm.cu
int tc_row_mode_cp(int c, int cp_plain, int kh, int kw, int stride, int pad, int out_cp);
bool tc_row_mode_ok(int c, int kh, int kw, int stride, int pad, int out_cp);
int tc_row_mode_kr(int cpx, int kw);
int tc_row_pack_weights(const GemmGeom& g, const int8_t* w_packed, int8_t* wr, int kr, cudaStream_t stream);
int tc_encode_act_map_row_mode(CUtensorMap* tm, const uint8_t* x, const GemmGeom& g, int kr);
GemmGeom tc_row_mode_geom(const GemmGeom& g, int kr);
int tc_conv_cluster(int bk, int bn);   // 2 = CTA-pair kernel (cta_group::2), 1 = single CTAs
int tc_pair_box_rows(int bn);          // weight-map box rows of the pair kernel
int tc_pick_bn_pair(const GemmGeom& g, int bk, int* mt_out);   // cost-model tile for conv plans: N width, accumulators per tile
int launch_tc_conv(const GemmGeom& g, const CUtensorMap& tmA, const CUtensorMap& tmB, int bk, int bn, int cluster,
                   const int32_t* border_tab, uint8_t* y, const EpiParams& ep, int zp_in, cudaStream_t stream,
                   int pair_mt = 1);
// cluster = 1: K split over a 4-CTA cluster folded through distributed shared memory (tc_fc_cluster_kernel)
void tc_fc_config(int m, int ldy, int k, int* bn, int* splits, int* kb_per, int* cluster);
// w_tiled: optional pre-swizzled 128 x 128-byte block copy of the weights (tc_fc_tile_weights), or nullptr
int launch_tc_fc(int m, int n, int k, int ldy, const CUtensorMap& tmA, const CUtensorMap& tmB, int bn, int splits,
                 int kb_per, uint8_t* y, const EpiParams& ep, cudaStream_t stream, const int8_t* w_tiled = nullptr,
                 int ldw = 0, int cluster = 0);
int tc_fc_tile_weights(const int8_t* w, int n_pad, int ldw, int8_t* wt, cudaStream_t stream);
int tc_read_error(int* out, bool reset);
int tc_error_sink_init();         // host-mapped mirror of the protocol-error flag (call outside stream capture)
void tc_error_sink_touch(cudaStream_t stream);   // launch paths: lazily creates the sink unless capturing
int tc_error_poll(bool reset);    // reads the mirror: no CUDA call, valid after a synchronisation

// stem path (small-C strided first-layer convs), see tc_gemm.cu
struct StemGeom {
  int c;          // real input channels (<= 4)
  int hp, wsp;    // rows / 16-byte superpixels per row of the bordered stem image
  int64_t bytes;  // size of the stem buffer for the plan's batch
};
bool tc_stem_eligible(const GemmGeom& g, int c);
StemGeom tc_stem_geom(const GemmGeom& g, int c);
int tc_stem_pack_weights(const GemmGeom& g, int c, const int8_t* w_packed, int8_t* ws, cudaStream_t stream);
int tc_stem_pack_input(const GemmGeom& g, const StemGeom& s, const uint8_t* x, uint8_t* xs, int zp, cudaStream_t stream);
int tc_stem_quantize_input(const GemmGeom& g, const StemGeom& s, const float* x, const float* const* xslot,
                           uint8_t* xs, float scale, int zp,
                           cudaStream_t stream);
int tc_encode_stem_act_map(CUtensorMap* tm, const uint8_t* xs, const GemmGeom& g, const StemGeom& s);
bool tc_stem2_eligible(const GemmGeom& g, int c);
// fp32 NCHW source of a fused input quantise (direct pointer or run-time address slot)
struct StemF32Src {
  const float* x;
  const float* const* xslot;
  float scale;
  int zp;
};
bool tc_stem2_can_fuse_quantize(const GemmGeom& g, int c, const float* x, const float* const* xslot, float scale);   // RGB stems
// f32 == nullptr: rows come from the bordered u8 stem image xs; otherwise the kernel quantises them itself
// k48: tmB maps the K-packed weights (tc_stem_pack_weights48: 48-byte windows, see Stem2Params::k48_n)
int launch_tc_stem2(const GemmGeom& g, const StemGeom& s, const uint8_t* xs, const StemF32Src* f32,
                    const CUtensorMap& tmB, int bn, uint8_t* y, const EpiParams& ep, cudaStream_t stream, bool k48 = false);
int tc_stem_k48_bytes(const GemmGeom& g, int c);
int tc_stem_pack_weights48(const GemmGeom& g, int c, const int8_t* w_packed, int8_t* ws, cudaStream_t stream);
int launch_tc_stem(const GemmGeom& g, const CUtensorMap& tmA, const CUtensorMap& tmB, int bn, uint8_t* y,
                   const EpiParams& ep, cudaStream_t stream);

}  // namespace i8ie
