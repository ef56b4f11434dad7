// tcgen05 (5th-gen tensor core) implicit-GEMM kernel for the i8ie INT8 path.
//
//   D[M, N] (s32, TMEM) = A[M, K] (u8) * W[N, K]^T (s8),  K-major operands in shared memory
//
// conv2d (conv2d.cc:100-142): A is never materialised. The TMA engine gathers it straight
//   from the NHWC activation tensor in IM2COL mode — one box = 128 consecutive output pixels
//   (walking W, then H, then N exactly like the reference's row index ti*ow+tj) x BK channels
//   of one filter tap — and zero-fills spatially out-of-range taps. The reference pads with
//   the input zero_point instead (conv2d.cc:24-25); the exact integer difference
//   zp * sum_{out-of-range taps} w is added back in the epilogue from a per-border-class
//   table, so the s32 accumulators are bit-identical.
// fully_connected (fully_connected.cc:22-52): A is a plain 2-D TMA tile of the [M, K] rows.
//
// Warp-specialised, one 128 x BN output tile per CTA:
//   warp 0     TMA producer (one elected lane), STAGES-deep mbarrier ring
//   warp 1     TMEM allocator + MMA issuer (one elected lane): tcgen05.mma kind::i8, M=128, N=BN, K=32
//   warps 2-5  epilogue: tcgen05.ld -> + oc (+ border correction) [-> fc float bias] -> requantise
//              (the reference's exact fp32 sequence) -> [relu] -> packed u8 NHWC store
// Several CTAs are co-resident per SM (TMEM: BN columns each), so one CTA's epilogue
// overlaps another's main loop.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "gemm_api.cuh"
#include "tc_ptx.cuh"

namespace i8ie {

__device__ int g_tc_error = 0;  // first protocol error seen by any tensor-core kernel (0 = none)
// Host-mapped mirror of g_tc_error (pinned, zero-copy): the host reads it after any synchronising
// call (numpy(), the end of a timed region, smoke()) WITHOUT another CUDA call, so a barrier timeout
// surfaces as an exception at the next result read instead of as silently wrong activations.
__device__ int* g_tc_error_host = nullptr;

// A bounded mbarrier wait expired (see ptx::mbar_wait): record the role that timed out.
//   1 producer waiting for a free operand stage   2 MMA issuer waiting for operands
//   3 epilogue waiting for an accumulator         4 MMA issuer waiting for a drained accumulator
//   5 stem MMA issuer waiting for the weights     6 stem converter waiting for fp32 rows
__device__ __noinline__ void tc_fail(int code) {
  atomicCAS(&g_tc_error, 0, code);
  int* h = g_tc_error_host;
  if (h != nullptr) {
    *reinterpret_cast<volatile int*>(h) = code;
    __threadfence_system();
  }
}

struct TcParams {
  int M, N, out_cp;
  int tiles_m, tiles_n;
  int num_kb;      // pipeline stages consumed per tile (each = ksub sub-blocks of BK bytes)
  int ksub;        // sub-blocks (TMA boxes per operand) per stage
  int stages;      // ring depth
  int cblocks;     // BK-byte channel blocks per filter tap (im2col)
  int kh, kw, stride_h, stride_w, pad, H, W, oh, ow;
  int zp_in;
  int fast_requant;           // requant_fast_ok(sa, sb, sc)
  const int32_t* border_tab;  // [(pad+1)^4][N] or nullptr
  uint8_t* y;
  EpiParams ep;
  // split-K (small-M fc): each tile covers kb_per stages of K and dumps raw s32 partials
  int splits, kb_per;
  int32_t* ws;                // [splits][M][ws_ld]
  int ws_ld;
  // 128-row sub-tiles per CTA tile (1 or 2): two accumulators share every weight stage, which
  // cuts the L2->SM bytes per MAC (the binding limit of a 128 x BN tile, ~43 B/clk/SM)
  int mt;
  // CTA-pair kernel: tiles_mp = ceil(tiles_m / 2) 256-row pair tiles; single-CTA kernel: tiles_mp = tiles_m
  int tiles_mp;
  // fc weight stream (MODE 0, BN a multiple of 128): copy of the weights stored as [n / 128][k / 128] blocks of
  // 128 rows x 128 bytes, each block already in the 128-byte-swizzled shared-memory image, so a weight stage is
  // ONE contiguous 16 KB bulk copy per 128 rows instead of a tiled box of 128-byte rows 'ldw' apart. Measured
  // (tools/ubench/i8_peak.cu, dram_bulk): contiguous 16 KB chunks stream from DRAM at 6.0 TB/s chip-wide, where
  // the row-strided boxes of the same weights reached 2.3 TB/s. nullptr = tiled map tmB.
  const int8_t* w_tiled;
  int wt_nkb;                 // 128-byte K blocks per row of the tiled copy (= ldw / 128)
  int dbg;                    // dev-only probes of the cluster fc kernel (I8IE_FC_DBG): 1 = no push, 2 = no fold, 4 = no loads
};

namespace {

constexpr int BM = 128;
// Four epilogue warps per TMEM lane quadrant (warp % 4), taking every fourth 32-column chunk. One 32 x 32
// chunk costs a warp ~1000 clk (TMEM load, ~250 dependent-ish instructions, two stores): with 8 warps the
// epilogue of a 128 x 256 tile (4000 clk) held its accumulator long enough to stall the MMA issuer and left a
// 2 us tail per launch; 16 warps took AlexNet conv2 at batch 100 from 40 to 32 us.
constexpr int kEpiWarps = 16;
constexpr int kThreads = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA, then the epilogue warps
constexpr int kEpiWarps2 = kEpiWarps;   // (pair kernel: same count)
constexpr int kThreads2 = 64 + 32 * kEpiWarps2;
__device__ __forceinline__ void epi_bar_sync2() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps2) : "memory"); }
// stem kernel: its K is tiny (22 MMAs per tile), so the fp32 epilogue is the longest stage; with two
// warps per sub-partition it runs latency-bound, four warps per sub-partition hide the dependent chains
constexpr int kStemEpiWarps = 16;
// fused-quantise variant: 8 producer warps (2 rows of a 15/16-row tile each) next to the epilogue
// warps — one epilogue warp per (quadrant, 32-column chunk) for the 96-channel tile, else 2 per quadrant
constexpr int kStemProdWarps = 8;
constexpr int kStemFBar = 8;   // first barrier slot of the fp32 row ring (operand ring has <= 6 stages)
template <int BN, bool FQ>
constexpr int stem_epi_warps() { return !FQ ? kStemEpiWarps : (BN == 96 ? 12 : 8); }
template <int BN, bool FQ>
constexpr int stem_threads() { return 64 + 32 * (stem_epi_warps<BN, FQ>() + (FQ ? kStemProdWarps : 0)); }
constexpr int kMaxStages = 16;

template <int BN>
constexpr uint32_t acc_stride() { return (BN + 31) / 32 * 32; }   // TMEM columns per accumulator buffer
// accumulator ring in TMEM: as many BN-column buffers as fit in the 512 columns, at most 4
template <int BN>
constexpr uint32_t num_acc() { return 512 / acc_stride<BN>() >= 4 ? 4 : 512 / acc_stride<BN>(); }
template <int BN>
constexpr uint32_t tmem_cols() {
  return num_acc<BN>() * acc_stride<BN>() <= 32 ? 32 : num_acc<BN>() * acc_stride<BN>() <= 64 ? 64
         : num_acc<BN>() * acc_stride<BN>() <= 128 ? 128 : num_acc<BN>() * acc_stride<BN>() <= 256 ? 256 : 512;
}

// control block placed after the operand ring
template <int BN>
struct alignas(16) TcControl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[4];
  uint64_t tmem_empty[4];
  uint32_t tmem_slot;
  uint32_t pad_[3];
  int32_t oc[2][BN];
  float bias[2][BN];
  float sbv[2][BN];   // per-output-channel weight scales of the tile (F4 extension; unused otherwise)
};

template <int BN, int BK>
__host__ __device__ constexpr int stage_bytes(int ksub, int mt) { return ksub * (mt * BM * BK + BN * BK); }

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); }


// One output row (= one TMEM lane) of a 128 x BN accumulator tile: + oc, [border correction],
// [fc float bias], requantise, [relu], packed u8 store. m < 0: row is padding (nothing stored).
// NSUB warps share each TMEM lane quadrant; warp `half` of them takes the chunks half, half + NSUB, ...
// `s_oc` / `s_bias` are shared-memory addresses of this tile's staged per-channel terms.
template <int BN, int NSUB = kEpiWarps / 4>   // NSUB warps share a quadrant and take every NSUB-th chunk
__device__ __forceinline__ void epilogue_row(const TcParams& p, uint32_t t_row, long long m, int n0,
                                             uint32_t s_oc, uint32_t s_bias, const int32_t* corr, float rcp,
                                             int half, uint32_t s_sb = 0) {
  const float zpf = (float)p.ep.zp_out;
  const float sa = p.ep.sa, sb = p.ep.sb, sc = p.ep.sc;
  const bool has_bias = p.ep.bias_f != nullptr;
  const RequantFast2 rq = make_requant_fast2(sa, sb, sc, rcp, zpf);
  uint8_t* yrow = p.y + (size_t)(m < 0 ? 0 : m) * p.out_cp;
#pragma unroll 1
  for (int c0 = half * 32; c0 < BN; c0 += 32 * NSUB) {
    if (n0 + c0 >= p.out_cp) break;   // warp-uniform
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)c0, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const uint4 o = ptx::lds128(s_oc + (uint32_t)(c0 + 4 * g) * 4);
      v[4 * g] += o.x; v[4 * g + 1] += o.y; v[4 * g + 2] += o.z; v[4 * g + 3] += o.w;
    }
    if (corr) {   // border pixels only: zero-fill -> zero-point padding correction
      // (table rows are padded to a multiple of 32 channels, so whole chunks are read as int4)
      const int4* c4 = reinterpret_cast<const int4*>(corr + n0 + c0);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int4 t = __ldg(c4 + g);
        v[4 * g] += (uint32_t)(p.zp_in * t.x); v[4 * g + 1] += (uint32_t)(p.zp_in * t.y);
        v[4 * g + 2] += (uint32_t)(p.zp_in * t.z); v[4 * g + 3] += (uint32_t)(p.zp_in * t.w);
      }
    }
    if (has_bias) {   // fully_connected.cc:44
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint4 b = ptx::lds128(s_bias + (uint32_t)(c0 + 4 * g) * 4);
        v[4 * g] = (uint32_t)fc_bias_add((int32_t)v[4 * g], __uint_as_float(b.x));
        v[4 * g + 1] = (uint32_t)fc_bias_add((int32_t)v[4 * g + 1], __uint_as_float(b.y));
        v[4 * g + 2] = (uint32_t)fc_bias_add((int32_t)v[4 * g + 2], __uint_as_float(b.z));
        v[4 * g + 3] = (uint32_t)fc_bias_add((int32_t)v[4 * g + 3], __uint_as_float(b.w));
      }
    }
    if (p.ep.acc_out && m >= 0) {   // parity-test dump
      for (int j = 0; j < 32; ++j)
        if (n0 + c0 + j < p.N) p.ep.acc_out[(size_t)m * p.N + n0 + c0 + j] = (int32_t)v[j];
    }
    if (p.fast_requant) {
      // exact requantise, two accumulators per packed fp32x2 instruction (see requant2_u8_fast);
      // clamp + truncation are one saturating convert. relu<u8> = max(y, zp) (functional.cc:22-23)
      // is applied to the ACCUMULATOR instead: with positive scales (requant_fast_ok) the requantised
      // value is >= zp exactly when acc >= 0, and acc = 0 maps to zp itself, so max(acc, 0) gives the
      // same byte — and fuses with the offset add above into one add-max instruction.
      if (p.ep.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (uint32_t)max((int32_t)v[j], 0);
      }
      if (s_sb == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) requant2_u8_fast<false>((int32_t)v[j], (int32_t)v[j + 1], rq, v[j], v[j + 1]);
      } else {   // per-output-channel weight scales (staged next to oc)
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint4 s4 = ptx::lds128(s_sb + (uint32_t)(c0 + 4 * g) * 4);
          requant2_u8_fast_sb<false>((int32_t)v[4 * g], (int32_t)v[4 * g + 1], rq,
                                     f2_pack(__uint_as_float(s4.x), __uint_as_float(s4.y)), v[4 * g], v[4 * g + 1]);
          requant2_u8_fast_sb<false>((int32_t)v[4 * g + 2], (int32_t)v[4 * g + 3], rq,
                                     f2_pack(__uint_as_float(s4.z), __uint_as_float(s4.w)), v[4 * g + 2], v[4 * g + 3]);
        }
      }
    } else {
      const uint32_t zlo = p.ep.relu ? (uint32_t)p.ep.zp_out : 0u;
      if (s_sb == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = max(requant_u8((int32_t)v[j], sa, sb, sc, zpf), zlo);
      } else {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint4 s4 = ptx::lds128(s_sb + (uint32_t)(c0 + 4 * g) * 4);
          v[4 * g] = max(requant_u8((int32_t)v[4 * g], sa, __uint_as_float(s4.x), sc, zpf), zlo);
          v[4 * g + 1] = max(requant_u8((int32_t)v[4 * g + 1], sa, __uint_as_float(s4.y), sc, zpf), zlo);
          v[4 * g + 2] = max(requant_u8((int32_t)v[4 * g + 2], sa, __uint_as_float(s4.z), sc, zpf), zlo);
          v[4 * g + 3] = max(requant_u8((int32_t)v[4 * g + 3], sa, __uint_as_float(s4.w), sc, zpf), zlo);
        }
      }
    }
    if (n0 + c0 + 32 > p.N) {   // pad lanes carry the zero point (warp-uniform branch)
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + c0 + j >= p.N) v[j] = (uint32_t)p.ep.zp_out;
    }
    uint32_t pk[8];
#pragma unroll
    for (int g = 0; g < 8; ++g)   // low bytes of four registers -> one word
      pk[g] = __byte_perm(__byte_perm(v[4 * g], v[4 * g + 1], 0x0040), __byte_perm(v[4 * g + 2], v[4 * g + 3], 0x0040),
                          0x5410);
    if (m >= 0) {
      *reinterpret_cast<uint4*>(yrow + n0 + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      if (n0 + c0 + 16 < p.out_cp)
        *reinterpret_cast<uint4*>(yrow + n0 + c0 + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
  }
}

// Persistent, warp-specialised implicit GEMM. MODE 0: A rows via a 2-D tiled map (fc);
// MODE 1: A gathered by the TMA im2col engine (conv, incl. the stem view).
template <int BN, int BK, int MODE, int MT>
__global__ void __launch_bounds__(kThreads, 1) tc_igemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kSubA = BM * BK, kSubB = BN * BK;
  constexpr uint32_t NACC = num_acc<BN>();
  constexpr int mt = MT;
  const int kStage = p.ksub * (mt * kSubA + kSubB);   // [ksub][mt] A sub-blocks, then [ksub] B sub-blocks
  TcControl<BN>* ctl = reinterpret_cast<TcControl<BN>*>(smem + (size_t)p.stages * kStage);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mn_tiles = p.tiles_m * p.tiles_n;
  const int num_tiles = mn_tiles * p.splits;
  const int tile0 = blockIdx.x, tile_step = gridDim.x;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < (int)NACC; ++b) {
      ptx::mbar_init(&ctl->tmem_full[b], 1);
      ptx::mbar_init(&ctl->tmem_empty[b], kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(&ctl->tmem_slot, tmem_cols<BN>());
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory from here on

  if (warp == 0) {
    // ===== TMA producer: runs ahead across tile boundaries, the ring never drains. The whole warp walks the
    // loop (warp-uniform operands), one elected lane issues; ring cursors advance incrementally =====
    const uint32_t smem_a = ptx::smem_u32(smem);
    const uint32_t full_a = ptx::smem_u32(&ctl->full[0]), empty_a = ptx::smem_u32(&ctl->empty[0]);
    const uint32_t nstages = (uint32_t)p.stages;
    uint32_t s = 0, ph = 0;
    bool alive = true;
    for (int tile = tile0; tile < num_tiles && alive; tile += tile_step) {
      const int split = tile / mn_tiles, mn = tile % mn_tiles;
      const int m0 = (mn / p.tiles_n) * BM * mt, n0 = (mn % p.tiles_n) * BN;
      const int kb0 = split * p.kb_per, kb1 = min(p.num_kb, kb0 + p.kb_per);
      int bw[MT] = {}, bh[MT] = {}, bn[MT] = {};
      if (MODE == 1) {
#pragma unroll
        for (int t = 0; t < mt; ++t) {
          const int mm = m0 + t * BM;
          const int q = mm % p.ow, r = mm / p.ow;
          bw[t] = q * p.stride_w - p.pad;
          bh[t] = (r % p.oh) * p.stride_h - p.pad;
          bn[t] = r / p.oh;
        }
      }
      int cb = 0, kx = 0, ky = 0;   // (split-K is only used with MODE 0, where these stay 0)
      int kcol = kb0 * p.ksub * BK;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!ptx::mbar_wait_a(empty_a + 8u * s, ph ^ 1u)) { tc_fail(1); alive = false; break; }
        const bool el = ptx::elect_one_sync();
        const uint32_t fb = full_a + 8u * s;
        if (el) ptx::mbar_arrive_expect_tx_a(fb, (uint32_t)kStage);
        const uint32_t sa = smem_a + s * (uint32_t)kStage;
        const uint32_t sb = sa + (uint32_t)(p.ksub * mt * kSubA);
        for (int j = 0; j < p.ksub; ++j) {
          if (el) {
#pragma unroll
            for (int t = 0; t < mt; ++t) {
              const uint32_t dst = sa + (uint32_t)((j * mt + t) * kSubA);
              if (MODE == 1)
                ptx::tma_load_im2col_4d_a(dst, &tmA, fb, cb, bw[t], bh[t], bn[t], (uint16_t)kx, (uint16_t)ky);
              else
                ptx::tma_load_2d_a(dst, &tmA, fb, kcol, m0 + t * BM);
            }
            if (MODE == 0 && BK == 128 && BN % 128 == 0 && p.w_tiled != nullptr) {
              const int8_t* blk = p.w_tiled + ((size_t)(n0 >> 7) * p.wt_nkb + (size_t)(kcol >> 7)) * (128 * 128);
#pragma unroll
              for (int h = 0; h < BN / 128; ++h)
                ptx::bulk_load_1d_a(sb + (uint32_t)(j * kSubB + h * 128 * 128), blk + (size_t)h * p.wt_nkb * (128 * 128),
                                    128 * 128, fb);
            } else {
              ptx::tma_load_2d_a(sb + (uint32_t)(j * kSubB), &tmB, fb, kcol, n0);
            }
          }
          if (MODE == 1) { cb += BK; if (cb == p.cblocks * BK) { cb = 0; if (++kx == p.kw) { kx = 0; ++ky; } } }
          kcol += BK;
        }
        __syncwarp();
        if (++s == nstages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: whole warp walks the loop, one elected lane issues; accumulators rotate through TMEM =====
    constexpr uint32_t idesc = ptx::make_idesc_i8(BM, BN);
    const uint32_t desc_hi = ptx::smem_desc_hi<BK>();
    const uint32_t full_a = ptx::smem_u32(&ctl->full[0]), empty_a = ptx::smem_u32(&ctl->empty[0]);
    const uint32_t tfull_a = ptx::smem_u32(&ctl->tmem_full[0]), tempty_a = ptx::smem_u32(&ctl->tmem_empty[0]);
    const uint32_t lo0 = ptx::smem_desc_lo(ptx::smem_u32(smem));
    const uint32_t nstages = (uint32_t)p.stages;
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t s = 0, ph = 0, slot = 0, sph = 0;
    bool alive = true;
    for (int tile = tile0; tile < num_tiles && alive; tile += tile_step) {
      uint32_t d_tmem[MT], slots[MT];
#pragma unroll
      for (int t = 0; t < mt; ++t) {
        if (alive && !ptx::mbar_wait_a(tempty_a + 8u * slot, sph ^ 1u)) { tc_fail(4); alive = false; }
        d_tmem[t] = tbase + slot * acc_stride<BN>();
        slots[t] = slot;
        if (++slot == NACC) { slot = 0; sph ^= 1u; }
      }
      if (!alive) break;
      ptx::tc_fence_after();
      const int kb0 = (tile / mn_tiles) * p.kb_per, kb1 = min(p.num_kb, kb0 + p.kb_per);
      uint32_t accf = 0;   // the first MMA of a tile overwrites the accumulator
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!ptx::mbar_wait_a(full_a + 8u * s, ph)) { tc_fail(2); alive = false; break; }
        ptx::tc_fence_after();
        if (ptx::elect_one_sync()) {
          // descriptor low words; every further operand is a constant (>>4) offset away
          uint32_t a_lo = lo0 + s * (uint32_t)(kStage >> 4);
          uint32_t b_lo = a_lo + (uint32_t)((p.ksub * mt * kSubA) >> 4);
          for (int j = 0; j < p.ksub; ++j) {
#pragma unroll
            for (int k = 0; k < BK / 32; ++k) {
#pragma unroll
              for (int t = 0; t < mt; ++t)
                ptx::mma_i8_ss_lohi(d_tmem[t], a_lo + (uint32_t)((t * kSubA + k * 32) >> 4), desc_hi,
                                    b_lo + (uint32_t)((k * 32) >> 4), desc_hi, idesc, accf);
              accf = 1;
            }
            a_lo += (uint32_t)((mt * kSubA) >> 4);
            b_lo += (uint32_t)(kSubB >> 4);
          }
          ptx::tc_commit_a(empty_a + 8u * s);   // slot reusable once these MMAs have read it
        }
        __syncwarp();
        if (++s == nstages) { s = 0; ph ^= 1u; }
      }
      if (alive && ptx::elect_one_sync()) {
#pragma unroll
        for (int t = 0; t < mt; ++t) ptx::tc_commit_a(tfull_a + 8u * slots[t]);
      }
      __syncwarp();
    }
    if (!alive && lane == 0)   // unblock the epilogue so that the CTA can exit
      for (int b = 0; b < (int)NACC; ++b) ptx::mbar_arrive(&ctl->tmem_full[b]);
  } else {
    // ===== epilogue: kEpiWarps warps, warp w owns TMEM lanes [32*(w%4), +32) = output rows =====
    const int quad = warp & 3;
    const int et = threadIdx.x - 64;
    const float rcp = __frcp_rn(p.ep.sc);
    const bool has_bias = p.ep.bias_f != nullptr;
    uint32_t tcount = 0, acc_it = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++tcount, acc_it += mt) {
      const uint32_t ob = tcount & 1;
      const int split = tile / mn_tiles, mn = tile % mn_tiles;
      const int m0 = (mn / p.tiles_n) * BM * mt, n0 = (mn % p.tiles_n) * BN;
      if (p.splits <= 1) {
        // stage this tile's per-channel offsets (double-buffered: one barrier per tile)
        for (int j = et; j < BN; j += 32 * kEpiWarps) {
          const int n = n0 + j;
          ctl->oc[ob][j] = (n < p.N) ? __ldg(p.ep.oc + n) : 0;
          ctl->bias[ob][j] = (n < p.N && has_bias) ? __ldg(p.ep.bias_f + n) : 0.f;
          if (p.ep.sb_vec) ctl->sbv[ob][j] = (n < p.N) ? __ldg(p.ep.sb_vec + n) : 1.f;
        }
        epi_bar_sync();
      }
#pragma unroll 1
      for (int t = 0; t < mt; ++t) {
        const uint32_t slot = (acc_it + t) % NACC, sph = ((acc_it + t) / NACC) & 1;
        const int m = m0 + t * BM + quad * 32 + lane;
        const uint32_t t_row = tmem_base + slot * acc_stride<BN>() + ((uint32_t)(quad * 32) << 16);
        if (p.splits > 1) {
          // split-K: dump the raw partial accumulators; fc_splitk_reduce_kernel finishes the job
          const bool ok = ptx::mbar_wait(&ctl->tmem_full[slot], sph);
          if (!ok) tc_fail(3);
          ptx::tc_fence_after();
          int32_t* wrow = p.ws + ((size_t)split * p.M + (m < p.M ? m : 0)) * p.ws_ld + n0;
#pragma unroll 1
          for (int c0 = ((warp - 2) >> 2) * 32; c0 < BN; c0 += 32 * (kEpiWarps / 4)) {
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)c0, v);
            ptx::tmem_ld_wait();
            if (m < p.M && ok) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                *reinterpret_cast<uint4*>(wrow + c0 + 4 * g) = make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
            }
          }
        } else {
          // spatial border class of this output pixel -> row of the zero-point correction table
          const int32_t* corr = nullptr;
          if (MODE == 1 && p.border_tab && m < p.M) {
            const int q = m % p.ow, pr = (m / p.ow) % p.oh;
            const int y0 = pr * p.stride_h - p.pad, x0 = q * p.stride_w - p.pad;
            const int th = min(max(-y0, 0), p.kh), bh = min(max(y0 + p.kh - p.H, 0), p.kh);
            const int tw = min(max(-x0, 0), p.kw), bw = min(max(x0 + p.kw - p.W, 0), p.kw);
            const int d = p.pad + 1;
            const int cls = ((th * d + bh) * d + tw) * d + bw;
            if (cls != 0) corr = p.border_tab + (size_t)cls * ((p.N + 31) & ~31);
          }
          const bool ok = ptx::mbar_wait(&ctl->tmem_full[slot], sph);
          if (!ok) tc_fail(3);
          ptx::tc_fence_after();
          epilogue_row<BN>(p, t_row, (m < p.M && ok) ? (long long)m : -1ll, n0, ptx::smem_u32(ctl->oc[ob]),
                           ptx::smem_u32(ctl->bias[ob]), corr, rcp, (warp - 2) >> 2,
                           p.ep.sb_vec ? ptx::smem_u32(ctl->sbv[ob]) : 0u);
        }
        // hand the accumulator buffer back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&ctl->tmem_empty[slot]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, tmem_cols<BN>());
}

// ---- small-M fc: split-K over a thread-block CLUSTER, partial sums folded through distributed shared memory ----
// A weight stream with one M tile has only N / 128 output tiles, so K is split over the S = 4 CTAs of ONE cluster
// (fc1 at batch <= 128: 32 clusters of 4 = 128 SMs pull the 37.7 MB of weights). Each CTA accumulates its K slice of
// the whole 128 x 128 tile in TMEM; then, instead of a round trip of s32 partials through global memory and a second
// kernel (fc_splitk_reduce_kernel: 14.7 MB written + read and 4.5 us for fc1), rank c of the cluster OWNS the 32
// columns [32c, 32c + 32): every epilogue warp stages its 32 x 32 accumulator chunk in local shared memory and one
// bulk copy (cp.async.bulk.shared::cluster.shared::cta, 4 KB) moves it into the owner's receive buffer, completing
// on the owner's mbarrier; the owner sums the S partials, requantises and stores its 128 x 32 slice. Integer adds
// commute, so the result is bit-identical to the unsplit sum. (Per-lane st.shared::cluster stores of the same
// chunks — 16-byte pieces 128 bytes apart — took 16 us for fc1: the SM-to-SM network wants bulk transfers.)
//   roles: warp 0 TMA / bulk producer, warp 1 MMA issuer, warps 2-17 chunk push + fold
//   receive buffer [S sources][128 rows][128 bytes] and the 16 x 4 KB staging area re-use the operand ring (free
//   once every CTA of the cluster has seen its last MMA complete: cluster barrier); 16-byte units of a row are
//   XOR-swizzled by (row & 7) so that the staging stores (one row per lane) and the fold (4 threads per row) are
//   bank-conflict free.
template <int BN>
__global__ void __launch_bounds__(kThreads, 1) tc_fc_cluster_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmB,
                                                                    const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int BK = 128, S = BN / 32;
  constexpr int kSubA = BM * BK, kSubB = BN * BK, kStage = kSubA + kSubB;
  static_assert(BN == 128 && kEpiWarps == 16, "cluster fc: one 32-column chunk per epilogue warp, one owner rank per chunk");
  constexpr uint32_t kRecvBytes = S * 128 * 128;
  TcControl<BN>* ctl = reinterpret_cast<TcControl<BN>*>(smem + (size_t)p.stages * kStage);
  uint64_t* recv_full = &ctl->tmem_full[1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const int tile = (int)blockIdx.x / S;                      // one (M tile, N tile) per cluster
  const int m0 = (tile / p.tiles_n) * BM, n0 = (tile % p.tiles_n) * BN;
  const int kb0 = (int)crank * p.kb_per, kb1 = min(p.num_kb, kb0 + p.kb_per);   // this CTA's K blocks (>= 1)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    ptx::mbar_init(&ctl->tmem_full[0], 1);
    ptx::mbar_init(recv_full, 1);
    ptx::fence_barrier_init();
    ptx::mbar_arrive_expect_tx(recv_full, (p.dbg & 1) ? 0u : kRecvBytes);   // S sources x 16 KB will land here
  }
  if (warp == 1) ptx::tmem_alloc(&ctl->tmem_slot, tmem_cols<BN>());
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_slot;
  // The weights do not depend on the previous kernel: with programmatic dependent launch this CTA may be resident
  // while its predecessor still runs, so the producer asks for the weight halves of the first ring pass BEFORE
  // griddepcontrol.wait and adds the activation tiles (the predecessor's output) after it.
  const int npre = min(p.stages, kb1 - kb0);
  auto load_w = [&](uint32_t sb, uint32_t fb, int kb, int kcol) {
    if (p.w_tiled != nullptr)   // one contiguous pre-swizzled 16 KB block (see TcParams::w_tiled)
      ptx::bulk_load_1d_a(sb, p.w_tiled + ((size_t)(n0 >> 7) * p.wt_nkb + (size_t)kb) * (128 * 128), 128 * 128, fb);
    else
      ptx::tma_load_2d_a(sb, &tmB, fb, kcol, n0);
  };
  if (warp == 0 && !(p.dbg & 4)) {
    if (ptx::elect_one_sync()) {
      for (int i = 0; i < npre; ++i) {
        const uint32_t fb = ptx::smem_u32(&ctl->full[i]);
        ptx::mbar_arrive_expect_tx_a(fb, (uint32_t)kStage);
        load_w(ptx::smem_u32(smem) + (uint32_t)i * (uint32_t)kStage + (uint32_t)kSubA, fb, kb0 + i, (kb0 + i) * BK);
      }
    }
    __syncwarp();
  }
  pdl_wait();

  bool ok = true;
  if (warp == 0) {
    const uint32_t smem_a = ptx::smem_u32(smem);
    const uint32_t full_a = ptx::smem_u32(&ctl->full[0]), empty_a = ptx::smem_u32(&ctl->empty[0]);
    const uint32_t nstages = (uint32_t)p.stages;
    uint32_t s = 0, ph = 0;
    int kcol = kb0 * BK;
    for (int kb = kb0; kb < kb1; ++kb) {
      const bool first_pass = kb - kb0 < npre;   // stage untouched so far, its weight half is already on the way
      if (!first_pass && !ptx::mbar_wait_a(empty_a + 8u * s, ph ^ 1u)) { tc_fail(1); break; }
      if (ptx::elect_one_sync()) {
        const uint32_t fb = full_a + 8u * s;
        if (p.dbg & 4) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb) : "memory");
        } else {
          const uint32_t sa = smem_a + s * (uint32_t)kStage, sb = sa + (uint32_t)kSubA;
          if (!first_pass) {
            ptx::mbar_arrive_expect_tx_a(fb, (uint32_t)kStage);
            load_w(sb, fb, kb, kcol);
          }
          ptx::tma_load_2d_a(sa, &tmA, fb, kcol, m0);
        }
      }
      __syncwarp();
      kcol += BK;
      if (++s == nstages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = ptx::make_idesc_i8(BM, BN);
    const uint32_t desc_hi = ptx::smem_desc_hi<BK>();
    const uint32_t full_a = ptx::smem_u32(&ctl->full[0]), empty_a = ptx::smem_u32(&ctl->empty[0]);
    const uint32_t lo0 = ptx::smem_desc_lo(ptx::smem_u32(smem));
    const uint32_t nstages = (uint32_t)p.stages;
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t s = 0, ph = 0, accf = 0;
    bool alive = true;
    for (int kb = kb0; kb < kb1; ++kb) {
      if (!ptx::mbar_wait_a(full_a + 8u * s, ph)) { tc_fail(2); alive = false; break; }
      ptx::tc_fence_after();
      if (ptx::elect_one_sync()) {
        const uint32_t a_lo = lo0 + s * (uint32_t)(kStage >> 4), b_lo = a_lo + (uint32_t)(kSubA >> 4);
#pragma unroll
        for (int k = 0; k < BK / 32; ++k) {
          ptx::mma_i8_ss_lohi(tbase, a_lo + (uint32_t)((k * 32) >> 4), desc_hi, b_lo + (uint32_t)((k * 32) >> 4), desc_hi,
                              idesc, accf);
          accf = 1;
        }
        ptx::tc_commit_a(empty_a + 8u * s);
      }
      __syncwarp();
      if (++s == nstages) { s = 0; ph ^= 1u; }
    }
    if (alive && ptx::elect_one_sync()) ptx::tc_commit(&ctl->tmem_full[0]);
    __syncwarp();
    if (!alive && lane == 0) ptx::mbar_arrive(&ctl->tmem_full[0]);
  }

  // ---- chunk push: warp (quadrant q, owner c) holds columns [32c, 32c + 32) of TMEM lanes [32q, 32q + 32) ----
  uint32_t v[32];
  const int quad = warp & 3, owner = (warp - 2) >> 2;
  // fold mapping: 4 threads per row, 8 columns each; their per-channel terms are fetched now, long before the fold
  const int et = threadIdx.x - 64;
  const int frow = et >> 2, cg = et & 3;
  const int nb = n0 + (int)crank * 32 + cg * 8;
  int32_t oc8[8];
  float bias8[8], sb8[8];
  if (warp >= 2) {
    const EpiParams& ep = p.ep;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool real = nb + j < p.N;
      oc8[j] = real ? __ldg(ep.oc + nb + j) : 0;
      bias8[j] = (real && ep.bias_f) ? __ldg(ep.bias_f + nb + j) : 0.f;
      sb8[j] = (real && ep.sb_vec) ? __ldg(ep.sb_vec + nb + j) : ep.sb;
    }
    ok = ptx::mbar_wait(&ctl->tmem_full[0], 0);
    if (!ok) tc_fail(3);
    ptx::tc_fence_after();
    ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(owner * 32), v);
    ptx::tmem_ld_wait();
    ptx::tc_fence_before();
  }
  __syncwarp();
  ptx::cluster_sync_all();   // every CTA's last MMA has completed: every operand ring may now be overwritten
  if (warp >= 2 && !(p.dbg & 1)) {
    const uint32_t stage = ptx::smem_u32(smem) + kRecvBytes + (uint32_t)(warp - 2) * 4096u;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      ptx::sts128(stage + (uint32_t)lane * 128u + (((uint32_t)u ^ ((uint32_t)lane & 7u)) << 4),
                  make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> the bulk copy's reads
    __syncwarp();
    if (lane == 0) {
      const uint32_t dst = ptx::mapa(ptx::smem_u32(smem) + (crank * 128u + (uint32_t)(quad * 32)) * 128u, (uint32_t)owner);
      ptx::bulk_copy_to_cluster(dst, stage, 4096u, ptx::mapa(ptx::smem_u32(recv_full), (uint32_t)owner));
    }
  }
  if (warp >= 2 && ok && !ptx::mbar_wait(recv_full, 0)) { tc_fail(8); ok = false; }   // the S x 16 KB of partial chunks have landed
  if (warp >= 2 && !(p.dbg & 2)) {
    // ---- fold + fused epilogue of this rank's 128 x 32 slice ----
    const int m = m0 + frow;
    const uint32_t base = ptx::smem_u32(smem) + (uint32_t)frow * 128u;
    const uint32_t o0 = ((uint32_t)(2 * cg) ^ ((uint32_t)frow & 7u)) << 4, o1 = ((uint32_t)(2 * cg + 1) ^ ((uint32_t)frow & 7u)) << 4;
    int32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const uint4 a = ptx::lds128(base + (uint32_t)s * (128u * 128u) + o0);
      const uint4 b = ptx::lds128(base + (uint32_t)s * (128u * 128u) + o1);
      acc[0] += (int32_t)a.x; acc[1] += (int32_t)a.y; acc[2] += (int32_t)a.z; acc[3] += (int32_t)a.w;
      acc[4] += (int32_t)b.x; acc[5] += (int32_t)b.y; acc[6] += (int32_t)b.z; acc[7] += (int32_t)b.w;
    }
    if (m < p.M && ok) {
      const EpiParams& ep = p.ep;
      const float zpf = (float)ep.zp_out, rcp = __frcp_rn(ep.sc);
      const uint32_t zlo = ep.relu ? (uint32_t)ep.zp_out : 0u;
      uint32_t word[2] = {0u, 0u};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = nb + j;
        uint32_t q = (uint32_t)ep.zp_out;   // pad lanes carry the zero point
        if (n < p.N) {
          int32_t a = acc[j] + oc8[j];
          if (ep.bias_f) a = fc_bias_add(a, bias8[j]);
          if (ep.acc_out) ep.acc_out[(size_t)m * p.N + n] = a;
          const uint32_t r = p.fast_requant ? requant_u8_fast(a, ep.sa, sb8[j], ep.sc, rcp, zpf) : requant_u8(a, ep.sa, sb8[j], ep.sc, zpf);
          q = max(r, zlo);
        }
        word[j >> 2] |= q << (8 * (j & 3));
      }
      if (nb < p.out_cp) *reinterpret_cast<uint2*>(p.y + (size_t)m * p.out_cp + nb) = make_uint2(word[0], word[1]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, tmem_cols<BN>());
  __syncwarp();
  ptx::cluster_sync_all();   // no CTA leaves while a peer's bulk copy may still be reading its staging area
}

// ---- CTA-pair kernel (cta_group::2): conv with wide N tiles ------------------------------------
// A 128 x BN tile per SM needs (128 + BN) * BK bytes per K block through the SM's L2 port, which is
// what bounds tc_igemm_kernel (ncu: ~46 B/clk/SM, tensor pipe ~50 % busy). Here two SMs of a
// cluster work on ONE 256 x BN tile: each CTA loads its own 128 rows of A and only HALF of the
// weight rows; tcgen05.mma.cta_group::2 (issued by the even CTA) reads both halves and fills both
// CTAs' TMEM (rows 0-127 / 128-255). Per-SM operand traffic drops to (128 + BN/2) * BK.
// BN = 384 (layers with 384 output channels): ONE A stage feeds two MMAs of N = 192 (columns 0-191 and
// 192-383 of a single 384-column accumulator), so A is fetched once instead of once per 192-wide N tile:
// 40 KB per 768 tensor clocks (52 B/clk/SM) instead of 28 KB per 384 (73 B/clk/SM) against the measured
// 46.6 B/clk/SM L2 -> SM ingest (profiles/i8_peak.json). TMEM then holds one accumulator, so the epilogue
// of a tile no longer overlaps the next tile's MMAs — the host picks BN per plan from a cost model.
// MT = 2 (BN = 256 only): a pair works on a 512 x 256 tile — two 256-row accumulators that share every weight
// stage (48 KB per 1024 tensor clocks = 47 B/clk/SM instead of 32 KB per 512 = 64 B/clk/SM). Both accumulators
// fill TMEM, so again there is no epilogue overlap; worth it when the layer has several waves of tiles (conv2).
//   full[s]       even CTA only, 1 arrival (its producer, expecting the bytes of BOTH CTAs)
//   empty[s]      each CTA, 1 arrival (the even CTA's multicast commit)
//   tmem_full[b]  each CTA, 1 arrival (multicast commit)
//   tmem_empty[b] even CTA only, 2 * kEpiWarps2 arrivals (the epilogue warps of both CTAs)
template <int BN, int MT>
__global__ void __launch_bounds__(kThreads2, 1) tc_igemm2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int BK = 128;
  constexpr int NSPLIT = BN > 256 ? 2 : 1;     // MMAs per K step (UMMA N <= 256)
  constexpr int BNI = BN / NSPLIT;             // N of one MMA instruction
  constexpr int kSubA = BM * BK, kSubBq = (BNI / 2) * BK, kSubBh = NSPLIT * kSubBq;
  constexpr int kStage = MT * kSubA + kSubBh;   // [MT A sub-tiles][this CTA's weight rows]
  static_assert(num_acc<BN>() >= (uint32_t)MT, "accumulators of one tile must fit TMEM");
  constexpr uint32_t NACC = num_acc<BN>() / MT;   // tile slots in TMEM, each MT accumulators of BN columns
  constexpr uint32_t kSlotCols = MT * acc_stride<BN>();
  TcControl<BN>* ctl = reinterpret_cast<TcControl<BN>*>(smem + (size_t)p.stages * kStage);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)(blockIdx.x & 1);     // cluster dims (2, 1, 1)
  const bool leader = crank == 0;
  const int mn_tiles = p.tiles_mp * p.tiles_n;
  const int tile0 = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < (int)NACC; ++b) {
      ptx::mbar_init(&ctl->tmem_full[b], 1);
      ptx::mbar_init(&ctl->tmem_empty[b], 2 * kEpiWarps2);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_2cta(&ctl->tmem_slot, tmem_cols<BN>());
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_slot;
  ptx::cluster_sync_all();   // the peer's barriers and TMEM exist before anything targets them
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 rows of A, own half of the weight rows =====
    // The WHOLE warp walks the loop (every value below is warp-uniform), one elected lane issues. The ring
    // cursor advances incrementally: a loop iteration is a few dozen instructions, so the single-thread issue
    // path is never what bounds a K block (with per-iteration div / mod and per-lane operands it was: ~850 clk).
    const uint32_t smem_a = ptx::smem_u32(smem);
    const uint32_t full_a = ptx::smem_u32(&ctl->full[0]), empty_a = ptx::smem_u32(&ctl->empty[0]);
    const uint32_t nstages = (uint32_t)p.stages;
    uint32_t s = 0, ph = 0;
    bool alive = true;
    for (int tile = tile0; tile < mn_tiles && alive; tile += tile_step) {
      const int n0 = (tile % p.tiles_n) * BN;
      int bw[MT], bh[MT], bimg[MT];
#pragma unroll
      for (int a = 0; a < MT; ++a) {   // accumulator a of this CTA = M tile (2 * MT) * T + 2 * a + crank
        // tail: M tiles past the end re-load the last one (their stores are masked)
        const int m0 = min((tile / p.tiles_n) * (2 * MT) + 2 * a + crank, p.tiles_m - 1) * BM;
        const int q = m0 % p.ow, r = m0 / p.ow;
        bw[a] = q * p.stride_w - p.pad; bh[a] = (r % p.oh) * p.stride_h - p.pad; bimg[a] = r / p.oh;
      }
      const int wrow = n0 + crank * (BNI / 2);
      int cb = 0, kx = 0, ky = 0, kcol = 0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        if (!ptx::mbar_wait_a(empty_a + 8u * s, ph ^ 1u)) { tc_fail(1); alive = false; break; }
        if (ptx::elect_one_sync()) {
          const uint32_t fb = full_a + 8u * s;
          if (leader) ptx::mbar_arrive_expect_tx_a(fb, 2u * (uint32_t)kStage);
          const uint32_t sa = smem_a + s * (uint32_t)kStage;
#pragma unroll
          for (int a = 0; a < MT; ++a)
            ptx::tma_load_im2col_4d_2cta_a(sa + (uint32_t)(a * kSubA), &tmA, fb, cb, bw[a], bh[a], bimg[a], (uint16_t)kx,
                                           (uint16_t)ky);
#pragma unroll
          for (int h = 0; h < NSPLIT; ++h)   // this CTA's half of the weight rows of every N = BNI instruction
            ptx::tma_load_2d_2cta_a(sa + (uint32_t)(MT * kSubA + h * kSubBq), &tmB, fb, kcol, wrow + h * BNI);
        }
        __syncwarp();
        kcol += BK;
        cb += BK;
        if (cb == p.cblocks * BK) { cb = 0; if (++kx == p.kw) { kx = 0; ++ky; } }
        if (++s == nstages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the even CTA's warp 1 drives the tensor cores of both SMs (whole warp walks the
    // loop, one elected lane issues; descriptors advance by constants) =====
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_i8(2 * BM, BNI);
      const uint32_t desc_hi = ptx::smem_desc_hi<BK>();
      const uint32_t full_a = ptx::smem_u32(&ctl->full[0]), empty_a = ptx::smem_u32(&ctl->empty[0]);
      const uint32_t tfull_a = ptx::smem_u32(&ctl->tmem_full[0]), tempty_a = ptx::smem_u32(&ctl->tmem_empty[0]);
      const uint32_t lo0 = ptx::smem_desc_lo(ptx::smem_u32(smem));
      const uint32_t nstages = (uint32_t)p.stages;
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t s = 0, ph = 0, slot = 0, sph = 0;
      bool alive = true;
      for (int tile = tile0; tile < mn_tiles && alive; tile += tile_step) {
        if (!ptx::mbar_wait_a(tempty_a + 8u * slot, sph ^ 1u)) { tc_fail(4); alive = false; break; }
        ptx::tc_fence_after();
        const uint32_t d_tmem = tbase + slot * kSlotCols;
        uint32_t accf = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          if (!ptx::mbar_wait_a(full_a + 8u * s, ph)) { tc_fail(2); alive = false; break; }
          ptx::tc_fence_after();
          if (ptx::elect_one_sync()) {
            const uint32_t a_lo = lo0 + s * (uint32_t)(kStage >> 4), b_lo = a_lo + (uint32_t)((MT * kSubA) >> 4);
#pragma unroll
            for (int k = 0; k < BK / 32; ++k) {
#pragma unroll
              for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int h = 0; h < NSPLIT; ++h)
                  ptx::mma_i8_ss_lohi_2cta(d_tmem + (uint32_t)(a * acc_stride<BN>() + h * BNI),
                                           a_lo + (uint32_t)((a * kSubA + k * 32) >> 4), desc_hi,
                                           b_lo + (uint32_t)((h * kSubBq + k * 32) >> 4), desc_hi, idesc, accf);
              accf = 1;
            }
            ptx::tc_commit_2cta_multicast_a(empty_a + 8u * s, 3);   // both CTAs' slots are free once these MMAs have read them
          }
          __syncwarp();
          if (++s == nstages) { s = 0; ph ^= 1u; }
        }
        if (alive && ptx::elect_one_sync()) ptx::tc_commit_2cta_multicast_a(tfull_a + 8u * slot, 3);
        __syncwarp();
        if (++slot == NACC) { slot = 0; sph ^= 1u; }
      }
      if (!alive && lane == 0)   // unblock both epilogues so that the CTAs can exit
        for (int b = 0; b < (int)NACC; ++b) { ptx::mbar_arrive(&ctl->tmem_full[b]); ptx::mbar_arrive_remote(&ctl->tmem_full[b], 1); }
    }
  } else {
    // ===== epilogue (both CTAs): this CTA's 128 rows of the 256-row tile =====
    const int quad = warp & 3;
    const int et = threadIdx.x - 64;
    const float rcp = __frcp_rn(p.ep.sc);
    uint32_t tcount = 0, acc_it = 0;
    for (int tile = tile0; tile < mn_tiles; tile += tile_step, ++tcount, ++acc_it) {
      const uint32_t ob = tcount & 1;
      const int n0 = (tile % p.tiles_n) * BN;
      for (int j = et; j < BN; j += 32 * kEpiWarps2) {
        const int n = n0 + j;
        ctl->oc[ob][j] = (n < p.N) ? __ldg(p.ep.oc + n) : 0;
        ctl->bias[ob][j] = 0.f;
        if (p.ep.sb_vec) ctl->sbv[ob][j] = (n < p.N) ? __ldg(p.ep.sb_vec + n) : 1.f;
      }
      epi_bar_sync2();
      const uint32_t slot = acc_it % NACC, sph = (acc_it / NACC) & 1;
      const bool ok = ptx::mbar_wait(&ctl->tmem_full[slot], sph);
      if (!ok) tc_fail(3);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int a = 0; a < MT; ++a) {
        const int mtile = (tile / p.tiles_n) * (2 * MT) + 2 * a + crank;
        const bool tile_valid = mtile < p.tiles_m;
        const int m = min(mtile, p.tiles_m - 1) * BM + quad * 32 + lane;
        const uint32_t t_row = tmem_base + slot * kSlotCols + (uint32_t)a * acc_stride<BN>() + ((uint32_t)(quad * 32) << 16);
        const int32_t* corr = nullptr;
        if (p.border_tab && m < p.M) {
          const int q = m % p.ow, pr = (m / p.ow) % p.oh;
          const int y0 = pr * p.stride_h - p.pad, x0 = q * p.stride_w - p.pad;
          const int th = min(max(-y0, 0), p.kh), bh = min(max(y0 + p.kh - p.H, 0), p.kh);
          const int tw = min(max(-x0, 0), p.kw), bw = min(max(x0 + p.kw - p.W, 0), p.kw);
          const int d = p.pad + 1;
          const int cls = ((th * d + bh) * d + tw) * d + bw;
          if (cls != 0) corr = p.border_tab + (size_t)cls * ((p.N + 31) & ~31);
        }
        epilogue_row<BN, kEpiWarps2 / 4>(p, t_row, (m < p.M && ok && tile_valid) ? (long long)m : -1ll, n0, ptx::smem_u32(ctl->oc[ob]),
                         ptx::smem_u32(ctl->bias[ob]), corr, rcp, (warp - 2) >> 2,
                         p.ep.sb_vec ? ptx::smem_u32(ctl->sbv[ob]) : 0u);
      }
      // hand the accumulator buffer (both halves) back to the even CTA's MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(&ctl->tmem_empty[slot]);
        else ptx::mbar_arrive_remote(&ctl->tmem_empty[slot], 0);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // no CTA leaves while its peer may still use its shared memory, barriers or TMEM
  if (warp == 1) ptx::tmem_dealloc_2cta(tmem_base, tmem_cols<BN>());
}

// One superpixel (4 px x 4 channel lanes = 16 bytes) of the bordered stem image, quantised straight
// from the fp32 NCHW input (quantize_utils.cc:44-52): row y / superpixel sx of the bordered image.
// Split into a load half (all global loads of a superpixel issued back to back, so that a caller
// can put several superpixels in flight) and a quantise half.
// VEC2: pad and w even -> the 4 pixels are two aligned float2 per channel plane.
// FAST: packed fp32x2 quantise behind one magnitude test (quant2_fast).
struct StemPixels {
  float f[4][4];   // [channel lane][pixel]; missing channels / out-of-image pixels stay 0
  uint32_t inside; // bit j: pixel j lies inside the image
};

template <bool VEC2, int CC = 0>   // CC > 0: channel count known at compile time (fewer live registers)
__device__ __forceinline__ void stem_sp_load(StemPixels& sp, const float* __restrict__ x, int img, int y, int sx, int c_rt,
                                             int h, int w, int pad) {
  const int c = CC > 0 ? CC : c_rt;
  const int row = y - pad;
  sp.inside = 0;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch)
#pragma unroll
    for (int j = 0; j < 4; ++j) sp.f[ch][j] = 0.f;
  if (row < 0 || row >= h) return;
  const int64_t plane = (int64_t)h * w;
  const int col0 = sx * 4 - pad;
  const float* src = x + ((int64_t)img * c * h + row) * w + col0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (col0 + j >= 0 && col0 + j < w) sp.inside |= 1u << j;
  if (VEC2) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      if (ch < c) {
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          if (sp.inside & (1u << j)) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(src + ch * plane + j));
            sp.f[ch][j] = v.x; sp.f[ch][j + 1] = v.y;
          }
        }
      }
    }
  } else {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      if (ch < c)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (sp.inside & (1u << j)) sp.f[ch][j] = __ldg(src + ch * plane + j);
  }
}

template <bool FAST, int CC = 0>
__device__ __forceinline__ uint4 stem_sp_quant(const StemPixels& sp, int c_rt, float scale, float zpf, uint32_t zp,
                                               float fast_lim) {
  const int c = CC > 0 ? CC : c_rt;
  const uint32_t zp4 = zp * 0x01010101u;
  uint32_t wd[4] = {zp4, zp4, zp4, zp4};
  if (sp.inside == 0) return make_uint4(zp4, zp4, zp4, zp4);
  int q[4][4];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch)
#pragma unroll
    for (int j = 0; j < 4; ++j) q[ch][j] = (int)zp;
  bool done = false;
  if (FAST) {
    float amax = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
      for (int j = 0; j < 4; ++j) amax = fmaxf(amax, fabsf(sp.f[ch][j]));
    if (amax < fast_lim) {
      const QuantFast2 qc = make_quant_fast2(scale, __frcp_rn(scale), zpf);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if (ch < c) {
          quant2_fast(sp.f[ch][0], sp.f[ch][1], qc, q[ch][0], q[ch][1]);
          quant2_fast(sp.f[ch][2], sp.f[ch][3], qc, q[ch][2], q[ch][3]);
        }
      }
      done = true;
    }
  }
  if (!done) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      if (ch < c)
#pragma unroll
        for (int j = 0; j < 4; ++j) q[ch][j] = (int)quant_u8_wrap(sp.f[ch][j], scale, zpf);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (sp.inside & (1u << j)) wd[j] = pack_low_bytes(q[0][j], q[1][j], q[2][j], q[3][j]);
  return make_uint4(wd[0], wd[1], wd[2], wd[3]);
}

template <bool VEC2, bool FAST>
__device__ __forceinline__ uint4 stem_superpixel(const float* __restrict__ x, int img, int y, int sx, int c, int h,
                                                 int w, int pad, float scale, float zpf, uint32_t zp,
                                                 float fast_lim) {
  StemPixels sp;
  stem_sp_load<VEC2>(sp, x, img, y, sx, c, h, w, pad);
  return stem_sp_quant<FAST>(sp, c, scale, zpf, zp, fast_lim);
}

// The whole bordered stem image in one pass (used when the stem kernel cannot fuse the quantise).
// One thread per superpixel; grid.y = image, so the only per-thread division is a 32-bit one.
template <bool VEC2, bool FAST>
__global__ void __launch_bounds__(256) stem_quantize_kernel(const float* __restrict__ x, uint8_t* __restrict__ xs,
                                                            int c, int h, int w, int pad, int hp, int wsp,
                                                            float scale, float zpf, uint32_t zp, float fast_lim,
                                                            const float* const* __restrict__ xslot) {
  pdl_launch_dependents();
  pdl_wait();
  if (xslot) x = *xslot;   // run-time source address (CUDA-graph replay on a new input buffer)
  const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= (uint32_t)(hp * wsp)) return;
  const int img = blockIdx.y;
  const int y = (int)(idx / (uint32_t)wsp), sx = (int)(idx - (uint32_t)y * (uint32_t)wsp);
  const uint4 v = stem_superpixel<VEC2, FAST>(x, img, y, sx, c, h, w, pad, scale, zpf, zp, fast_lim);
  *reinterpret_cast<uint4*>(xs + (((int64_t)img * hp + y) * wsp + sx) * 16) = v;
}

// ---- stem2: smem-resident stem rows, sliding windows expressed by overlapping descriptors ---
// For stride-4 stems the window of output pixel q starts at superpixel q, i.e. 16 bytes after
// the window of q-1: exactly the fixed row pitch of a NO-SWIZZLE K-major UMMA core matrix
// (8 rows x 16 B, rows 16 B apart). So the raw stem rows are copied to shared memory ONCE and
// the A descriptor (LBO = 16 B between the two 16-byte K chunks, SBO = 128 B between 8-row
// groups) walks them as 64 overlapping windows — no im2col copy, no duplicated fetch.
// An M=128 tile is two output rows (p, p+1) x 64 columns: input rows 4 apart are stored
// exactly 1024 B apart (region = row % 4, slot = row / 4), which keeps SBO uniform across the
// switch from output row p to p+1. The whole packed weight [kh][BN x 64 B] stays resident.
struct Stem2Params {
  int n_img, oh, ow, kh, hp, wsp;
  int pairs;        // ceil(oh / 2) output-row pairs per image
  int nsl;          // 1 KB slots per region = ceil((kh + 4) / 4)
  int stages;
  int dbg;          // dev-only bottleneck probes (I8IE_STEM2_DBG): 1 = no epilogue, 2 = no MMA, 4 = no loads, 8 = no stores, 16 = no quantise
  const uint8_t* xs;
  // fused input quantise (FQ kernels): the fp32 NCHW image, read directly by the producer warps
  const float* xf;
  const float* const* xslot;
  int c, h, w, pad;
  float in_scale, fast_lim;
  int in_zp;
  // K packing (kw <= 12): the window of one filter row is 3 superpixels = 48 bytes, so the GEMM K axis is the
  // concatenation of kh 48-byte windows (kh * 48 bytes, rounded up to 32) instead of kh 64-byte ones: AlexNet conv1
  // issues 17 MMAs per tile instead of 22. K32 block j covers bytes [32j, 32j + 32) of that axis = two 16-byte
  // chunks that either lie next to each other in one stem row (LBO = 16) or straddle two filter rows (chunk 0 =
  // superpixel 2 of row r, chunk 1 = superpixel 0 of row r + 1: LBO = their distance in the operand stage).
  // k48_a[j] = (chunk-0 offset in the stage >> 4) | (LBO >> 4) << 16, added to the stage's descriptor low word.
  int k48_n;                     // K32 blocks per tile (0 = 64-byte windows)
  uint32_t k48_a[24];
  int f_stages, f_stage_bytes;   // fp32 row ring: [rows_per_tile][c][w] floats per stage
  int pf_tiles;                  // L2 prefetch distance of the fp32 rows, in tiles (0 = off)
  // dev-only timeline (build with I8IE_NVCC_EXTRA=-DI8IE_STEM_TRACE, run with I8IE_STEM2_TRACE=<file>): CTA 0
  // records clock64() at kTraceEvents points of each of its first kTraceTiles tiles — trace[tile][event];
  // compiled out of the production library (the per-tile checks cost 2.4 % of the kernel's instructions)
  long long* trace;
};
constexpr int kTraceTiles = 24, kTraceEvents = 16;
__device__ __forceinline__ void stem_trace(const Stem2Params& sp, uint32_t it, int ev) {
#ifdef I8IE_STEM_TRACE
  if (sp.trace != nullptr && blockIdx.x == 0 && it < (uint32_t)kTraceTiles) sp.trace[it * kTraceEvents + ev] = clock64();
#else
  (void)sp; (void)it; (void)ev;
#endif
}

// Per-role tile / ring cursors of the stem kernel, advanced incrementally: every role used to divide the
// tile index by `pairs` and the tile counter by the ring depths once per tile (runtime divisors: ~25
// instructions each through the slow conversion pipes), ~150 instructions per warp per tile x 22 warps.
struct StemCursor {
  int img, pr;          // image, output-row pair within the image
  uint32_t s, ph;       // ring slot and phase bit
  __device__ __forceinline__ void init(int tile, int pairs) { img = tile / pairs; pr = tile - img * pairs; s = 0; ph = 0; }
  __device__ __forceinline__ void next_tile(int pairs) { if (++pr == pairs) { pr = 0; ++img; } }
  __device__ __forceinline__ void next_slot(uint32_t n) { if (++s == n) { s = 0; ph ^= 1u; } }
};

// FQ: the input quantise is fused. Warp 0 bulk-copies the raw fp32 image rows of a tile
// (cp.async.bulk, one row of one channel plane per copy) into a second shared-memory ring, 8
// converter warps quantise them into superpixel rows of the operand ring (generic-proxy stores +
// fence.proxy.async), so the global-load latency hides behind the ring and the bordered u8 stem
// image is never materialised: 21 MB written + 39 MB re-read per 100 images disappear, and so does
// the stem_quantize launch.
template <int BN, int KH, bool FQ>   // KH > 0: filter height known at compile time (fully unrolled issue loop)
__global__ void __launch_bounds__((stem_threads<BN, FQ>()), 1) tc_stem2_kernel(const __grid_constant__ CUtensorMap tmB,
                                                               const TcParams p, const Stem2Params sp) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // resident weights: one [BN x 64 B] SW64 tile per filter row, or k48_n [BN x 32 B] SW32 tiles (K packing)
  const int w_bytes = sp.k48_n > 0 ? sp.k48_n * BN * 32 : sp.kh * BN * 64;
  const int a_stage = 4 * sp.nsl * 1024;
  uint8_t* sW = smem;
  uint8_t* sA = smem + w_bytes;
  uint8_t* sF = sA + (size_t)sp.stages * a_stage;   // FQ: fp32 row ring
  TcControl<BN>* ctl = reinterpret_cast<TcControl<BN>*>(sF + (FQ ? (size_t)sp.f_stages * sp.f_stage_bytes : 0));
  uint64_t* w_full = &ctl->full[kMaxStages - 1];   // the ring never uses more than kMaxStages-1 slots here
  uint64_t* f_full = &ctl->full[kStemFBar];        // fp32 ring barriers (operand ring: slots < kStemFBar)
  uint64_t* f_empty = &ctl->empty[kStemFBar];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = sp.n_img * sp.pairs;
  constexpr int kEpiW = stem_epi_warps<BN, FQ>();                  // epilogue warps: 2 .. 2 + kEpiW
  constexpr int kProdW = FQ ? kStemProdWarps : 0;                  // producer warps (FQ): the rest

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < sp.stages; ++s) {
      ptx::mbar_init(&ctl->full[s], FQ ? kProdW : 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    ptx::mbar_init(w_full, 1);
    if (FQ)
      for (int s = 0; s < sp.f_stages; ++s) {
        ptx::mbar_init(&f_full[s], 1);
        ptx::mbar_init(&f_empty[s], kProdW);
      }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&ctl->tmem_full[b], 1);
      ptx::mbar_init(&ctl->tmem_empty[b], kEpiW);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(&ctl->tmem_slot, tmem_cols<BN>());
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_slot;
  pdl_wait();
  const int rows_per_tile = sp.kh + 4;
  // Tile sequence of this CTA. FQ: one contiguous run of tiles, so that consecutive tiles of an image
  // share their overlapping kh - 4 rows (copied inside shared memory instead of loaded and quantised
  // again); otherwise static striding.
  const int t_begin = FQ ? (int)((long long)blockIdx.x * num_tiles / gridDim.x) : (int)blockIdx.x;
  const int t_end = FQ ? (int)((long long)(blockIdx.x + 1) * num_tiles / gridDim.x) : num_tiles;
  const int t_step = FQ ? 1 : (int)gridDim.x;
  const int i_new = sp.kh > 4 ? sp.kh - 4 : 0;   // FQ: first tile row a follow-up tile has to produce itself
  const uint32_t row_bytes = (uint32_t)sp.wsp * 16;

  if (warp == 0) {
    // ===== producer: the whole warp walks the tile loop, lane i copies row i of the tile (one
    // bulk copy each, issued in parallel instead of 15 back-to-back from one thread) =====
    if (lane == 0) {   // weights once
      ptx::mbar_arrive_expect_tx(w_full, (uint32_t)w_bytes);
      if (sp.k48_n > 0)
        for (int j = 0; j < sp.k48_n; ++j) ptx::tma_load_2d(sW + (size_t)j * BN * 32, &tmB, w_full, j * 32, 0);
      else
        for (int r = 0; r < sp.kh; ++r) ptx::tma_load_2d(sW + (size_t)r * BN * 64, &tmB, w_full, r * 64, 0);
    }
    uint32_t it = 0;
    bool alive = true;
    if (FQ) {
      // fp32 rows of the tile -> fp32 ring; rows above / below the image are skipped (the converter
      // warps emit zero-point rows for them)
      const float* xf = sp.xslot ? *sp.xslot : sp.xf;
      const uint32_t frow_bytes = (uint32_t)sp.w * 4;
      // The image comes from DRAM: lane ch prefetches channel plane ch of the rows of a LATER tile into L2
      // (see ptx::prefetch_l2_bulk), so that the ring's bulk copies hit L2 when their turn comes.
      const int pf_ahead = sp.pf_tiles;
      auto prefetch_tile = [&](int t) {
        if (t >= t_end || lane >= sp.c) return;
        const int pimg = t / sp.pairs, pp0 = (t % sp.pairs) * 2;
        const int pr0 = 4 * pp0 - sp.pad;
        int plo = pr0 < 0 ? -pr0 : 0, phi = sp.h - pr0;
        if (phi > rows_per_tile) phi = rows_per_tile;
        if (!(t == t_begin || pp0 == 0) && plo < i_new) plo = i_new;
        if (phi > plo)
          ptx::prefetch_l2_bulk(xf + (((int64_t)pimg * sp.c + lane) * sp.h + (pr0 + plo)) * sp.w, (uint32_t)(phi - plo) * frow_bytes);
      };
      for (int t = t_begin; t < t_begin + pf_ahead; ++t) prefetch_tile(t);
      StemCursor cur;
      cur.init(t_begin, sp.pairs);
      for (int tile = t_begin; tile < t_end && alive; tile += t_step, ++it, cur.next_tile(sp.pairs), cur.next_slot((uint32_t)sp.f_stages)) {
        const int img = cur.img, p0 = cur.pr * 2;
        const uint32_t s = cur.s, ph = cur.ph;
        if (pf_ahead > 0) prefetch_tile(tile + pf_ahead);
        if (!ptx::mbar_wait(&f_empty[s], ph ^ 1)) { tc_fail(1); alive = false; break; }
        if (lane == 0) stem_trace(sp, it, 0);
        const int r0 = 4 * p0 - sp.pad;                 // image row of tile row 0
        int lo = r0 < 0 ? -r0 : 0, hi = sp.h - r0;      // tile rows [lo, hi) lie inside the image
        if (hi > rows_per_tile) hi = rows_per_tile;
        const bool first = tile == t_begin || p0 == 0;  // no previous tile of the same image in this CTA
        if (!first && lo < i_new) lo = i_new;           // rows below i_new are carried over by the converters
        if (hi < lo) hi = lo;
        // NCHW: the rows [lo, hi) of one channel plane are one contiguous run -> one bulk copy per plane
        // (stage layout [c][rows_per_tile][w] floats)
        const uint32_t run = (sp.dbg & 4) ? 0u : (uint32_t)(hi - lo) * frow_bytes;
        if (lane == 0) ptx::mbar_arrive_expect_tx(&f_full[s], run * (uint32_t)sp.c);
        uint8_t* st = sF + (size_t)s * sp.f_stage_bytes;
        if (lane < sp.c && run != 0)
          ptx::bulk_load_1d(st + (size_t)(lane * rows_per_tile + lo) * frow_bytes,
                            xf + (((int64_t)img * sp.c + lane) * sp.h + (r0 + lo)) * sp.w, run, &f_full[s]);
        __syncwarp();
      }
      alive = false;
    }

    for (int tile = t_begin; tile < t_end && alive; tile += t_step, ++it) {
      const int img = tile / sp.pairs, p0 = (tile % sp.pairs) * 2;
      const uint32_t s = it % (uint32_t)sp.stages;
      const uint32_t ph = (it / (uint32_t)sp.stages) & 1;
      if (!ptx::mbar_wait(&ctl->empty[s], ph ^ 1)) { tc_fail(1); alive = false; break; }
      const int i0 = 4 * p0;
      int nrows = sp.hp - i0;
      if (nrows > rows_per_tile) nrows = rows_per_tile;
      if (sp.dbg & 4) { if (lane == 0) ptx::mbar_arrive(&ctl->full[s]); __syncwarp(); continue; }
      // the arrival (with the byte count) and the copies may land in any order: the phase cannot
      // complete before the single arrival, and by then the transaction count is balanced
      if (lane == 0) ptx::mbar_arrive_expect_tx(&ctl->full[s], (uint32_t)nrows * row_bytes);
      uint8_t* st = sA + (size_t)s * a_stage;
      const uint8_t* src = sp.xs + ((size_t)img * sp.hp + i0) * row_bytes;
      for (int i = lane; i < nrows; i += 32)
        ptx::bulk_load_1d(st + (size_t)((i & 3) * sp.nsl + (i >> 2)) * 1024, src + (size_t)i * row_bytes, row_bytes,
                          &ctl->full[s]);
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer. The whole warp walks the loop so that every address below is warp-uniform
    // (descriptor words live in uniform registers, no per-MMA broadcast); lane 0 issues. =====
    constexpr uint32_t idesc = ptx::make_idesc_i8(BM, BN);
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
    // A: no-swizzle K-major, LBO = 16 B (the two K chunks), SBO = 128 B (8-row groups), version 1
    constexpr uint32_t a_flags = (16u >> 4) << 16, a_hi = (128u >> 4) | (1u << 14);
    const uint32_t b_hi = ptx::smem_desc_hi<64>();
    const uint32_t b_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sW));
    const uint32_t sa_base = ptx::smem_u32(sA);
    bool alive = ptx::mbar_wait(w_full, 0);
    if (!alive) tc_fail(5);
    uint32_t it = 0;
    StemCursor mcur;
    mcur.s = 0; mcur.ph = 0;
    for (int tile = t_begin; tile < t_end && alive; tile += t_step, ++it) {
      const uint32_t buf = it & 1, bph = (it >> 1) & 1;
      const uint32_t s = mcur.s, ph = mcur.ph;
      mcur.next_slot((uint32_t)sp.stages);
      if (!ptx::mbar_wait(&ctl->tmem_empty[buf], bph ^ 1)) { tc_fail(4); alive = false; break; }
      if (lane == 0) stem_trace(sp, it, 8);
      if (!ptx::mbar_wait(&ctl->full[s], ph)) { tc_fail(2); alive = false; break; }
      if (lane == 0) stem_trace(sp, it, 9);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tbase + buf * acc_stride<BN>();
      const uint32_t a_lo0 = (((sa_base + s * (uint32_t)a_stage) & 0x3FFFFu) >> 4) | a_flags;
      const bool leader = ptx::elect_one_sync();
      if (leader && !(sp.dbg & 2) && sp.k48_n > 0) {
        const uint32_t a_base = ((sa_base + s * (uint32_t)a_stage) & 0x3FFFFu) >> 4;
        const uint32_t b32_hi = ptx::smem_desc_hi<32>();
        for (int j = 0; j < sp.k48_n; ++j)
          ptx::mma_i8_ss_lohi(d_tmem, a_base + sp.k48_a[j], a_hi, b_lo0 + (uint32_t)(j * BN * 2), b32_hi, idesc, j != 0 ? 1u : 0u);
      } else if (leader && !(sp.dbg & 2)) {
        if (KH > 0) {
          constexpr int NSL = (KH + 4 + 3) / 4;
#pragma unroll
          for (int r = 0; r < KH; ++r) {
            constexpr int kDummy = 0; (void)kDummy;
            const uint32_t a_lo = a_lo0 + (uint32_t)(((r & 3) * NSL + (r >> 2)) * 64);   // 1 KB slots, >> 4
            const uint32_t b_lo = b_lo0 + (uint32_t)(r * BN * 4);                        // BN x 64 B per filter row
            ptx::mma_i8_ss_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, r != 0 ? 1u : 0u);
            ptx::mma_i8_ss_lohi(d_tmem, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
          }
        } else {
          for (int r = 0; r < sp.kh; ++r) {
            const uint32_t a_lo = a_lo0 + (uint32_t)(((r & 3) * sp.nsl + (r >> 2)) * 64);
            const uint32_t b_lo = b_lo0 + (uint32_t)(r * BN * 4);
            ptx::mma_i8_ss_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, r != 0 ? 1u : 0u);
            ptx::mma_i8_ss_lohi(d_tmem, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
          }
        }
      }
      if (leader) {
        ptx::tc_commit(&ctl->empty[s]);
        ptx::tc_commit(&ctl->tmem_full[buf]);
      }
      if (lane == 0) stem_trace(sp, it, 10);
      __syncwarp();
    }
    if (!alive && lane == 0) {
      ptx::mbar_arrive(&ctl->tmem_full[0]);
      ptx::mbar_arrive(&ctl->tmem_full[1]);
    }
  } else if (FQ && warp >= 2 + kEpiW) {
    // ===== converter warps (fused quantise): fp32 ring -> u8 superpixel rows in the operand ring.
    // Rows a tile shares with its predecessor (same image, same CTA) are copied from the previous
    // operand stage; of the new rows, warp pw takes every kProdW-th, and a lane owns superpixels
    // lane and lane + 32 (wsp <= 64), processed together for instruction-level parallelism.
    // RGB, even pad / width (two float2 per channel plane and superpixel). =====
    const int pw = warp - 2 - kEpiW;
    const float zpf = (float)sp.in_zp;
    const QuantFast2 qc = make_quant_fast2(sp.in_scale, __frcp_rn(sp.in_scale), zpf);
    const uint32_t zp4 = (uint32_t)sp.in_zp * 0x01010101u;
    const uint32_t sA_u = ptx::smem_u32(sA), sF_u = ptx::smem_u32(sF);
    // per-lane column state of the two superpixels (same for every row)
    bool in01[2], in23[2], sx_ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int sx = lane + 32 * u;
      const int col0 = sx * 4 - sp.pad;
      sx_ok[u] = sx < sp.wsp;
      in01[u] = sx_ok[u] && col0 >= 0 && col0 < sp.w;            // pixels 0,1 (pad and w are even)
      in23[u] = sx_ok[u] && col0 + 2 >= 0 && col0 + 2 < sp.w;    // pixels 2,3
    }
    uint32_t it = 0;
    bool dead = false;
    StemCursor ccur, fcur;
    ccur.init(t_begin, sp.pairs);
    fcur.s = 0; fcur.ph = 0;
    uint32_t sprev_slot = 0, sprev_ph = 0;
    for (int tile = t_begin; tile < t_end; tile += t_step, ++it, ccur.next_tile(sp.pairs)) {
      const int p0 = ccur.pr * 2;
      const uint32_t s = ccur.s, ph = ccur.ph;
      const uint32_t fs = fcur.s, fph = fcur.ph;
      ccur.next_slot((uint32_t)sp.stages);
      fcur.next_slot((uint32_t)sp.f_stages);
      const int i0 = 4 * p0;
      int nrows = sp.hp - i0;
      if (nrows > rows_per_tile) nrows = rows_per_tile;
      const bool first = tile == t_begin || p0 == 0;
      // (a timed-out warp keeps walking the loop without waiting, so that nobody hangs in bar.sync)
      if (!dead && !ptx::mbar_wait(&f_full[fs], fph)) { tc_fail(6); dead = true; }
      if (pw == 0 && lane == 0) stem_trace(sp, it, 1);
      if (!dead && !ptx::mbar_wait(&ctl->empty[s], ph ^ 1)) { tc_fail(1); dead = true; }
      if (pw == 0 && lane == 0) stem_trace(sp, it, 2);
      // every converter warp is done with the previous tile: its rows can be read, and the stage
      // before it (this tile's target) is no longer being read by a slower warp's copy
      asm volatile("bar.sync 2, %0;" ::"n"(32 * kProdW) : "memory");
      if (pw == 0 && lane == 0) stem_trace(sp, it, 3);
      const uint32_t st = sA_u + s * (uint32_t)a_stage;
      const uint32_t sf = sF_u + fs * (uint32_t)sp.f_stage_bytes;
      int i_start = 0;
      // The previous tile's `full` phase is complete — i.e. ITS bulk copy (below) has finished reading the stage this
      // tile is about to overwrite — before any row of this tile is written.
      if (it != 0 && !dead && !ptx::mbar_wait(&ctl->full[sprev_slot], sprev_ph)) { tc_fail(6); dead = true; }
      if (!first) {
        // Tile row i is row i + 8 of the previous tile: same region (i & 3), two 1 KB slots further on. The carried
        // rows of a region are contiguous in both stages, so ONE thread hands them to the bulk-copy engine (shared ->
        // shared, completing on this stage's `full` barrier next to the converters' arrivals) instead of every warp
        // moving a row through registers: those lds / sts stalled for ~1500 clk per tile behind the operand reads of
        // the previous tile's MMAs, on the converters' critical path.
        const uint32_t sprev = sA_u + sprev_slot * (uint32_t)a_stage;
        if (pw == 0 && lane == 0 && !dead) {
          uint32_t bytes = 0;
          for (int r = 0; r < 4 && r < i_new; ++r) bytes += (uint32_t)((i_new - r + 3) >> 2) * 1024u;
          ptx::mbar_expect_tx(&ctl->full[s], bytes);
          const uint32_t bar = ptx::smem_u32(&ctl->full[s]);
          for (int r = 0; r < 4 && r < i_new; ++r) {
            const uint32_t cnt = (uint32_t)((i_new - r + 3) >> 2);      // carried rows r, r + 4, ... < i_new
            ptx::bulk_copy_to_cluster(st + (uint32_t)(r * sp.nsl) * 1024u, sprev + (uint32_t)(r * sp.nsl + 2) * 1024u,
                                      cnt * 1024u, bar);
          }
        }
        i_start = i_new;
      }
      if (pw == 0 && lane == 0) stem_trace(sp, it, 4);
      for (int i = i_start + pw; i < nrows && !(sp.dbg & 16); i += kProdW) {
        const int row = i0 + i - sp.pad;
        const bool row_ok = row >= 0 && row < sp.h;
        const uint32_t rp = sf + (uint32_t)((i * sp.w - sp.pad) * 4);   // channel plane 0 of tile row i, bordered column 0
        const uint32_t drow = st + (uint32_t)((i & 3) * sp.nsl + (i >> 2)) * 1024u;
        float2 v[2][3][2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const uint32_t src = rp + (uint32_t)((ch * rows_per_tile * sp.w + (lane + 32 * u) * 4) * 4);
            v[u][ch][0] = (row_ok && in01[u]) ? ptx::lds64_f2(src) : make_float2(0.f, 0.f);
            v[u][ch][1] = (row_ok && in23[u]) ? ptx::lds64_f2(src + 8) : make_float2(0.f, 0.f);
          }
        float amax = 0.f;
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int ch = 0; ch < 3; ++ch)
            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v[u][ch][0].x), fabsf(v[u][ch][0].y))),
                         fmaxf(fabsf(v[u][ch][1].x), fabsf(v[u][ch][1].y)));
        int q[2][3][4];
        if (amax < sp.fast_lim) {
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              quant2_fast(v[u][ch][0].x, v[u][ch][0].y, qc, q[u][ch][0], q[u][ch][1]);
              quant2_fast(v[u][ch][1].x, v[u][ch][1].y, qc, q[u][ch][2], q[u][ch][3]);
            }
        } else {   // also taken for NaN (the compare is false)
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              q[u][ch][0] = (int)quant_u8_wrap(v[u][ch][0].x, sp.in_scale, zpf);
              q[u][ch][1] = (int)quant_u8_wrap(v[u][ch][0].y, sp.in_scale, zpf);
              q[u][ch][2] = (int)quant_u8_wrap(v[u][ch][1].x, sp.in_scale, zpf);
              q[u][ch][3] = (int)quant_u8_wrap(v[u][ch][1].y, sp.in_scale, zpf);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint32_t wd[4] = {zp4, zp4, zp4, zp4};
          if (row_ok && in01[u]) {
            wd[0] = pack_low_bytes(q[u][0][0], q[u][1][0], q[u][2][0], sp.in_zp);
            wd[1] = pack_low_bytes(q[u][0][1], q[u][1][1], q[u][2][1], sp.in_zp);
          }
          if (row_ok && in23[u]) {
            wd[2] = pack_low_bytes(q[u][0][2], q[u][1][2], q[u][2][2], sp.in_zp);
            wd[3] = pack_low_bytes(q[u][0][3], q[u][1][3], q[u][2][3], sp.in_zp);
          }
          if (sx_ok[u]) ptx::sts128(drow + (uint32_t)(lane + 32 * u) * 16u, make_uint4(wd[0], wd[1], wd[2], wd[3]));
        }
      }
      if (pw == 0 && lane == 0) stem_trace(sp, it, 5);
      // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&ctl->full[s]);
        ptx::mbar_arrive(&f_empty[fs]);
      }
      if (pw == 0 && lane == 0) stem_trace(sp, it, 6);
      sprev_slot = s; sprev_ph = ph;
    }
  } else {
    const int quad = warp & 3;
    const int et = threadIdx.x - 64;
    const float rcp = __frcp_rn(p.ep.sc);
    // per-channel offsets are the same for every tile (single N tile)
    for (int j = et; j < BN; j += 32 * kEpiW) {
      ctl->oc[0][j] = (j < p.N) ? __ldg(p.ep.oc + j) : 0;
      ctl->bias[0][j] = 0.f;
      if (p.ep.sb_vec) ctl->sbv[0][j] = (j < p.N) ? __ldg(p.ep.sb_vec + j) : 1.f;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiW) : "memory");
    uint32_t it = 0;
    StemCursor ecur;
    ecur.init(t_begin, sp.pairs);
    for (int tile = t_begin; tile < t_end; tile += t_step, ++it) {
      const uint32_t buf = it & 1, bph = (it >> 1) & 1;
      // FQ: contiguous tile run (cursor); otherwise tiles are strided over the grid
      const int img = FQ ? ecur.img : tile / sp.pairs, p0 = (FQ ? ecur.pr : tile % sp.pairs) * 2;
      if (FQ) ecur.next_tile(sp.pairs);
      const int l = quad * 32 + lane;            // TMEM lane = tile row
      const int prow = p0 + (l >> 6), q = l & 63;
      const bool valid = (prow < sp.oh) && (q < sp.ow);
      const long long m = valid ? ((long long)img * sp.oh + prow) * sp.ow + q : -1ll;
      const bool ok = ptx::mbar_wait(&ctl->tmem_full[buf], bph);
      if (!ok) tc_fail(3);
      if (warp == 2 && lane == 0) stem_trace(sp, it, 12);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + buf * acc_stride<BN>() + ((uint32_t)(quad * 32) << 16);
      if (!(sp.dbg & 1)) {
        epilogue_row<BN, kEpiW / 4>(p, t_row, (ok && !(sp.dbg & 8)) ? m : -1ll, 0, ptx::smem_u32(ctl->oc[0]),
                                    ptx::smem_u32(ctl->bias[0]), nullptr, rcp, (warp - 2) >> 2,
                                    p.ep.sb_vec ? ptx::smem_u32(ctl->sbv[0]) : 0u);
        // output pitch wider than the N tile (e.g. 96 channels stored at pitch 128): the pad lanes
        // carry the zero point; written by the warp of each pair that had fewer chunks
        if (BN < p.out_cp && ((warp - 2) >> 2) == ((BN / 32) % (kEpiW / 4)) && ok && m >= 0) {
          const uint32_t z4 = (uint32_t)p.ep.zp_out * 0x01010101u;
          for (int c = BN; c < p.out_cp; c += 16)
            *reinterpret_cast<uint4*>(p.y + (size_t)m * p.out_cp + c) = make_uint4(z4, z4, z4, z4);
        }
      }
      if (warp == 2 && lane == 0) stem_trace(sp, it, 13);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ctl->tmem_empty[buf]);
      if (warp == 2 && lane == 0) stem_trace(sp, it, 14);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, tmem_cols<BN>());
}

// ---- row mode (stride-1 convs whose channel count is not a multiple of 128) ----------------------
// With the input stored PHYSICALLY padded as [n][h + 2p][w + 2p][cpx] (border = zero point), the kw taps of
// one filter row of one output pixel are ONE contiguous run of kw * cpx bytes. The GEMM K index becomes
// (filter row, byte of the run) with the run rounded up to KR = a multiple of 128 bytes (the tail reads the
// neighbouring pixel against zero weights), instead of (tap, channel) with every tap padded to 128 channels:
// AlexNet conv2 (C = 96, 5 x 5) runs 5 x 512 = 2560 bytes of K per pixel instead of 25 x 128 = 3200, and needs no
// zero-point border table (the padding is real, exactly conv2d.cc:24-25). The kernel is unchanged: the TMA
// im2col map describes a virtual tensor {KR "channels", ow positions at a pitch of cpx bytes, h + 2p rows, n}
// with a kh x 1 filter — overlapping "pixels", as in the stem view.
// Weights: wr[n][r][KR], byte (x * cpx + ch) of row r = w_packed[n][r][x][ch], zero beyond kw * cpx.
__global__ void row_weight_kernel(const int8_t* __restrict__ wp, int8_t* __restrict__ wr, int kc_pad, int kh, int kw,
                                  int cp, int kr) {
  const int64_t total = (int64_t)kc_pad * kh * kr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i % kr);
    const int64_t nr = i / kr;   // n * kh + r
    wr[i] = b < kw * cp ? wp[nr * kw * cp + b] : (int8_t)0;
  }
}

// zero-point border correction table: tab[cls][n] = sum over spatially out-of-range taps of
// the packed weight, cls = ((th*(pad+1)+bh)*(pad+1)+tw)*(pad+1)+bw  (rows [0,th) and
// [kh-bh,kh), cols [0,tw) and [kw-bw,kw) are out of range).
__global__ void border_table_kernel(const int8_t* __restrict__ wp, int32_t* __restrict__ tab, int N, int kh,
                                    int kw, int cp, int pad) {
  const int d = pad + 1;
  const int ncls = d * d * d * d;
  const int npitch = (N + 31) & ~31;   // rows padded to 32 channels (zeros) for 128-bit lookups
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ncls * npitch) return;
  const int n = idx % npitch, cls = idx / npitch;
  const int bw = cls % d, tw = (cls / d) % d, bh = (cls / (d * d)) % d, th = cls / (d * d * d);
  int32_t s = 0;
  if (n < N) {
    for (int r = 0; r < kh; ++r)
      for (int c = 0; c < kw; ++c) {
        const bool inb = (r >= th) && (r < kh - bh) && (c >= tw) && (c < kw - bw);
        if (inb) continue;
        const int8_t* w = wp + (((size_t)n * kh + r) * kw + c) * cp;
        for (int ch = 0; ch < cp; ++ch) s += w[ch];
      }
  }
  tab[idx] = s;
}

// ---- stem path: small-C strided first-layer convs (AlexNet conv1: C=3, k=11, s=4, p=2) ------
// A 3-channel NHWC image gives TMA/UMMA nothing to chew on (K blocks must be >= 32 bytes).
// The stem layout packs 4 pixels x 4 channel lanes into one 16-byte "superpixel" and bakes the
// spatial zero-point border in physically (the window start 4*ox - pad must land on a
// superpixel boundary, so the image is stored shifted by `pad`):
//     xs[n][y][wsp][16],  y in [0, hp), hp = (oh-1)*s + kh,  wsp = (ow-1)*s/4 + 4
//     pixel (y, x) of the bordered image lives at superpixel x/4, lanes 4*(x%4) + ch
// One filter row of an output pixel is then ONE contiguous 64-byte run (16 px x 4 lanes, the
// first kw px used), fetched by an im2col tensor map whose W stride is 16 bytes (windows
// overlap). GEMM K = kh * 64.
__global__ void stem_pack_kernel(const uint8_t* __restrict__ x, uint8_t* __restrict__ xs, int n, int c, int h,
                                 int w, int cp, int pad, int hp, int wsp, uint32_t zp) {
  const int64_t total = (int64_t)n * hp * wsp;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int sx = (int)(t % wsp);
    const int y = (int)((t / wsp) % hp);
    const int img = (int)(t / ((int64_t)wsp * hp));
    const int row = y - pad;
    const uint32_t zp4 = zp * 0x01010101u;
    uint32_t wd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = sx * 4 + j - pad;
      uint32_t v = zp4;
      if (row >= 0 && row < h && col >= 0 && col < w) {
        const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(x + (((int64_t)img * h + row) * w + col) * cp));
        const uint32_t keep = (c >= 4) ? 0xffffffffu : ((1u << (8 * c)) - 1u);
        v = (raw & keep) | (zp4 & ~keep);
      }
      wd[j] = v;
    }
    *reinterpret_cast<uint4*>(xs + t * 16) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
  }
}

// stem weights: ws[n][r][64], lane j = 4*px + ch
__global__ void stem_weight_kernel(const int8_t* __restrict__ wp, int8_t* __restrict__ ws, int kc_pad, int c,
                                   int kh, int kw, int cp) {
  const int total = kc_pad * kh * 64;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = idx % 64, r = (idx / 64) % kh, n = idx / (64 * kh);
  const int px = j >> 2, ch = j & 3;
  ws[idx] = (px < kw && ch < c) ? wp[(((size_t)n * kh + r) * kw + px) * cp + ch] : (int8_t)0;
}


// stem weights, K-packed: ws[n][k48], byte 48 * r + 4 * px + ch of the K axis (px < 12), zero beyond kh * 48
__global__ void stem_weight48_kernel(const int8_t* __restrict__ wp, int8_t* __restrict__ ws, int kc_pad, int c,
                                     int kh, int kw, int cp, int k48) {
  const int total = kc_pad * k48;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int b = idx % k48, n = idx / k48;
  const int r = b / 48, o = b % 48, px = o >> 2, ch = o & 3;
  ws[idx] = (r < kh && px < kw && ch < c) ? wp[(((size_t)n * kh + r) * kw + px) * cp + ch] : (int8_t)0;
}

// split-K finish: y[m][n] = requant(sum_s ws[s][m][n] + oc[n] (+ float bias)).
// One thread per (row, 4 channels): a warp reads 512 contiguous bytes of every split (fully
// coalesced), eight independent 128-bit loads in flight per thread. Integer adds commute, so
// the reduction order is irrelevant to the result.
// fc weights [n_pad][ldw] (K-major rows) -> [n_pad / 128][ldw / 128] blocks of 128 rows x 128 bytes in the
// SWIZZLE_128B shared-memory image: 16-byte chunk c of row r sits at r * 128 + ((c ^ (r & 7)) * 16).
__global__ void fc_tile_weight_kernel(const int8_t* __restrict__ w, int8_t* __restrict__ wt, int n_pad, int ldw) {
  const int nkb = ldw >> 7;
  const int64_t total = (int64_t)n_pad * (ldw >> 4);   // 16-byte chunks
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / (ldw >> 4)), ch = (int)(i % (ldw >> 4));
    const int kb = ch >> 3, c = ch & 7, r = row & 127, nt = row >> 7;
    const uint4 v = *reinterpret_cast<const uint4*>(w + (size_t)row * ldw + (size_t)ch * 16);
    *reinterpret_cast<uint4*>(wt + ((size_t)nt * nkb + kb) * (128 * 128) + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
}

__global__ void __launch_bounds__(256) fc_splitk_reduce_kernel(const int32_t* __restrict__ ws, int splits, int M,
                                                               int N, int ws_ld, int ldy, uint8_t* __restrict__ y,
                                                               const EpiParams ep, int fast) {
  pdl_launch_dependents();
  pdl_wait();
  const int quads = ldy >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * quads) return;
  const int m = (int)(idx / quads), n4 = (int)(idx % quads) * 4;
  const size_t split_stride = (size_t)M * ws_ld;
  const int32_t* src = ws + (size_t)m * ws_ld + n4;
  int4 a = make_int4(0, 0, 0, 0);
  int s = 0;
  for (; s + 8 <= splits; s += 8) {
    int4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(reinterpret_cast<const int4*>(src + (size_t)(s + j) * split_stride));
#pragma unroll
    for (int j = 0; j < 8; ++j) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
  }
  for (; s < splits; ++s) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(src + (size_t)s * split_stride));
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  const float zpf = (float)ep.zp_out;
  const float rcp = __frcp_rn(ep.sc);
  const uint32_t zlo = ep.relu ? (uint32_t)ep.zp_out : 0u;
  const int32_t acc[4] = {a.x, a.y, a.z, a.w};
  uint32_t word = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n4 + j;
    uint32_t q = (uint32_t)ep.zp_out;   // pad lanes carry the zero point
    if (n < N) {
      int32_t v = acc[j] + __ldg(ep.oc + n);
      if (ep.bias_f) v = fc_bias_add(v, __ldg(ep.bias_f + n));
      if (ep.acc_out) ep.acc_out[(size_t)m * N + n] = v;
      const float sbn = ep.sb_vec ? __ldg(ep.sb_vec + n) : ep.sb;
      const uint32_t r = fast ? requant_u8_fast(v, ep.sa, sbn, ep.sc, rcp, zpf) : requant_u8(v, ep.sa, sbn, ep.sc, zpf);
      q = max(r, zlo);
    }
    word |= q << (8 * j);
  }
  *reinterpret_cast<uint32_t*>(y + (size_t)m * ldy + n4) = word;
}

// process-wide split-K scratch, grown outside of stream capture (the first eager call sizes it)
std::mutex g_ws_mu;
int32_t* g_ws = nullptr;
size_t g_ws_bytes = 0;

int ensure_workspace(size_t bytes, cudaStream_t stream, int32_t** out) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  if (bytes > g_ws_bytes) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &st);
    I8IE_REQUIRE(st == cudaStreamCaptureStatusNone,
                 "fc split-K workspace must grow (%zu bytes) while the stream is capturing: run the shape once eagerly first", bytes);
    I8IE_CUDA_OK(cudaDeviceSynchronize());
    if (g_ws) cudaFree(g_ws);
    g_ws = nullptr; g_ws_bytes = 0;
    const size_t want = ((bytes + (bytes >> 2)) + 255) & ~size_t(255);
    I8IE_CUDA_OK(cudaMalloc(&g_ws, want));
    g_ws_bytes = want;
  }
  *out = g_ws;
  return I8IE_OK;
}

// ---- host side: tensor maps -------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
using EncodeIm2colFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename Fn>
Fn driver_fn(const char* name) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<Fn>(f);
}

CUtensorMapSwizzle swizzle_for(int bk) {
  return bk == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : bk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

int encode_tiled_2d(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch, int box_cols,
                    int box_rows) {
  static EncodeTiledFn fn = driver_fn<EncodeTiledFn>("cuTensorMapEncodeTiled");
  I8IE_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(box_cols), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  I8IE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): cols=%llu rows=%llu pitch=%llu box=%dx%d",
               (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch, box_cols, box_rows);
  return I8IE_OK;
}

int encode_im2col_4d(CUtensorMap* tm, const void* base, const GemmGeom& g, int bk) {
  static EncodeIm2colFn fn = driver_fn<EncodeIm2colFn>("cuTensorMapEncodeIm2col");
  I8IE_REQUIRE(fn != nullptr, "cuTensorMapEncodeIm2col entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)g.cp, (cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)g.n};
  cuuint64_t strides[3] = {(cuuint64_t)g.cp, (cuuint64_t)g.cp * g.w, (cuuint64_t)g.cp * g.w * g.h};
  // bounding box of the filter's top-left tap, (W, H) order: [-pad, dim - 1 + pad - (k - 1)]
  int lower[2] = {-g.pad, -g.pad};
  int upper[2] = {g.pad - (g.kw - 1), g.pad - (g.kh - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims, strides, lower, upper,
                  (cuuint32_t)bk, (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bk),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  I8IE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (%d): c=%d w=%d h=%d n=%d k=%dx%d s=%d p=%d bk=%d",
               (int)r, g.cp, g.w, g.h, g.n, g.kh, g.kw, g.stride, g.pad, bk);
  return I8IE_OK;
}

template <int BN, int BK, int MODE, int MT>
int launch_kernel(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, int smem, cudaStream_t stream) {
  static int attr_smem = 0;
  auto kern = tc_igemm_kernel<BN, BK, MODE, MT>;
  if (attr_smem < smem) {
    I8IE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  const int tiles = p.tiles_m * p.tiles_n * p.splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  PdlFamily fam_(kPdlTc);
  launch_pdl(kern, dim3(grid), dim3(kThreads), (size_t)smem, stream, tmA, tmB, p);
  return check_launch("tc_igemm_kernel");
}

template <int BN, int BK, int MODE>
int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, TcParams p, cudaStream_t stream) {
  constexpr int kMaxSmem = 227 * 1024;
  constexpr bool kHasMT2 = false;   // 256-row single-CTA tiles: superseded by the CTA-pair kernel, not instantiated
  const int ctl_bytes = (int)sizeof(TcControl<BN>);
  if (p.mt < 1 || !kHasMT2) p.mt = 1;
  if (p.mt == 2) {
    // two 128-row sub-tiles only pay off when the wave count stays healthy and 3+ stages fit
    const int tiles2 = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.out_cp + BN - 1) / BN);
    const bool fits = (kMaxSmem - 1024 - ctl_bytes) / stage_bytes<BN, BK>(p.ksub, 2) >= 3;
    // measured on B200: the lost epilogue/main-loop overlap (both accumulators busy) costs more
    // than the saved weight traffic except in a few wave-quantisation cases -> opt-in only
    if (!fits || tiles2 * 100 < num_sms() * 85 || std::getenv("I8IE_TC_MT2") == nullptr) p.mt = 1;
  }
  const int kStage = stage_bytes<BN, BK>(p.ksub, p.mt);
  int stages = (kMaxSmem - 1024 - ctl_bytes) / kStage;
  if (stages > kMaxStages) stages = kMaxStages;
  if (const char* e = std::getenv("I8IE_TC_STAGES")) {
    const int v = std::atoi(e);
    if (v >= 1 && v < stages) stages = v;
  }
  I8IE_REQUIRE(stages >= 2, "tcgen05: stage of %d bytes leaves no room for a pipeline", kStage);
  p.stages = stages;
  const int smem = stages * kStage + ctl_bytes + 1024;
  p.tiles_m = (p.M + BM * p.mt - 1) / (BM * p.mt);
  p.tiles_n = (p.out_cp + BN - 1) / BN;
  p.tiles_mp = p.tiles_m;
  if (p.splits < 1) { p.splits = 1; p.kb_per = p.num_kb; }
  if constexpr (kHasMT2) {
    if (p.mt == 2) return launch_kernel<BN, BK, MODE, 2>(tmA, tmB, p, smem, stream);
  }
  return launch_kernel<BN, BK, MODE, 1>(tmA, tmB, p, smem, stream);
}

template <int BK, int MODE>
int launch_bn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t stream) {
  switch (bn) {
    case 32:  return launch_cfg<32, BK, MODE>(tmA, tmB, p, stream);
    case 64:  return launch_cfg<64, BK, MODE>(tmA, tmB, p, stream);
    case 96:  return launch_cfg<96, BK, MODE>(tmA, tmB, p, stream);
    case 128: return launch_cfg<128, BK, MODE>(tmA, tmB, p, stream);
    case 192: return launch_cfg<192, BK, MODE>(tmA, tmB, p, stream);
    case 256: return launch_cfg<256, BK, MODE>(tmA, tmB, p, stream);
  }
  set_error("tcgen05: unsupported BN %d", bn);
  return I8IE_EINVAL;
}

template <int MODE>
int launch_bk(int bk, int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t stream) {
  if (bk == 128) return launch_bn<128, MODE>(bn, tmA, tmB, p, stream);
  if constexpr (MODE == 1) {  // the row-tiled (fc) operand always uses 128-byte K blocks
    if (bk == 64) return launch_bn<64, MODE>(bn, tmA, tmB, p, stream);
    if (bk == 32) return launch_bn<32, MODE>(bn, tmA, tmB, p, stream);
  }
  set_error("tcgen05: unsupported BK %d", bk);
  return I8IE_EINVAL;
}

// CTA-pair launch (see tc_igemm2_kernel): BK = 128, one K block per stage, MT 256-row accumulators per tile
template <int BN, int MT = 1>
int launch_pair_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, TcParams p, cudaStream_t stream) {
  constexpr int kMaxSmem = 227 * 1024;
  const int ctl_bytes = (int)sizeof(TcControl<BN>);
  const int kStage = MT * BM * 128 + (BN / 2) * 128;
  int stages = (kMaxSmem - 1024 - ctl_bytes) / kStage;
  if (stages > kMaxStages) stages = kMaxStages;
  I8IE_REQUIRE(stages >= 3, "tcgen05 pair: no room for a pipeline");
  p.stages = stages;
  const int smem = stages * kStage + ctl_bytes + 1024;
  p.mt = MT; p.splits = 1; p.kb_per = p.num_kb;
  p.tiles_m = (p.M + BM - 1) / BM;
  p.tiles_n = (p.out_cp + BN - 1) / BN;
  p.tiles_mp = (p.tiles_m + 2 * MT - 1) / (2 * MT);
  static int attr_smem = 0;
  auto kern = tc_igemm2_kernel<BN, MT>;
  if (attr_smem < smem) {
    I8IE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  const int tiles = p.tiles_mp * p.tiles_n;
  const int max_clusters = num_sms() / 2;
  const int grid = (tiles < max_clusters ? tiles : max_clusters) * 2;
  PdlFamily fam_(kPdlPairConv);
  launch_cluster_pdl(2, kern, dim3(grid), dim3(kThreads2), (size_t)smem, stream, tmA, tmB, p);
  return check_launch("tc_igemm2_kernel");
}

// sub-blocks per pipeline stage: short K blocks are batched so that one mbarrier round trip
// covers >= 96..128 bytes of K (must divide the per-tap block count)
int pick_ksub(int bk, int cblocks) {
  if (bk == 128) return 1;
  const int want = bk == 64 ? 2 : 4;
  for (int k = want; k > 1; --k)
    if (cblocks % k == 0) return k;
  return 1;
}

int pick_bn(int n) {
  if (const char* e = std::getenv("I8IE_TC_BN")) {
    const int v = std::atoi(e);
    if (v == 32 || v == 64 || v == 96 || v == 128 || v == 192 || v == 256) return v;
  }
  if (n <= 32) return 32;
  if (n <= 64) return 64;
  if (n <= 96) return 96;
  if (n <= 128) return 128;
  if (n <= 192) return 192;
  if (n <= 256) return 256;
  // fewest padded columns; ties -> larger tile
  int best = 256, best_waste = 1 << 30;
  for (int bn : {256, 192, 128}) {
    const int waste = (n + bn - 1) / bn * bn - n;
    if (waste < best_waste) { best = bn; best_waste = waste; }
  }
  return best;
}

}  // namespace

// ---- entry points used by gemm_api.cu --------------------------------------------------------

bool tc_conv_eligible(const GemmGeom& g) {
  // K blocks are whole 32-byte multiples of one tap's channel run; corner offsets fit TMA's s8 range
  return g.cp % 32 == 0 && g.pad <= 3 && g.stride <= 8 && g.kh <= 128 && g.kw <= 128 && g.out_cp % 16 == 0;
}

int tc_conv_bk(const GemmGeom& g) { return g.cp % 128 == 0 ? 128 : g.cp % 64 == 0 ? 64 : 32; }

int tc_border_table_size(const GemmGeom& g) {
  if (g.pad == 0) return 0;
  const int d = g.pad + 1;
  return d * d * d * d * ((g.N + 31) & ~31);
}

int tc_build_border_table(const GemmGeom& g, const int8_t* w_packed, int32_t* tab, cudaStream_t stream) {
  const int total = tc_border_table_size(g);
  if (total == 0) return I8IE_OK;
  border_table_kernel<<<(total + 127) / 128, 128, 0, stream>>>(w_packed, tab, g.N, g.kh, g.kw, g.cp, g.pad);
  return check_launch("border_table_kernel");
}

int tc_encode_weight_map(CUtensorMap* tm, const int8_t* w, int rows, int ldw, int bk, int bn) {
  return encode_tiled_2d(tm, w, (uint64_t)ldw, (uint64_t)rows, (uint64_t)ldw, bk, bn);
}

int tc_encode_act_map_im2col(CUtensorMap* tm, const uint8_t* x, const GemmGeom& g, int bk) {
  return encode_im2col_4d(tm, x, g, bk);
}

// ---- row mode host side (see row_weight_kernel) ----
int tc_row_mode_kr(int cpx, int kw) { return (kw * cpx + 127) / 128 * 128; }

// Hard constraints of row mode: stride 1, a run that fits a few K blocks.
bool tc_row_mode_ok(int c, int kh, int kw, int stride, int pad, int out_cp) {
  if (stride != 1 || kw < 2 || pad > 8 || out_cp % 16 != 0 || kh > 16) return false;
  return tc_row_mode_kr((c + 15) / 16 * 16, kw) <= 2048;
}

// Channel pitch (bytes per pixel) the physically padded input must have, or 0 when the layer should
// stay on the plain im2col path (auto dispatch). Row mode is taken when it saves at least 10 % of the K
// bytes (C = 96: 5 x 512 instead of 25 x 128), and for narrow-channel layers whose plain path would run
// K blocks of 16 / 32 / 64 bytes (or the SIMT kernel): there the run of kw pixels is one 128-byte-wide K
// block per filter row instead of kw small ones, as long as that costs at most 2x the K bytes.
int tc_row_mode_cp(int c, int cp_plain, int kh, int kw, int stride, int pad, int out_cp) {
  if (std::getenv("I8IE_NO_ROW_MODE") != nullptr || !tc_row_mode_ok(c, kh, kw, stride, pad, out_cp)) return 0;
  const int cpx = (c + 15) / 16 * 16;
  const long long k_plain = (long long)kh * kw * cp_plain, k_row = (long long)kh * tc_row_mode_kr(cpx, kw);
  if (k_row * 10 <= k_plain * 9) return cpx;
  if (cp_plain % 128 != 0 && k_row <= 2 * k_plain) return cpx;
  return 0;
}

int tc_row_pack_weights(const GemmGeom& g, const int8_t* w_packed, int8_t* wr, int kr, cudaStream_t stream) {
  const int64_t total = (int64_t)g.n_pad * g.kh * kr;
  int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16);
  row_weight_kernel<<<blocks, 256, 0, stream>>>(w_packed, wr, g.n_pad, g.kh, g.kw, g.cp, kr);
  return check_launch("row_weight_kernel");
}

// x = padded input [n][h + 2p][w + 2p][cp] (g is the layer's REAL geometry with cp = the padded tensor's pitch)
int tc_encode_act_map_row_mode(CUtensorMap* tm, const uint8_t* x, const GemmGeom& g, int kr) {
  static EncodeIm2colFn fn = driver_fn<EncodeIm2colFn>("cuTensorMapEncodeIm2col");
  I8IE_REQUIRE(fn != nullptr, "cuTensorMapEncodeIm2col entry point not available");
  const int hp = g.h + 2 * g.pad, wp = g.w + 2 * g.pad;
  cuuint64_t dims[4] = {(cuuint64_t)kr, (cuuint64_t)g.ow, (cuuint64_t)hp, (cuuint64_t)g.n};
  cuuint64_t strides[3] = {(cuuint64_t)g.cp, (cuuint64_t)wp * g.cp, (cuuint64_t)hp * wp * g.cp};
  int lower[2] = {0, 0};
  int upper[2] = {0, -(g.kh - 1)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t*>(x), dims, strides, lower, upper, 128,
                  (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  I8IE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col (row-mode view) failed (%d): kr=%d ow=%d hp=%d n=%d cp=%d", (int)r,
               kr, g.ow, hp, g.n, g.cp);
  return I8IE_OK;
}

// The virtual geometry the pair kernel runs a row-mode layer with: kh x 1 filter over KR "channels".
GemmGeom tc_row_mode_geom(const GemmGeom& g, int kr) {
  GemmGeom v = g;
  v.h = g.h + 2 * g.pad; v.w = g.ow; v.cp = kr;
  v.kw = 1; v.pad = 0; v.ldw = g.kh * kr;
  return v;
}

int tc_encode_act_map_rows(CUtensorMap* tm, const uint8_t* x, int m, int k, int ldx) {
  return encode_tiled_2d(tm, x, (uint64_t)k, (uint64_t)m, (uint64_t)ldx, 128, BM);
}

int tc_pick_bn(int n) { return pick_bn(n); }

// Cluster mode of a conv plan: 1 = single CTAs; 2 = CTA pairs, tcgen05.mma.cta_group::2 with the
// weight tile split across the pair (wide tiles with 128-byte K blocks; the weight tensor map's box
// then holds bn / 2 rows = whole 1 KB swizzle atoms). I8IE_NO_CLUSTER=1 forces single CTAs (tests).
int tc_conv_cluster(int bk, int bn) {
  if (std::getenv("I8IE_NO_CLUSTER") != nullptr) return 1;
  const bool ok = bk == 128 && (bn == 256 || bn == 192 || bn == 128);
  return ok ? 2 : 1;
}

// Rows of the weight tensor map's box for the pair kernel: each CTA loads its half of the rows of every
// MMA instruction (N <= 256 per instruction; BN = 384 is two instructions of N = 192).
int tc_pair_box_rows(int bn) { return bn > 256 ? bn / 4 : bn / 2; }

// N tile (and accumulators per tile) of a conv plan served by the pair kernel: the narrowest tile that
// covers the output pitch with the fewest padded columns (pick_bn), one 256-row accumulator per tile.
// Measured on B200 with 16 epilogue warps (profiles/r02_pair_tile_ab.md, AlexNet conv2-5, batch 100-1000):
// this beats both wide variants the kernel template can express — a 384-column tile (two N = 192 MMAs per A
// stage) and 512 x 256 tiles (two accumulators per weight stage) cut the L2 -> SM bytes per MAC by 19-25 %,
// but fill TMEM with ONE tile, and the lost epilogue / main-loop overlap costs more (conv4 batch 250:
// 59.5 us vs 39.1 us) — so they are not instantiated. I8IE_TC_BN overrides the width (tests).
int tc_pick_bn_pair(const GemmGeom& g, int bk, int* mt_out) {
  (void)bk;
  *mt_out = 1;
  return pick_bn(g.out_cp);
}

int launch_tc_conv(const GemmGeom& g, const CUtensorMap& tmA, const CUtensorMap& tmB, int bk, int bn, int cluster,
                   const int32_t* border_tab, uint8_t* y, const EpiParams& ep, int zp_in, cudaStream_t stream, int pair_mt) {
  TcParams p{};
  p.M = g.M; p.N = g.N; p.out_cp = g.out_cp;
  p.cblocks = g.cp / bk;
  p.ksub = pick_ksub(bk, p.cblocks);
  p.num_kb = g.kh * g.kw * p.cblocks / p.ksub;
  p.kh = g.kh; p.kw = g.kw; p.stride_h = p.stride_w = g.stride; p.pad = g.pad; p.H = g.h; p.W = g.w; p.oh = g.oh; p.ow = g.ow;
  p.zp_in = zp_in; p.border_tab = border_tab; p.y = y; p.ep = ep;
  p.fast_requant = requant_fast_ok(ep);
  if (cluster == 2) {   // CTA pairs
    I8IE_REQUIRE(bk == 128, "tcgen05 pair: needs 128-byte K blocks");
    (void)pair_mt;
    switch (bn) {
      case 128: return launch_pair_bn<128>(tmA, tmB, p, stream);
      case 192: return launch_pair_bn<192>(tmA, tmB, p, stream);
      case 256: return launch_pair_bn<256>(tmA, tmB, p, stream);
    }
    set_error("tcgen05 pair: unsupported BN %d", bn);
    return I8IE_EINVAL;
  }
  p.mt = 2;   // launch_cfg falls back to 1 when the shape is too small for it
  return launch_bk<1>(bk, bn, tmA, tmB, p, stream);
}

// Tile width and K split for an fc shape. Large M: one wide tile per CTA. Small M (a weight
// stream): every SM should pull an equal, minimal share of bytes through its ~43 B/clk L2
// port, so pick the (BN, splits) pair with the fewest bytes per CTA, counting the activation
// rows each N tile re-reads and the s32 partials it writes and the reduce kernel re-reads.
bool fc_cluster_ok(int bn, int num_kb, int ldy);

void tc_fc_config(int m, int ldy, int k, int* bn_out, int* splits_out, int* kb_per_out, int* cluster_out) {
  const int num_kb = (k + 127) / 128;
  const int tiles_m = (m + BM - 1) / BM;
  int bn = pick_bn(ldy);
  int splits = 1, kb_per = num_kb, cluster = 0;
  const bool allow = std::getenv("I8IE_NO_SPLITK") == nullptr && std::getenv("I8IE_TC_BN") == nullptr;
  if (allow && tiles_m * ((ldy + bn - 1) / bn) * 2 <= num_sms() && num_kb >= 2) {
    long long best = -1;
    const int rows = m < BM ? m : BM;
    for (int cand : {256, 192, 128, 96, 64, 32}) {
      if (cand > bn && cand > ((ldy + 31) / 32) * 32) continue;
      const int tiles = tiles_m * ((ldy + cand - 1) / cand);
      if (tiles > num_sms()) continue;
      int s = num_sms() / tiles;
      if (s > num_kb) s = num_kb;
      if (s > 32) s = 32;
      const int per = (num_kb + s - 1) / s;
      s = (num_kb + per - 1) / per;
      const long long cost = (long long)per * 128 * (rows + cand) + (s > 1 ? 8ll * rows * cand : 0);
      if (best < 0 || cost < best) { best = cost; bn = cand; splits = s; kb_per = per; }
    }
  }
  // cluster split-K (tc_fc_cluster_kernel: 128-wide N tiles, K over the 4 CTAs of a cluster, partials folded through
  // distributed shared memory — no scratch round trip, no second kernel) whenever that shape fills most of the chip
  // Measured (fc1 / fc2 of AlexNet, us): batch 16 11.4 / 12.0 vs 10.1 / 9.8 for independent CTAs + reduce kernel (the
  // partials are tiny there), batch 64 11.6 / 9.8 vs 12.3 / 9.3, batch 100 11.7 / 9.9 vs 14.7 / 10.7, batch 128
  // 11.7 / 9.9 vs 17.2 / 12.0; with two M tiles (batch 250) every cluster wave streams the weights again:
  // 23.0 / 19.5 vs 20.8 / 17.9 — so: exactly one M tile of at least 64 rows.
  if (splits > 1 && tiles_m == 1 && m >= 64 && fc_cluster_ok(128, num_kb, ldy)) {
    const int ctas = ((ldy + 127) / 128) * 4;
    if (ctas * 3 >= num_sms() * 2 && ctas <= num_sms()) { bn = 128; splits = 4; kb_per = (num_kb + 3) / 4; cluster = 1; }
  }
  *cluster_out = cluster;
  *bn_out = bn; *splits_out = splits; *kb_per_out = kb_per;
}

int tc_fc_tile_weights(const int8_t* w, int n_pad, int ldw, int8_t* wt, cudaStream_t stream) {
  I8IE_REQUIRE(w && wt && n_pad > 0 && n_pad % 128 == 0 && ldw > 0 && ldw % 128 == 0 &&
                   (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(wt) & 15) == 0,
               "fc weight tiling needs n_pad and ldw to be multiples of 128 and 16-byte aligned buffers (n_pad=%d ldw=%d)", n_pad, ldw);
  const int64_t total = (int64_t)n_pad * (ldw >> 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  fc_tile_weight_kernel<<<blocks, 256, 0, stream>>>(w, wt, n_pad, ldw);
  return check_launch("fc_tile_weight_kernel");
}

// Cluster split-K fc (tc_fc_cluster_kernel): S = bn / 32 CTAs per (M tile, N tile), S <= 8 (portable cluster size).
// Eligible when every rank gets at least one K block and the output pitch is a multiple of 8.
template <int BN>
int launch_fc_cluster_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, TcParams p, cudaStream_t stream) {
  constexpr int S = BN / 32;
  constexpr int kStage = BM * 128 + BN * 128;
  const int ctl_bytes = (int)sizeof(TcControl<BN>);
  int stages = (227 * 1024 - 1024 - ctl_bytes) / kStage;
  if (stages > kMaxStages) stages = kMaxStages;
  I8IE_REQUIRE(stages * kStage >= S * 128 * 128 + kEpiWarps * 4096, "cluster fc: operand ring smaller than receive buffer + staging");
  p.stages = stages;
  p.tiles_m = (p.M + BM - 1) / BM;
  p.tiles_n = (p.out_cp + BN - 1) / BN;
  p.splits = S;
  p.kb_per = (p.num_kb + S - 1) / S;
  const int smem = stages * kStage + ctl_bytes + 1024;
  static int attr_smem = 0;
  auto kern = tc_fc_cluster_kernel<BN>;
  if (attr_smem < smem) {
    I8IE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  if (const char* e = std::getenv("I8IE_FC_DBG")) {
    p.dbg = std::atoi(e);
    static bool once = false;
    if (!once) {
      once = true;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(p.tiles_m * p.tiles_n * S)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = (size_t)smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nc = -1;
      cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
      std::fprintf(stderr, "[i8ie] cluster fc BN=%d S=%d smem=%d stages=%d grid=%d: max active clusters %d\n", BN, S, smem, stages,
                   p.tiles_m * p.tiles_n * S, nc);
    }
  }
  PdlFamily fam_(kPdlFcCluster);
  launch_cluster_pdl(S, kern, dim3((unsigned)(p.tiles_m * p.tiles_n * S)), dim3(kThreads), (size_t)smem, stream, tmA, tmB, p);
  return check_launch("tc_fc_cluster_kernel");
}

bool fc_cluster_ok(int bn, int num_kb, int ldy) {
  static const bool off = std::getenv("I8IE_NO_FC_CLUSTER") != nullptr;
  if (off || bn != 128 || ldy % 8 != 0) return false;
  const int S = bn / 32, per = (num_kb + S - 1) / S;
  return per * (S - 1) < num_kb;   // the last rank still has a K block
}

int launch_tc_fc(int m, int n, int k, int ldy, const CUtensorMap& tmA, const CUtensorMap& tmB, int bn, int splits,
                 int kb_per, uint8_t* y, const EpiParams& ep, cudaStream_t stream, const int8_t* w_tiled, int ldw,
                 int cluster) {
  TcParams p{};
  p.M = m; p.N = n; p.out_cp = ldy;
  if (w_tiled != nullptr && bn % 128 == 0 && ldw % 128 == 0) { p.w_tiled = w_tiled; p.wt_nkb = ldw / 128; }
  p.cblocks = 1; p.ksub = 1; p.num_kb = (k + 127) / 128;
  p.kh = p.kw = 1; p.stride_h = p.stride_w = 1; p.pad = 0; p.H = p.W = p.oh = p.ow = 1;
  p.zp_in = 0; p.border_tab = nullptr; p.y = y; p.ep = ep;
  p.fast_requant = requant_fast_ok(ep);
  if (splits <= 1) return launch_bk<0>(128, bn, tmA, tmB, p, stream);
  // split K across the CTAs of a cluster and fold the partial sums through distributed shared memory ...
  if (cluster && fc_cluster_ok(bn, p.num_kb, ldy) && splits == 4) return launch_fc_cluster_bn<128>(tmA, tmB, p, stream);
  // ... or across independent CTAs, folding through a global scratch buffer (exact either way: integer adds commute)
  p.splits = splits; p.kb_per = kb_per;
  p.ws_ld = ((ldy + bn - 1) / bn) * bn;
  int rc = ensure_workspace(sizeof(int32_t) * (size_t)p.splits * m * p.ws_ld, stream, &p.ws);
  if (rc != I8IE_OK) return rc;
  rc = launch_bk<0>(128, bn, tmA, tmB, p, stream);
  if (rc != I8IE_OK) return rc;
  const long long threads = (long long)m * (ldy / 4);
  launch_pdl(fc_splitk_reduce_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, stream, p.ws, p.splits, m,
             n, p.ws_ld, ldy, y, ep, p.fast_requant);
  return check_launch("fc_splitk_reduce_kernel");
}

// ---- stem path host side ---------------------------------------------------------------------
bool tc_stem_eligible(const GemmGeom& g, int c) {
  return c <= 4 && (g.stride == 4 || g.stride == 8) && g.kw <= 16 && g.kh <= 128 && g.pad <= g.kw &&
         g.out_cp % 16 == 0;
}

StemGeom tc_stem_geom(const GemmGeom& g, int c) {
  StemGeom s;
  s.c = c;
  s.hp = (g.oh - 1) * g.stride + g.kh;
  s.wsp = (g.ow - 1) * (g.stride / 4) + 4;
  s.bytes = (int64_t)g.n * s.hp * s.wsp * 16;
  return s;
}

int tc_stem_pack_weights(const GemmGeom& g, int c, const int8_t* w_packed, int8_t* ws, cudaStream_t stream) {
  const int total = g.n_pad * g.kh * 64;
  stem_weight_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w_packed, ws, g.n_pad, c, g.kh, g.kw, g.cp);
  return check_launch("stem_weight_kernel");
}

// K packing of the stem2 kernel (see Stem2Params::k48_n): bytes of one packed weight row, 0 = not applicable
int tc_stem_k48_bytes(const GemmGeom& g, int c) {
  if (!tc_stem2_eligible(g, c) || g.kw > 12 || std::getenv("I8IE_NO_STEM_K48") != nullptr) return 0;
  const int n32 = (g.kh * 48 + 31) / 32;
  return n32 <= 24 ? n32 * 32 : 0;
}

int tc_stem_pack_weights48(const GemmGeom& g, int c, const int8_t* w_packed, int8_t* ws, cudaStream_t stream) {
  const int k48 = tc_stem_k48_bytes(g, c);
  I8IE_REQUIRE(k48 > 0, "stem K packing does not apply to this geometry");
  const int total = g.n_pad * k48;
  stem_weight48_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w_packed, ws, g.n_pad, c, g.kh, g.kw, g.cp, k48);
  return check_launch("stem_weight48_kernel");
}

int tc_stem_pack_input(const GemmGeom& g, const StemGeom& s, const uint8_t* x, uint8_t* xs, int zp,
                       cudaStream_t stream) {
  const int64_t total = (int64_t)g.n * s.hp * s.wsp;
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  stem_pack_kernel<<<blocks, 256, 0, stream>>>(x, xs, g.n, s.c, g.h, g.w, g.cp, g.pad, s.hp, s.wsp, (uint32_t)zp);
  return check_launch("stem_pack_kernel");
}

int tc_stem_quantize_input(const GemmGeom& g, const StemGeom& s, const float* x, const float* const* xslot,
                           uint8_t* xs, float scale, int zp, cudaStream_t stream) {
  const dim3 grid((unsigned)((s.hp * s.wsp + 255) / 256), (unsigned)g.n);
  const bool vec2 = (g.pad % 2 == 0) && (g.w % 2 == 0) && (xslot || (reinterpret_cast<uintptr_t>(x) & 7) == 0);
  const float lim = quant_fast_limit(scale);
  const bool fast = lim > 0.f;
  I8IE_REQUIRE(g.n <= 65535, "stem quantise: batch %d exceeds the grid.y limit", g.n);
#define I8IE_STEMQ(V, F) launch_pdl(stem_quantize_kernel<V, F>, grid, dim3(256), 0, stream, \
      x, xs, s.c, g.h, g.w, g.pad, s.hp, s.wsp, scale, (float)zp, (uint32_t)zp, lim, xslot)
  if (vec2 && fast) I8IE_STEMQ(true, true);
  else if (vec2) I8IE_STEMQ(true, false);
  else if (fast) I8IE_STEMQ(false, true);
  else I8IE_STEMQ(false, false);
#undef I8IE_STEMQ
  return check_launch("stem_quantize_kernel");
}

int tc_encode_stem_act_map(CUtensorMap* tm, const uint8_t* xs, const GemmGeom& g, const StemGeom& s) {
  static EncodeIm2colFn fn = driver_fn<EncodeIm2colFn>("cuTensorMapEncodeIm2col");
  I8IE_REQUIRE(fn != nullptr, "cuTensorMapEncodeIm2col entry point not available");
  const int sw = g.stride / 4;
  // overlapping view: "pixel" = 64-byte run starting at superpixel w; consecutive pixels are 16 bytes apart
  cuuint64_t dims[4] = {64, (cuuint64_t)((g.ow - 1) * sw + 1), (cuuint64_t)s.hp, (cuuint64_t)g.n};
  cuuint64_t strides[3] = {16, (cuuint64_t)s.wsp * 16, (cuuint64_t)s.hp * s.wsp * 16};
  int lower[2] = {0, 0};
  int upper[2] = {0, -(g.kh - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)sw, (cuuint32_t)g.stride, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t*>(xs), dims, strides, lower, upper, 64,
                  (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  I8IE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col (stem view) failed (%d)", (int)r);
  return I8IE_OK;
}

int launch_tc_stem(const GemmGeom& g, const CUtensorMap& tmA, const CUtensorMap& tmB, int bn, uint8_t* y,
                   const EpiParams& ep, cudaStream_t stream) {
  TcParams p{};
  p.M = g.M; p.N = g.N; p.out_cp = g.out_cp;
  p.cblocks = 1; p.ksub = 1; p.num_kb = g.kh;
  p.kh = g.kh; p.kw = 1; p.stride_h = g.stride; p.stride_w = g.stride / 4; p.pad = 0;
  p.H = (g.oh - 1) * g.stride + g.kh; p.W = (g.ow - 1) * (g.stride / 4) + 1; p.oh = g.oh; p.ow = g.ow;
  p.zp_in = 0; p.border_tab = nullptr; p.y = y; p.ep = ep;
  p.fast_requant = requant_fast_ok(ep);
  return launch_bk<1>(64, bn, tmA, tmB, p, stream);
}

bool tc_stem2_eligible(const GemmGeom& g, int c) {
  // stride 4 (16-byte window step), both row windows inside one 1 KB slot, resident weights fit
  const int wsp = (g.ow - 1) + 4;
  return tc_stem_eligible(g, c) && g.stride == 4 && wsp <= 64 && g.kh <= 12 && g.out_cp <= 128 &&
         g.kh * tc_pick_bn(g.out_cp) * 64 <= 100 * 1024;
}

template <int BN>
int launch_stem2_bn(const GemmGeom& g, const StemGeom& s, const uint8_t* xs, const StemF32Src* f32,
                    const CUtensorMap& tmB, TcParams p, cudaStream_t stream, bool k48) {
  Stem2Params sp{};
  sp.n_img = g.n; sp.oh = g.oh; sp.ow = g.ow; sp.kh = g.kh; sp.hp = s.hp; sp.wsp = s.wsp;
  sp.pairs = (g.oh + 1) / 2;
  sp.nsl = (g.kh + 4 + 3) / 4;
  if (k48) {
    sp.k48_n = (g.kh * 48 + 31) / 32;
    auto rowaddr = [&](int i) { return ((i & 3) * sp.nsl + (i >> 2)) * 1024; };
    for (int j = 0; j < sp.k48_n; ++j) {
      const int b = 32 * j, r = b / 48, o = b % 48;
      int start = rowaddr(r < g.kh ? r : g.kh - 1) + (r < g.kh ? o : 0), lbo = 16;
      if (r < g.kh && o == 32 && r + 1 < g.kh) lbo = rowaddr(r + 1) - rowaddr(r) - 32;   // straddles two filter rows
      I8IE_REQUIRE(lbo > 0 && lbo % 16 == 0 && (lbo >> 4) < (1 << 14), "stem2 K packing: unsupported row distance %d", lbo);
      sp.k48_a[j] = (uint32_t)(start >> 4) | ((uint32_t)(lbo >> 4) << 16);
    }
  }
  sp.xs = xs;
  const bool fq = f32 != nullptr;
  if (fq) {
    sp.xf = f32->x; sp.xslot = f32->xslot; sp.c = s.c; sp.h = g.h; sp.w = g.w; sp.pad = g.pad;
    sp.in_scale = f32->scale; sp.in_zp = f32->zp; sp.fast_lim = quant_fast_limit(f32->scale);
  }
  sp.dbg = 0;
  if (const char* e = std::getenv("I8IE_STEM2_DBG")) sp.dbg = std::atoi(e);
  sp.pf_tiles = 8;
  if (const char* e = std::getenv("I8IE_STEM2_PF")) sp.pf_tiles = std::atoi(e);
  sp.trace = nullptr;
#ifdef I8IE_STEM_TRACE
  const char* trace_path = fq ? std::getenv("I8IE_STEM2_TRACE") : nullptr;
#else
  const char* trace_path = nullptr;
#endif
  static long long* d_trace = nullptr;
  if (trace_path != nullptr) {   // dev only: eager launches, synchronises
    if (d_trace == nullptr) I8IE_CUDA_OK(cudaMalloc(&d_trace, sizeof(long long) * kTraceTiles * kTraceEvents));
    I8IE_CUDA_OK(cudaMemset(d_trace, 0, sizeof(long long) * kTraceTiles * kTraceEvents));
    sp.trace = d_trace;
  }

  const int w_bytes = k48 ? sp.k48_n * BN * 32 : g.kh * BN * 64;
  const int a_stage = 4 * sp.nsl * 1024;
  const int ctl_bytes = (int)sizeof(TcControl<BN>);
  const int budget = 227 * 1024 - 1024 - ctl_bytes - w_bytes;
  int stages, smem;
  if (fq) {
    // two operand stages (a tile's MMAs are short), the rest of shared memory is the fp32 row ring
    stages = 2;
    sp.f_stage_bytes = (((g.kh + 4) * s.c * g.w * 4) + 127) & ~127;
    sp.f_stages = (budget - stages * a_stage) / sp.f_stage_bytes;
    if (sp.f_stages > kMaxStages - 1 - kStemFBar) sp.f_stages = kMaxStages - 1 - kStemFBar;
    I8IE_REQUIRE(sp.f_stages >= 2, "stem2: not enough shared memory for the fp32 row ring");
    smem = w_bytes + stages * a_stage + sp.f_stages * sp.f_stage_bytes + ctl_bytes + 1024;
  } else {
    stages = budget / a_stage;
    if (stages > 6) stages = 6;
    I8IE_REQUIRE(stages >= 2, "stem2: not enough shared memory for the row ring");
    smem = w_bytes + stages * a_stage + ctl_bytes + 1024;
  }
  sp.stages = stages;
  static int attr_smem = 0;
  auto kern = fq ? (g.kh == 11 ? tc_stem2_kernel<BN, 11, true> : tc_stem2_kernel<BN, 0, true>)
                 : (g.kh == 11 ? tc_stem2_kernel<BN, 11, false> : tc_stem2_kernel<BN, 0, false>);
  if (attr_smem < smem) {
    I8IE_CUDA_OK(cudaFuncSetAttribute(tc_stem2_kernel<BN, 11, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    I8IE_CUDA_OK(cudaFuncSetAttribute(tc_stem2_kernel<BN, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    I8IE_CUDA_OK(cudaFuncSetAttribute(tc_stem2_kernel<BN, 11, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    I8IE_CUDA_OK(cudaFuncSetAttribute(tc_stem2_kernel<BN, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  const int tiles = sp.n_img * sp.pairs;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  const int threads = fq ? stem_threads<BN, true>() : stem_threads<BN, false>();
  PdlFamily fam_(kPdlStem);
  launch_cluster_pdl(fq ? -1 : 1, kern, dim3(grid), dim3(threads), (size_t)smem, stream, tmB, p, sp);
  if (trace_path != nullptr) {
    std::vector<long long> h(kTraceTiles * kTraceEvents);
    I8IE_CUDA_OK(cudaDeviceSynchronize());
    I8IE_CUDA_OK(cudaMemcpy(h.data(), d_trace, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    if (FILE* f = std::fopen(trace_path, "w")) {
      long long t0 = 0;
      for (long long v : h) if (v != 0 && (t0 == 0 || v < t0)) t0 = v;
      std::fprintf(f, "# tile: ev0 prod f_empty | 1 conv f_full 2 conv empty 3 conv bar 4 conv copied 5 conv converted 6 conv arrived | "
                      "8 mma tmem_empty 9 mma full 10 mma committed | 12 epi tmem_full 13 epi done 14 epi arrived (clocks since first event)\n");
      for (int t = 0; t < kTraceTiles; ++t) {
        std::fprintf(f, "%2d:", t);
        for (int e = 0; e < kTraceEvents; ++e) std::fprintf(f, " %7lld", h[t * kTraceEvents + e] ? h[t * kTraceEvents + e] - t0 : -1ll);
        std::fprintf(f, "\n");
      }
      std::fclose(f);
    }
  }
  return check_launch("tc_stem2_kernel");
}

// The quantise can be fused into the stem kernel when its fast path applies: RGB, even pad (float2
// reads of the staged rows), image rows that are whole 16-byte units from a 16-byte aligned source
// (bulk copies; slot sources are allocator-aligned), two fp32 tile stages fit, mid-range scale.
bool tc_stem2_can_fuse_quantize(const GemmGeom& g, int c, const float* x, const float* const* xslot, float scale) {
  if (std::getenv("I8IE_NO_STEM_FUSEQ") != nullptr) return false;
  const int f_stage = (g.kh + 4) * c * g.w * 4;
  return c == 3 && (g.pad % 2 == 0) && (g.w % 4 == 0) && (xslot || (reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
         2 * f_stage <= 100 * 1024 && quant_fast_limit(scale) > 0.f;
}

int launch_tc_stem2(const GemmGeom& g, const StemGeom& s, const uint8_t* xs, const StemF32Src* f32,
                    const CUtensorMap& tmB, int bn, uint8_t* y, const EpiParams& ep, cudaStream_t stream, bool k48) {
  TcParams p{};
  p.M = g.M; p.N = g.N; p.out_cp = g.out_cp;
  p.zp_in = 0; p.border_tab = nullptr; p.y = y; p.ep = ep;
  p.fast_requant = requant_fast_ok(ep);
  I8IE_REQUIRE(bn % 32 == 0 || bn >= g.out_cp, "stem2: N tile %d narrower than the output pitch must be a multiple of 32", bn);
  switch (bn) {
    case 32:  return launch_stem2_bn<32>(g, s, xs, f32, tmB, p, stream, k48);
    case 64:  return launch_stem2_bn<64>(g, s, xs, f32, tmB, p, stream, k48);
    case 96:  return launch_stem2_bn<96>(g, s, xs, f32, tmB, p, stream, k48);
    case 128: return launch_stem2_bn<128>(g, s, xs, f32, tmB, p, stream, k48);
  }
  set_error("stem2: unsupported BN %d", bn);
  return I8IE_EINVAL;
}

// ---- protocol-error sink -------------------------------------------------------------------------
namespace {
constexpr int kMaxDevices = 64;
std::mutex g_sink_mu;
int* g_sink_host[kMaxDevices] = {};   // pinned, mapped; one int per device
int* g_sink_dev[kMaxDevices] = {};    // the same words as device addresses
}  // namespace

// Creates the current device's host-mapped error flag and publishes its device address to the
// kernels. Synchronising calls inside: must run outside of stream capture (plan creation and
// i8ie_device_check() call it; a model's eager warm-up calls always precede its graph capture).
int tc_error_sink_init() {
  int dev = 0;
  I8IE_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return I8IE_OK;
  std::lock_guard<std::mutex> lk(g_sink_mu);
  if (g_sink_host[dev] != nullptr) return I8IE_OK;
  int* h = nullptr;
  I8IE_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&h), sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
  *h = 0;
  int* d = nullptr;
  I8IE_CUDA_OK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0));
  I8IE_CUDA_OK(cudaMemcpyToSymbol(g_tc_error_host, &d, sizeof(d)));

  g_sink_dev[dev] = d;
  g_sink_host[dev] = h;
  return I8IE_OK;
}

// Device address of the current device's flag (nullptr before tc_error_sink_init): kernels of other
// translation units (the peer result exchange) report their bounded-wait timeouts through it.
int* tc_error_sink_device_ptr() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  return g_sink_dev[dev];
}

// Called on the launch paths: creates the sink on first use unless the stream is capturing.
void tc_error_sink_touch(cudaStream_t stream) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices || g_sink_host[dev] != nullptr) return;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return;
  tc_error_sink_init();
}

// First protocol error recorded on the current device since the last reset (0 = none). Plain host
// memory read: meaningful after the work in question has been synchronised.
int tc_error_poll(bool reset) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  int* h = g_sink_host[dev];
  if (h == nullptr) return 0;
  const int v = *reinterpret_cast<volatile int*>(h);
  if (reset && v != 0) *reinterpret_cast<volatile int*>(h) = 0;
  return v;
}

int tc_read_error(int* out, bool reset) {
  int v = 0;
  I8IE_CUDA_OK(cudaMemcpyFromSymbol(&v, g_tc_error, sizeof(int)));
  if (reset && v != 0) {
    const int z = 0;
    I8IE_CUDA_OK(cudaMemcpyToSymbol(g_tc_error, &z, sizeof(int)));
    tc_error_poll(true);
  }
  *out = v;
  return I8IE_OK;
}

}  // namespace i8ie
