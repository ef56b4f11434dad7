// SIMT (dp4a) implicit-GEMM kernel for the shapes the tcgen05 path does not take
// (tiny / oddly-shaped layers, forced impl=1) and the reference against which the
// tensor-core kernels are unit-tested on the device. Same fused epilogue as the
// tcgen05 kernels: + oc, [fc float bias], requantise, [relu], u8 NHWC store.
//
// conv (conv2d.cc:100-142): A[m][k] is gathered on the fly from the NHWC input
// (m = (img, oy, ox), k = (ky, kx, c)); spatially out-of-range taps are filled with
// the input zero_point exactly like im2col_tile (conv2d.cc:24-25). No im2col
// buffer is materialised. FC (fully_connected.cc:22-52) is the same kernel with a
// 1x1 "image" per row.
#include "common.cuh"
#include "gemm_api.cuh"

namespace i8ie {
namespace {

constexpr int BM = 64, BN = 64, NT = 256;

__global__ void __launch_bounds__(NT) simt_igemm_kernel(const GemmGeom g, const uint8_t* __restrict__ x,
                                                       const int8_t* __restrict__ w,
                                                       uint8_t* __restrict__ y, const EpiParams ep,
                                                       const uint32_t zp_in4) {
  __shared__ uint4 sA[2][BM];
  __shared__ uint4 sB[2][BN];
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // loader roles: threads [0,64) fetch one A row chunk, [64,128) one B row chunk
  const bool isA = tid < BM, isB = (tid >= BM) && (tid < BM + BN);
  const int r = tid & 63;
  const int cgroups = g.cp >> 4;
  const int ksteps = g.kh * g.kw * cgroups;

  const uint8_t* abase = nullptr;
  int iy0 = 0, ix0 = 0;
  bool mvalid = false;
  const int8_t* wrow = nullptr;
  if (isA) {
    const int m = m0 + r;
    mvalid = m < g.M;
    if (mvalid) {
      const int ox = m % g.ow;
      const int t = m / g.ow;
      const int oy = t % g.oh;
      const int img = t / g.oh;
      iy0 = oy * g.stride - g.pad;
      ix0 = ox * g.stride - g.pad;
      abase = x + (size_t)img * g.h * g.w * g.cp;
    }
  } else if (isB) {
    const int n = n0 + r;
    if (n < g.n_pad) wrow = w + (size_t)n * g.ldw;
  }

  int ky = 0, kx = 0, cg = 0;  // loader's running (tap, channel-group) position
  auto fetch = [&]() -> uint4 {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (isA) {
      v = make_uint4(zp_in4, zp_in4, zp_in4, zp_in4);
      const int iy = iy0 + ky, ix = ix0 + kx;
      if (mvalid && iy >= 0 && iy < g.h && ix >= 0 && ix < g.w)
        v = __ldg(reinterpret_cast<const uint4*>(abase + ((size_t)iy * g.w + ix) * g.cp + cg * 16));
    } else if (isB) {
      if (wrow) v = __ldg(reinterpret_cast<const uint4*>(wrow + (size_t)(ky * g.kw + kx) * g.cp + cg * 16));
    }
    if (++cg == cgroups) { cg = 0; if (++kx == g.kw) { kx = 0; ++ky; } }
    return v;
  };

  int32_t acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;

  uint4 pre = fetch();
  if (isA) sA[0][r] = pre; else if (isB) sB[0][r] = pre;
  __syncthreads();
  for (int ks = 0; ks < ksteps; ++ks) {
    const int cur = ks & 1;
    if (ks + 1 < ksteps) pre = fetch();
    uint4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = sA[cur][ty + 16 * i];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = sB[cur][tx + 16 * j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int32_t s = acc[i][j];
        s = dp4a_u8s8(a[i].x, b[j].x, s);
        s = dp4a_u8s8(a[i].y, b[j].y, s);
        s = dp4a_u8s8(a[i].z, b[j].z, s);
        s = dp4a_u8s8(a[i].w, b[j].w, s);
        acc[i][j] = s;
      }
    if (ks + 1 < ksteps) {
      if (isA) sA[cur ^ 1][r] = pre; else if (isB) sB[cur ^ 1][r] = pre;
    }
    __syncthreads();
  }

  // fused epilogue
  const float zpf = (float)ep.zp_out;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx + 16 * j;
    if (n >= g.out_cp) continue;
    const bool real = n < g.N;
    const int32_t ocn = real ? __ldg(ep.oc + n) : 0;
    const float bf = (real && ep.bias_f) ? __ldg(ep.bias_f + n) : 0.f;
    const float sbn = (real && ep.sb_vec) ? __ldg(ep.sb_vec + n) : ep.sb;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty + 16 * i;
      if (m >= g.M) continue;
      uint32_t q = (uint32_t)ep.zp_out;  // pad lanes carry the zero point
      if (real) {
        int32_t v = acc[i][j] + ocn;
        if (ep.bias_f) v = fc_bias_add(v, bf);
        if (ep.acc_out) ep.acc_out[(size_t)m * g.N + n] = v;
        q = requant_u8(v, ep.sa, sbn, ep.sc, zpf);
        if (ep.relu) q = max(q, (uint32_t)ep.zp_out);
      }
      y[(size_t)m * g.out_cp + n] = (uint8_t)q;
    }
  }
}

// ---- fc with a handful of outputs (classifier heads: N <= 16) ------------------------------------
// One block per input row, each warp takes a K slice: a lane reads 16 bytes of the row and the
// matching 16 bytes of all 16 weight rows (17 independent loads in flight), 64 dp4a.u32.s32, then a
// transpose-reduce (16 shuffles) leaves output j = lane>>1 in every lane; warps fold through shared
// memory and lane j of warp 0 runs the fused epilogue. Replaces a 128-row tensor-core tile that
// would be >90 % padding plus its split-K reduction.
constexpr int kHeadN = 16;
constexpr int kHeadWarps = 8;

__global__ void __launch_bounds__(kHeadWarps * 32) fc_head_kernel(const uint8_t* __restrict__ x, int ldx,
                                                                  const int8_t* __restrict__ w, int ldw,
                                                                  uint8_t* __restrict__ y, int ldy, int n,
                                                                  int kvec, const EpiParams ep, int fast) {
  __shared__ int32_t part[kHeadWarps][kHeadN];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)m * ldx);
  int32_t acc[kHeadN];
#pragma unroll
  for (int j = 0; j < kHeadN; ++j) acc[j] = 0;
  for (int v = warp * 32 + lane; v < kvec; v += kHeadWarps * 32) {
    const uint4 a = __ldg(xr + v);
    uint4 b[kHeadN];
#pragma unroll
    for (int j = 0; j < kHeadN; ++j) b[j] = __ldg(reinterpret_cast<const uint4*>(w + (size_t)j * ldw) + v);
#pragma unroll
    for (int j = 0; j < kHeadN; ++j) {
      int32_t s = acc[j];
      s = dp4a_u8s8(a.x, b[j].x, s); s = dp4a_u8s8(a.y, b[j].y, s);
      s = dp4a_u8s8(a.z, b[j].z, s); s = dp4a_u8s8(a.w, b[j].w, s);
      acc[j] = s;
    }
  }
  // transpose-reduce: after the step with offset o, a lane keeps the half of its values selected by
  // that lane bit; 8 + 4 + 2 + 1 exchanges, then one plain fold over the last lane bit
#pragma unroll
  for (int o = 16, cnt = kHeadN / 2; cnt >= 1; o >>= 1, cnt >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < cnt; ++j) {
      const int32_t send = up ? acc[j] : acc[j + cnt];
      const int32_t keep = up ? acc[j + cnt] : acc[j];
      acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
  // lane holds output index bits (lane>>4 &1)*8 + (lane>>3 &1)*4 + (lane>>2 &1)*2 + (lane>>1 &1) = lane >> 1
  if ((lane & 1) == 0) part[warp][lane >> 1] = acc[0];
  __syncthreads();
  if (warp != 0 || lane >= kHeadN) return;
  int32_t sum = 0;
#pragma unroll
  for (int k = 0; k < kHeadWarps; ++k) sum += part[k][lane];
  uint32_t q = (uint32_t)ep.zp_out;   // pad lanes carry the zero point
  if (lane < n) {
    int32_t v = sum + __ldg(ep.oc + lane);
    if (ep.bias_f) v = fc_bias_add(v, __ldg(ep.bias_f + lane));
    if (ep.acc_out) ep.acc_out[(size_t)m * n + lane] = v;
    const float zpf = (float)ep.zp_out;
    const float sbn = ep.sb_vec ? __ldg(ep.sb_vec + lane) : ep.sb;
    q = fast ? requant_u8_fast(v, ep.sa, sbn, ep.sc, __frcp_rn(ep.sc), zpf) : requant_u8(v, ep.sa, sbn, ep.sc, zpf);
    if (ep.relu) q = max(q, (uint32_t)ep.zp_out);
    // Module.__call__'s final dequantize (module.py:22-24 -> quantize_utils.cc:54-58) of a classifier head
    if (ep.deq_out) ep.deq_out[(size_t)m * n + lane] = dequant_f32((uint8_t)q, ep.zp_out, ep.sc);
  }
  y[(size_t)m * ldy + lane] = (uint8_t)q;
}

}  // namespace

bool fc_head_eligible(int n_pad, int ldx, int ldw, int ldy, const void* x, const void* w) {
  return n_pad == kHeadN && ldy == kHeadN && ldx % 16 == 0 && ldw % 16 == 0 && ldw >= ldx &&
         (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0;
}

int launch_fc_head(const uint8_t* x, int ldx, const int8_t* w, int ldw, uint8_t* y, int ldy, int m, int n, int k,
                   const EpiParams& ep, cudaStream_t stream) {
  const int kvec = (k + 15) / 16;   // the [k, ldx) tail multiplies zero weight lanes
  PdlFamily fam_(kPdlFcHead);
  launch_pdl(fc_head_kernel, dim3(m), dim3(kHeadWarps * 32), 0, stream, x, ldx, w, ldw, y, ldy, n, kvec, ep,
             requant_fast_ok(ep) ? 1 : 0);
  return check_launch("fc_head_kernel");
}

int launch_simt_igemm(const GemmGeom& g, const uint8_t* x, const int8_t* w, uint8_t* y,
                      const EpiParams& ep, int zp_in, cudaStream_t stream) {
  const uint32_t z = (uint32_t)(zp_in & 0xff);
  dim3 grid((g.M + BM - 1) / BM, (g.out_cp + BN - 1) / BN);
  launch_pdl(simt_igemm_kernel, grid, dim3(NT), 0, stream, g, x, w, y, ep, z | (z << 8) | (z << 16) | (z << 24));
  return check_launch("simt_igemm_kernel");
}

}  // namespace i8ie
