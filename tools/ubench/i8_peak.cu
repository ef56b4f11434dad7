// Microbenchmarks behind the roofline denominators of the tcgen05 conv / fc kernels (B200, sm_100a).
//
//   mma   tcgen05.mma kind::i8 issue rate with shared-memory-resident operands and NO loads:
//         the tensor-pipe peak the conv kernels are measured against (UTCIMMA only).
//         Variants: cta_group::1 (M=128) and cta_group::2 (M=256), N in {128, 192, 256}.
//   l2    L2 -> SM delivery rate: every SM streams an L2-resident buffer into shared memory with
//         16 KB bulk copies (UBLKCP), 8 copies in flight, nothing else running.
//   both  the two at once on the same SMs (do they share a limit / a power budget?).
//
// Build (tools/ubench/build.sh):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -cudart shared
// Prints one JSON object per measurement on stdout. Dev tool: not part of libi8ie_sm100.so.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../int8inferenceengine_b200/csrc/tc_ptx.cuh"

using namespace i8ie;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

__device__ int g_err = 0;

struct Ctl {
  uint64_t done[2];
  uint64_t full[8];
  uint32_t tmem_slot;
};

// ---- MMA issue rate -----------------------------------------------------------------------------
// One CTA (or CTA pair) per SM. Operands: A 128 x 128 B and B (BN or BN/2) x 128 B, 128-byte swizzle
// K-major (content irrelevant for timing; filled with a byte pattern so the datapath toggles).
// Thread 0 of warp 0 issues `batches` batches of `per` K blocks (4 MMAs of K = 32 each), two
// accumulators alternating, one commit per batch on alternating barriers (the issuer never runs
// more than two batches ahead of the tensor pipe).
template <int BN, int CG>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int batches, int per, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kA = 128 * 128, kB = (BN / CG) * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + kA;
  Ctl* ctl = reinterpret_cast<Ctl*>(sB + kB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (kA + kB) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x01020304u * (uint32_t)(i % 61 + 1);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&ctl->done[0], 1);
    ptx::mbar_init(&ctl->done[1], 1);
    ptx::fence_barrier_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (CG == 2) ptx::tmem_alloc_2cta(&ctl->tmem_slot, 512);
    else ptx::tmem_alloc(&ctl->tmem_slot, 512);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = ctl->tmem_slot;
  if (CG == 2) ptx::cluster_sync_all();
  const bool leader = CG == 1 || (blockIdx.x & 1) == 0;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && lane == 0 && leader) {
    constexpr uint32_t idesc = ptx::make_idesc_i8(128 * CG, BN);
    const uint32_t hi = ptx::smem_desc_hi<128>();
    const uint32_t a_lo = ptx::smem_desc_lo(ptx::smem_u32(sA)), b_lo = ptx::smem_desc_lo(ptx::smem_u32(sB));
    t0 = clock64();
    for (int b = 0; b < batches; ++b) {
      if (b >= 2 && !ptx::mbar_wait(&ctl->done[b & 1], ((b - 2) >> 1) & 1)) { atomicExch(&g_err, 1); break; }
      const uint32_t d = tmem + (uint32_t)(b & 1) * 256u;
      for (int i = 0; i < per; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (CG == 2) ptx::mma_i8_ss_lohi_2cta(d, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, (i | k) ? 1u : 0u);
          else ptx::mma_i8_ss_lohi(d, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, (i | k) ? 1u : 0u);
        }
      }
      if (CG == 2) ptx::tc_commit_2cta_multicast(&ctl->done[b & 1], 1);
      else ptx::tc_commit(&ctl->done[b & 1]);
    }
    for (int b = (batches >= 2 ? batches - 2 : 0); b < batches; ++b)
      if (!ptx::mbar_wait(&ctl->done[b & 1], (b >> 1) & 1)) atomicExch(&g_err, 2);
    t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync_all();
  if (warp == 0) {
    if (CG == 2) ptx::tmem_dealloc_2cta(tmem, 512);
    else ptx::tmem_dealloc(tmem, 512);
  }
}

// ---- L2 -> SM delivery ---------------------------------------------------------------------------
// Each CTA streams `chunks` 16 KB chunks of its own slice (slice_bytes, revisited round-robin) into an
// 8-slot shared-memory ring with cp.async.bulk; the same thread waits for a slot and refills it.
__global__ void __launch_bounds__(128, 1) l2_stream_kernel(const uint8_t* __restrict__ src, size_t slice_bytes,
                                                           int chunks, int stride_ctas, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kChunk = 16384, kSlots = 8;
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + kSlots * kChunk);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) ptx::mbar_init(&ctl->full[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* base = src + (size_t)(blockIdx.x % stride_ctas) * slice_bytes;
    const int per_slice = (int)(slice_bytes / kChunk);
    const long long t0 = clock64();
    for (int c = 0; c < chunks + kSlots; ++c) {
      const int s = c % kSlots;
      if (c >= kSlots && !ptx::mbar_wait(&ctl->full[s], ((c - kSlots) / kSlots) & 1)) { atomicExch(&g_err, 3); break; }
      if (c < chunks) {
        ptx::mbar_arrive_expect_tx(&ctl->full[s], kChunk);
        ptx::bulk_load_1d(smem + s * kChunk, base + (size_t)(c % per_slice) * kChunk, kChunk, &ctl->full[s]);
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

// ---- DRAM -> SM through bulk copies: same ring, chunk size and ring depth as parameters, every chunk a
// DRAM miss (each CTA walks its own region of a buffer far larger than L2, touched once) ------------------
__global__ void __launch_bounds__(128, 1) dram_stream_kernel(const uint8_t* __restrict__ src, size_t region_bytes,
                                                             int chunk, int slots, int chunks, long long* cycles,
                                                             int prefetch_ahead) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + 8 * 16384);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) ptx::mbar_init(&ctl->full[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* base = src + (size_t)blockIdx.x * region_bytes;
    for (int c = 0; c < prefetch_ahead && c < chunks; ++c) ptx::prefetch_l2_bulk(base + (size_t)c * chunk, chunk);
    const long long t0 = clock64();
    for (int c = 0; c < chunks + slots; ++c) {
      const int s = c % slots;
      if (c >= slots && !ptx::mbar_wait(&ctl->full[s], ((c - slots) / slots) & 1)) { atomicExch(&g_err, 4); break; }
      if (c < chunks) {
        if (prefetch_ahead > 0 && c + prefetch_ahead < chunks) ptx::prefetch_l2_bulk(base + (size_t)(c + prefetch_ahead) * chunk, chunk);
        ptx::mbar_arrive_expect_tx(&ctl->full[s], chunk);
        ptx::bulk_load_1d(smem + s * 16384, base + (size_t)c * chunk, chunk, &ctl->full[s]);
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

// ---- both at once: warp 0 issues MMAs (single CTA, N = 256), warp 1 streams from L2 ----------------
__global__ void __launch_bounds__(128, 1) both_kernel(int batches, int per, const uint8_t* __restrict__ src,
                                                      size_t slice_bytes, int chunks, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kA = 128 * 128, kB = 256 * 128, kChunk = 16384, kSlots = 8;
  uint8_t* sA = smem;
  uint8_t* sB = smem + kA;
  uint8_t* sR = sB + kB;
  Ctl* ctl = reinterpret_cast<Ctl*>(sR + kSlots * kChunk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (kA + kB) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x01020304u * (uint32_t)(i % 61 + 1);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&ctl->done[0], 1);
    ptx::mbar_init(&ctl->done[1], 1);
    for (int s = 0; s < kSlots; ++s) ptx::mbar_init(&ctl->full[s], 1);
    ptx::fence_barrier_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) ptx::tmem_alloc(&ctl->tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = ctl->tmem_slot;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = ptx::make_idesc_i8(128, 256);
    const uint32_t hi = ptx::smem_desc_hi<128>();
    const uint32_t a_lo = ptx::smem_desc_lo(ptx::smem_u32(sA)), b_lo = ptx::smem_desc_lo(ptx::smem_u32(sB));
    const long long t0 = clock64();
    for (int b = 0; b < batches; ++b) {
      if (b >= 2 && !ptx::mbar_wait(&ctl->done[b & 1], ((b - 2) >> 1) & 1)) { atomicExch(&g_err, 1); break; }
      const uint32_t d = tmem + (uint32_t)(b & 1) * 256u;
      for (int i = 0; i < per; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mma_i8_ss_lohi(d, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, (i | k) ? 1u : 0u);
      ptx::tc_commit(&ctl->done[b & 1]);
    }
    for (int b = (batches >= 2 ? batches - 2 : 0); b < batches; ++b)
      if (!ptx::mbar_wait(&ctl->done[b & 1], (b >> 1) & 1)) atomicExch(&g_err, 2);
    cycles[blockIdx.x] = clock64() - t0;
  } else if (warp == 1 && lane == 0) {
    const uint8_t* base = src + (size_t)blockIdx.x * slice_bytes;
    const int per_slice = (int)(slice_bytes / kChunk);
    const long long t0 = clock64();
    for (int c = 0; c < chunks + kSlots; ++c) {
      const int s = c % kSlots;
      if (c >= kSlots && !ptx::mbar_wait(&ctl->full[s], ((c - kSlots) / kSlots) & 1)) { atomicExch(&g_err, 3); break; }
      if (c < chunks) {
        ptx::mbar_arrive_expect_tx(&ctl->full[s], kChunk);
        ptx::bulk_load_1d(sR + s * kChunk, base + (size_t)(c % per_slice) * kChunk, kChunk, &ctl->full[s]);
      }
    }
    cycles[gridDim.x + blockIdx.x] = clock64() - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

static double med(std::vector<long long> v) {
  std::vector<long long> w;
  for (auto x : v) if (x > 0) w.push_back(x);
  if (w.empty()) return 0;
  std::sort(w.begin(), w.end());
  return (double)w[w.size() / 2];
}

template <int BN, int CG>
static void run_mma(int sms, int batches, int per, long long* d_cyc) {
  const int smem = 128 * 128 + (BN / CG) * 128 + (int)sizeof(Ctl) + 1024;
  auto kern = mma_rate_kernel<BN, CG>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sms); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  std::vector<long long> cyc(sms);
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaMemset(d_cyc, 0, sizeof(long long) * 1024));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kern, batches, per, d_cyc));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) { best = ms; CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost)); }
  }
  const double mmas = (double)batches * per * 4;
  const double ops = 2.0 * 128 * BN * 32 * mmas * sms;   // per CTA: its 128 rows of every instruction
  const double clk = med(cyc);
  printf("{\"bench\": \"mma\", \"cta_group\": %d, \"M\": %d, \"N\": %d, \"ctas\": %d, \"mma_per_cta\": %.0f, \"ms\": %.4f, "
         "\"tops\": %.1f, \"clk_per_mma\": %.2f, \"mhz_effective\": %.0f}\n",
         CG, 128 * CG, BN, sms, mmas, best, ops / (best * 1e-3) / 1e12, clk / mmas, clk / (best * 1e-3) / 1e6);
  fflush(stdout);
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, sizeof(long long) * 1024));
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);

  // short (~30 us: the duration of a conv layer at batch 100) and long (~1 ms: sustained, power capped) runs
  for (int pass = 0; pass < 2; ++pass) {
    const int batches = pass == 0 ? 8 : 256, per = 16;   // 512 / 16384 MMAs per CTA
    run_mma<256, 1>(sms, batches, per, d_cyc);
    run_mma<192, 1>(sms, batches, per, d_cyc);
    run_mma<128, 1>(sms, batches, per, d_cyc);
    run_mma<256, 2>(sms, batches, per, d_cyc);
    run_mma<192, 2>(sms, batches, per, d_cyc);
    run_mma<128, 2>(sms, batches, per, d_cyc);
  }

  // L2 delivery: 512 KB per CTA (74 MB in total: resident in the 126 MB L2 after the first pass)
  const size_t slice = 512 << 10;
  uint8_t* src;
  CK(cudaMalloc(&src, slice * sms));
  CK(cudaMemset(src, 1, slice * sms));
  const int smem_l2 = 8 * 16384 + (int)sizeof(Ctl) + 1024;
  CK(cudaFuncSetAttribute(l2_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_l2));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int ctas : {sms, sms / 2, 8, 1}) {
    for (int shared : {0, 1}) {   // shared: every CTA reads the SAME slice (what conv CTAs do with the weights)
      const int chunks = 2048;    // 32 MB per CTA
      float best = 1e30f;
      std::vector<long long> cyc(ctas);
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemset(d_cyc, 0, sizeof(long long) * 1024));
        CK(cudaEventRecord(e0));
        l2_stream_kernel<<<ctas, 128, smem_l2>>>(src, slice, chunks, shared ? 1 : ctas, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) { best = ms; CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost)); }
      }
      const double bytes = (double)chunks * 16384 * ctas;
      const double clk = med(cyc);
      printf("{\"bench\": \"l2\", \"ctas\": %d, \"same_slice\": %d, \"ms\": %.4f, \"gbs\": %.1f, \"bytes_per_clk_per_sm\": %.2f, "
             "\"bytes_per_clk_chip\": %.0f}\n",
             ctas, shared, best, bytes / (best * 1e-3) / 1e9, (double)chunks * 16384 / clk, (double)chunks * 16384 / clk * ctas);
      fflush(stdout);
    }
  }

  // DRAM -> SM bulk-copy stream (every chunk misses L2): how much must be in flight per SM?
  {
    const size_t region = 8u << 20;   // 8 MB per CTA, 1.2 GB in all: nothing is re-read
    uint8_t* big;
    CK(cudaMalloc(&big, region * sms));
    CK(cudaMemset(big, 1, region * sms));
    CK(cudaFuncSetAttribute(dram_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_l2));
    uint8_t* flush;
    CK(cudaMalloc(&flush, 256u << 20));
    for (int chunk : {7168, 16384}) {
      for (int slots : {3, 8}) {
        for (int pf : {0, 16}) {
          const int chunks = (int)(region / chunk);
          CK(cudaMemset(flush, 2, 256u << 20));   // evict the region from L2
          CK(cudaMemset(d_cyc, 0, sizeof(long long) * 1024));
          CK(cudaEventRecord(e0));
          dram_stream_kernel<<<sms, 128, smem_l2>>>(big, region, chunk, slots, chunks, d_cyc, pf);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          std::vector<long long> cyc(sms);
          CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
          const double bytes = (double)chunks * chunk * sms;
          printf("{\"bench\": \"dram_bulk\", \"chunk\": %d, \"in_flight\": %d, \"l2_prefetch_ahead\": %d, \"ms\": %.4f, "
                 "\"gbs\": %.1f, \"bytes_per_clk_per_sm\": %.2f}\n",
                 chunk, slots, pf, ms, bytes / (ms * 1e-3) / 1e9, (double)chunks * chunk / med(cyc));
          fflush(stdout);
        }
      }
    }
    CK(cudaFree(flush));
    CK(cudaFree(big));
  }

  // both at once
  {
    const int smem = 128 * 128 + 256 * 128 + 8 * 16384 + (int)sizeof(Ctl) + 1024;
    CK(cudaFuncSetAttribute(both_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int chunks : {0, 256, 512, 768, 1024}) {
      const int batches = 32, per = 16;   // 2048 MMAs of 128 clk = 262 k clk
      float best = 1e30f;
      std::vector<long long> cyc(2 * sms);
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemset(d_cyc, 0, sizeof(long long) * 1024));
        CK(cudaEventRecord(e0));
        both_kernel<<<sms, 128, smem>>>(batches, per, src, slice, chunks, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) { best = ms; CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost)); }
      }
      std::vector<long long> cm(cyc.begin(), cyc.begin() + sms), cl(cyc.begin() + sms, cyc.end());
      const double mmas = (double)batches * per * 4;
      printf("{\"bench\": \"both\", \"l2_chunks_per_cta\": %d, \"ms\": %.4f, \"mma_clk_per_mma\": %.2f, "
             "\"l2_bytes_per_clk_per_sm\": %.2f, \"tops\": %.1f}\n",
             chunks, best, med(cm) / mmas, chunks ? (double)chunks * 16384 / med(cl) : 0.0,
             2.0 * 128 * 256 * 32 * mmas * sms / (best * 1e-3) / 1e12);
      fflush(stdout);
    }
  }
  int err = 0;
  CK(cudaMemcpyFromSymbol(&err, g_err, sizeof(int)));
  printf("{\"protocol_error\": %d}\n", err);
  return err ? 3 : 0;
}
