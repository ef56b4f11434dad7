// Multi-GPU result exchange over NVLink / NVSwitch peer memory (the only exchange of
// the batch-sharded forward is the gather of the fp32 logits and the top-1 agreement count).
//
// One process per GPU. Every rank owns ONE shareable device buffer
//     [ flags: kMaxWorld x u64 | gathered[2][world][chunk] ]
// opened by all peers through CUDA IPC. A step is two kernels on the forward's stream, no host
// involvement, no NCCL call — so forward + exchange replay as ONE CUDA graph (one enqueue per step):
//   top1_pack_push_kernel   argmax agreement count + logits of this rank, written with peer stores
//                           into slot `rank` of gathered[seq & 1] on EVERY rank (all-gather by push),
//                           then seq is published in word `rank` of every rank's flags (release.sys)
//   top1_wait_unpack_kernel waits until all `world` flag words reached seq (acquire.sys, bounded),
//                           then unpacks gathered[seq & 1] from LOCAL memory
// seq is a device-resident step counter (incremented by the pack kernel), so a replayed graph needs
// no per-step argument. Two gather buffers suffice: a peer can run at most one step ahead, because
// its step seq+2 push comes after its step seq+1 unpack, which needs this rank's seq+1 flag.
#include <cstring>

#include "common.cuh"

namespace i8ie {

int* tc_error_sink_device_ptr();   // tc_gemm.cu: device address of the host-mapped protocol-error flag (or nullptr)

namespace {

constexpr int kMaxWorld = 16;
constexpr int kFlagBytes = 256;   // kMaxWorld x u64, padded

struct ExchangeHeader {   // same chunk layout as elementwise.cu: int64 agree | int32 rows | int32 cols | fp32 logits
  long long agree;
  int rows, cols;
};

struct PeerTable {
  uint8_t* base[kMaxWorld];   // base[p] = rank p's shared buffer as mapped in THIS process
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(1024) top1_pack_push_kernel(const float* __restrict__ logits,
                                                              const long long* __restrict__ ref_argmax, int rows,
                                                              int cols, PeerTable peers, int world, int rank,
                                                              long long chunk_bytes, unsigned long long* seq_counter) {
  __shared__ int s_cnt[32];
  pdl_launch_dependents();
  pdl_wait();   // resident while the classifier head still runs; its logits are read from here on
  const unsigned long long seq = *seq_counter + 1ull;
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
    if (threadIdx.x == 0) s_cnt[0] = cnt;
  }
  __syncthreads();
  const long long agree = s_cnt[0];
  const size_t slot_off = (size_t)kFlagBytes + ((size_t)(seq & 1ull) * world + rank) * (size_t)chunk_bytes;
  const int n = rows * cols;
  for (int p = 0; p < world; ++p) {
    uint8_t* chunk = peers.base[p] + slot_off;
    if (threadIdx.x == 0) {
      ExchangeHeader* h = reinterpret_cast<ExchangeHeader*>(chunk);
      h->agree = agree; h->rows = rows; h->cols = cols;
    }
    float* dst = reinterpret_cast<float*>(chunk + sizeof(ExchangeHeader));
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = logits[i];
  }
  __threadfence_system();   // every thread's peer stores are ordered before the flag below
  __syncthreads();
  if (threadIdx.x < world)
    st_release_sys_u64(reinterpret_cast<unsigned long long*>(peers.base[threadIdx.x]) + rank, seq);
  if (threadIdx.x == 0) *seq_counter = seq;
}

__global__ void __launch_bounds__(1024) top1_wait_unpack_kernel(const uint8_t* __restrict__ mine, int world,
                                                                long long chunk_bytes,
                                                                const unsigned long long* __restrict__ seq_counter,
                                                                float* __restrict__ logits_all,
                                                                long long* __restrict__ agree_total, int* __restrict__ err) {
  __shared__ int s_bad;
  pdl_launch_dependents();
  pdl_wait();   // resident while the pack kernel runs; it has completed (and bumped seq) from here on
  const unsigned long long seq = *seq_counter;   // the pack kernel of this step ran before us on this stream
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  if (threadIdx.x < world) {
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(mine) + threadIdx.x;
    long long t0 = 0;
    for (unsigned n = 0;; ++n) {
      if (ld_acquire_sys_u64(flag) >= seq) break;
      if ((n & 255u) == 255u) {
        const long long t = clock64();
        if (t0 == 0) t0 = t;
        else if (t - t0 > 30000000000ll) { s_bad = 1; break; }   // ~15 s: a peer died or never launched
      }
    }
  }
  __syncthreads();
  if (s_bad) {
    if (threadIdx.x == 0 && err) { *reinterpret_cast<volatile int*>(err) = 7; __threadfence_system(); }
    return;
  }
  const uint8_t* gathered = mine + kFlagBytes + (size_t)(seq & 1ull) * world * (size_t)chunk_bytes;
  long long total = 0;
  int row0 = 0;
  for (int k = 0; k < world; ++k) {
    const uint8_t* chunk = gathered + (size_t)k * chunk_bytes;
    ExchangeHeader h;   // written by a peer: read past L1
    h.agree = __ldcv(reinterpret_cast<const long long*>(chunk));
    h.rows = __ldcv(reinterpret_cast<const int*>(chunk + 8));
    h.cols = __ldcv(reinterpret_cast<const int*>(chunk + 12));
    const float* src = reinterpret_cast<const float*>(chunk + sizeof(ExchangeHeader));
    float* dst = logits_all + (size_t)row0 * h.cols;
    for (int i = threadIdx.x; i < h.rows * h.cols; i += blockDim.x) dst[i] = __ldcv(src + i);
    total += h.agree;
    row0 += h.rows;
  }
  if (threadIdx.x == 0) *agree_total = total;
}

}  // namespace
}  // namespace i8ie

using namespace i8ie;

extern "C" {

int64_t i8ie_peer_exchange_bytes(int world, int64_t chunk_bytes) {
  return (int64_t)kFlagBytes + 2 * (int64_t)world * chunk_bytes;
}

int i8ie_peer_alloc(int64_t bytes, void** ptr, void* handle64) {
  I8IE_REQUIRE(bytes > 0 && ptr && handle64, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  I8IE_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
  I8IE_CUDA_OK(cudaMemset(p, 0, (size_t)bytes));
  I8IE_CUDA_OK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return I8IE_ECUDA;
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return I8IE_OK;
}

int i8ie_peer_open(const void* handle64, void** ptr) {
  I8IE_REQUIRE(handle64 && ptr, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  I8IE_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return I8IE_OK;
}

int i8ie_peer_close(void* ptr) {
  if (ptr) I8IE_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return I8IE_OK;
}

int i8ie_peer_free(void* ptr) {
  if (ptr) I8IE_CUDA_OK(cudaFree(ptr));
  return I8IE_OK;
}

int i8ie_top1_pack_push(const float* logits, const int64_t* ref_argmax, int rows, int cols, void* const* peer_bases,
                        int world, int rank, int64_t chunk_bytes, void* seq_counter, void* stream) {
  I8IE_REQUIRE(logits && peer_bases && seq_counter && rows >= 0 && cols > 0 && world >= 1 && world <= kMaxWorld &&
                   rank >= 0 && rank < world && chunk_bytes >= (int64_t)sizeof(ExchangeHeader) + 4ll * rows * cols &&
                   chunk_bytes % 16 == 0,
               "top1_pack_push: bad arguments (world=%d rank=%d rows=%d cols=%d chunk=%lld)", world, rank, rows, cols,
               (long long)chunk_bytes);
  PeerTable t{};
  for (int p = 0; p < world; ++p) {
    I8IE_REQUIRE(peer_bases[p] != nullptr, "top1_pack_push: peer %d is not mapped", p);
    t.base[p] = reinterpret_cast<uint8_t*>(peer_bases[p]);
  }
  PdlFamily fam_(kPdlExchange);
  launch_pdl(top1_pack_push_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, logits,
             reinterpret_cast<const long long*>(ref_argmax), rows, cols, t, world, rank, (long long)chunk_bytes,
             reinterpret_cast<unsigned long long*>(seq_counter));
  return check_launch("top1_pack_push_kernel");
}

int i8ie_top1_wait_unpack(const void* mine, int world, int64_t chunk_bytes, const void* seq_counter, float* logits_all,
                          int64_t* agree_total, void* stream) {
  I8IE_REQUIRE(mine && seq_counter && logits_all && agree_total && world >= 1 && world <= kMaxWorld && chunk_bytes > 0,
               "top1_wait_unpack: bad arguments");
  PdlFamily fam_(kPdlExchange);
  launch_pdl(top1_wait_unpack_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream,
             reinterpret_cast<const uint8_t*>(mine), world, (long long)chunk_bytes,
             reinterpret_cast<const unsigned long long*>(seq_counter), logits_all, reinterpret_cast<long long*>(agree_total),
             tc_error_sink_device_ptr());
  return check_launch("top1_wait_unpack_kernel");
}

}  // extern "C"
