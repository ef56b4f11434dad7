// Dev microbenchmark: issue rate of the instructions the requantise epilogue is made of
// (cycles per warp-instruction per SM sub-partition), B200. Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 512
#define CHAINS 8

template <int OP>
__global__ void bench(uint32_t* out, long long* cyc, float s, int si) {
  uint32_t v[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x * 977u + c * 131u + 12345u;
  unsigned long long pk[CHAINS / 2];
#pragma unroll
  for (int c = 0; c < CHAINS / 2; ++c) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pk[c]) : "r"(v[2 * c] & 0x3fffffffu | 0x3f000000u), "r"(v[2 * c + 1] & 0x3fffffffu | 0x3f000000u));
  unsigned long long ss;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (OP == 0) asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(v[c]));
      if (OP == 1) asm volatile("cvt.rzi.s32.f32 %0, %0;" : "+r"(v[c]));
      if (OP == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+r"(v[c]) : "r"(__float_as_uint(s)));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+r"(v[c]) : "r"(__float_as_uint(s)));
      if (OP == 4) asm volatile("add.rm.f32 %0, %0, %1;" : "+r"(v[c]) : "r"(__float_as_uint(s)));
      if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+r"(v[c]) : "r"(__float_as_uint(s)));
      if (OP == 6) asm volatile("add.s32 %0, %0, %1;" : "+r"(v[c]) : "r"(si));
      if (OP == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(si), "r"(it));
      if (OP == 8) asm volatile("prmt.b32 %0, %0, %1, 0x7650;" : "+r"(v[c]) : "r"(si));
      if (OP == 9) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(si), "r"(it));
      if (OP == 10 && c < CHAINS / 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(pk[c]) : "l"(ss));
      if (OP == 11 && c < CHAINS / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(pk[c]) : "l"(ss));
      if (OP == 12 && c < CHAINS / 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pk[c]) : "l"(ss));
      if (OP == 13) asm volatile("cvt.rzi.sat.u8.f32 %0, %0;" : "+r"(v[c]));   // PTX clamps narrow int results
      if (OP == 14) asm volatile("shr.s32 %0, %0, 16;" : "+r"(v[c]));
      if (OP == 15) asm volatile("vmax4.u32.u32.u32 %0, %0, %1, %1;" : "+r"(v[c]) : "r"(si));
      if (OP == 16) {   // mixed: mul (fma pipe) + add.s32 (alu pipe) alternating
        if (c & 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+r"(v[c]) : "r"(__float_as_uint(s)));
        else asm volatile("add.s32 %0, %0, %1;" : "+r"(v[c]) : "r"(si));
      }
      if (OP == 17) {   // mixed: f32x2 mul + add.s32
        if (c < CHAINS / 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(pk[c]) : "l"(ss));
        else asm volatile("add.s32 %0, %0, %1;" : "+r"(v[c]) : "r"(si));
      }
      if (OP == 18) {   // mixed: cvt (xu) + mul
        if (c & 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+r"(v[c]) : "r"(__float_as_uint(s)));
        else asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(v[c]));
      }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc ^= v[c];
#pragma unroll
  for (int c = 0; c < CHAINS / 2; ++c) acc ^= (uint32_t)pk[c] ^ (uint32_t)(pk[c] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_iter) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  printf("%-28s", name);
  for (int wps : {1, 2, 4, 8}) {   // warps per SMSP
    const int threads = 128 * wps;
    bench<OP><<<148, threads>>>(out, cyc, 1.0000001f, 3);
    cudaDeviceSynchronize();
    bench<OP><<<148, threads>>>(out, cyc, 1.0000001f, 3);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    // cycles per warp-instruction per SMSP
    printf("  w%d: %6.2f", wps, avg / ((double)ITERS * instr_per_iter * wps));
  }
  printf("   (cyc / warp-instr / SMSP)\n");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("cvt.rn.f32.s32 (I2FP)", CHAINS);
  run<1>("cvt.rzi.s32.f32 (F2I)", CHAINS);
  run<13>("cvt.rzi.sat.u8.f32", CHAINS);
  run<2>("mul.rn.f32", CHAINS);
  run<3>("fma.rn.f32", CHAINS);
  run<4>("add.rm.f32", CHAINS);
  run<5>("max.f32", CHAINS);
  run<6>("add.s32", CHAINS);
  run<7>("lop3", CHAINS);
  run<8>("prmt", CHAINS);
  run<14>("shr.s32", CHAINS);
  run<15>("vmax4.u32", CHAINS);
  run<9>("cvt.pack.sat.u8.s32", CHAINS);
  run<10>("mul.rn.f32x2", CHAINS / 2);
  run<11>("fma.rn.f32x2", CHAINS / 2);
  run<12>("add.rn.f32x2", CHAINS / 2);
  run<16>("mix mul.f32 + add.s32", CHAINS);
  run<17>("mix mul.f32x2 + add.s32", CHAINS / 2 + CHAINS / 2);
  run<18>("mix cvt.f32.s32 + mul.f32", CHAINS);
  return 0;
}
