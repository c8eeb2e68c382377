// Microbenchmark: issue throughput of scalar FADD / FMUL vs packed FADD2 / FMUL2 (f32x2) on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float mul1(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8]; u64 p[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
  const float c = seed * 0.999f; const u64 c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) a[i] = add1(a[i], c);
        if (MODE == 1) a[i] = mul1(a[i], c);
        if (MODE == 2) p[i] = add2(p[i], c2);
        if (MODE == 3) p[i] = mul2(p[i], c2);
        if (MODE == 4) { a[i] = (i & 1) ? add1(a[i], c) : mul1(a[i], c); }       // alternate pipes?
        if (MODE == 5) { p[i] = (i & 1) ? add2(p[i], c2) : mul2(p[i], c2); }
      }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  if (s == 12345.678f) out[0] = s;
}
template <int MODE> void run(const char* name, float* out) {
  const int iters = 4096, blocks = 148 * 2, threads = 1024;
  k<MODE><<<blocks, threads>>>(out, 16, 1.0f); cudaDeviceSynchronize();
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  cudaEventRecord(s); k<MODE><<<blocks, threads>>>(out, iters, 1.0f); cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  const double inst = (double)blocks * threads / 32 * iters * 32;   // warp instructions
  const double per_sm_clk = inst / 148 / (ms * 1e-3 * 1.965e9);
  printf("%-22s %8.3f ms  %.2f warp-instr/clk/SM  (%.0f fp32 lane-ops/clk/SM)\n", name, ms, per_sm_clk, per_sm_clk * 32 * (MODE == 2 || MODE == 3 || MODE == 5 ? 2 : 1));
}
int main() {
  float* out; cudaMalloc(&out, 4);
  run<0>("FADD", out); run<1>("FMUL", out); run<2>("FADD2 (f32x2)", out); run<3>("FMUL2 (f32x2)", out);
  run<4>("FADD/FMUL alternating", out); run<5>("FADD2/FMUL2 alternating", out);
  return 0;
}
