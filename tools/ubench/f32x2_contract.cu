// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into ONE FFMA2 (it never contracts the scalar .rn forms):
//   nvcc -O3 [-fmad=false] -gencode arch=compute_100a,code=sm_100a -cubin -o c.cubin f32x2_contract.cu && cuobjdump -sass c.cubin | grep -E "FMUL|FADD|FFMA|Function"
// CUDA 12.9: k_scalar -> FMUL + FADD;  k_packed, k_add_as_fma, k_mul_as_fma, k_opaque64 -> a single FFMA2;
//            k_opaque32 (value split into two 32-bit halves through an empty volatile asm) -> FMUL2 + FADD2.
// A chain of separately rounded multiplies and adds (the reference's MODE-DOTA M-step) therefore cannot be written with
// f32x2 without changing bits; multiplying by a power of two before the add is exact and safe (the kNN distance).
typedef unsigned long long u64;
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__global__ void k_scalar(const float* a, const float* b, const float* c, float* o) {
  const int i = threadIdx.x;
  float d, e;
  asm("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a[i]), "f"(b[i]));
  asm("add.rn.f32 %0, %1, %2;" : "=f"(e) : "f"(d), "f"(c[i]));
  o[i] = e;
}
__global__ void k_packed(const u64* a, const u64* b, const u64* c, u64* o) { const int i = threadIdx.x; o[i] = add2(mul2(a[i], b[i]), c[i]); }
__global__ void k_add_as_fma(const u64* a, const u64* b, const u64* c, u64* o) {
  const int i = threadIdx.x;
  o[i] = fma2(mul2(a[i], b[i]), 0x3f8000003f800000ull, c[i]);
}
__global__ void k_mul_as_fma(const u64* a, const u64* b, const u64* c, u64* o) {
  const int i = threadIdx.x;
  o[i] = add2(fma2(a[i], b[i], 0x8000000080000000ull), c[i]);
}
__global__ void k_opaque64(const u64* a, const u64* b, const u64* c, u64* o) {
  const int i = threadIdx.x;
  u64 m = mul2(a[i], b[i]);
  asm volatile("" : "+l"(m));
  o[i] = add2(m, c[i]);
}
__global__ void k_opaque32(const u64* a, const u64* b, const u64* c, u64* o) {
  const int i = threadIdx.x;
  const u64 m = mul2(a[i], b[i]);
  unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
  asm volatile("" : "+r"(lo), "+r"(hi));
  o[i] = add2(((u64)hi << 32) | lo, c[i]);
}
