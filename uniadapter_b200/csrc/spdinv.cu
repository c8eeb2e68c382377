// DOTA.update on the device: Lambda = inverse((1-eps)*overall_Sigma + eps*I).half()   (dota.py:66-69).
//
// The matrix is symmetric positive definite by construction (a mean of sigma*I plus outer products, plus eps*I), so
// Gauss-Jordan elimination needs no pivoting. The library routes (cuSOLVER LU / Cholesky + solves) are chains of small
// launches: 1.3 / 0.5 ms at D=512 for 0.27 GFLOP. Here ONE cooperative launch does the whole inversion:
//
//   * the D x D matrix is cut into (16*RM) x (16*RN) tiles, one CTA per tile (<= 148 CTAs, all co-resident);
//     a CTA keeps its tile in REGISTERS from the first load to the final fp16 store (thread (ty,tx) of a 16x16
//     thread layout owns the elements (ty + 16r, tx + 16c));
//   * block Gauss-Jordan with 16-wide pivot blocks: at step k every CTA needs only the pivot row panel A[k,:] over its
//     columns, the pivot column panel A[:,k] over its rows and the 16x16 pivot block. The owners publish those slices
//     to a ping-pong panel buffer in global memory (L2-resident), ONE grid barrier per step, readers fetch them with
//     ld.global.cg; every CTA inverts the pivot block itself (one warp, registers + shuffles), forms
//     R = P * A[k,:] and applies the rank-16 update A -= A[:,k] * R to its register tile;
//   * in-place bookkeeping of Gauss-Jordan inversion (pivot rows become R, pivot columns become -A[:,k]*P, the pivot
//     block becomes P) is done by uniform per-micro-block selects: pivot rows / columns of a tile are one whole
//     micro-block row r* / column c*, so no thread diverges.
//
// After D/16 steps the tile holds the inverse. fp32 accuracy equals LAPACK's LU / Cholesky inverses on the matrices of
// the path (tests/test_gpu_adapters.py; numpy prototype: 6e-7..9e-7 of max|Lambda| vs float64 at cond 20..200).
#include "common.cuh"

namespace ua {
namespace {

constexpr int kNB = 16;          // pivot block width = micro-block width
constexpr int kInvThreads = 256; // 16 x 16 thread layout

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Correctly-rounded reciprocal of a normal positive float (MUFU.RCP + one Newton step in FMA): the same value as
// 1.0f / v without the range checks and slow path of the division sequence.
__device__ __forceinline__ float rcp_rn_pos(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  const float e = -fmaf(v, r, -1.0f);
  return fmaf(r, e, r);
}

// Inverse of a 16x16 SPD block by one warp: lane owns row (lane>>1), columns 8*(lane&1) .. +7. Unpivoted scalar
// Gauss-Jordan, fully unrolled so every register index is static. The chain of 16 dependent pivots is the serial core
// of the whole inversion (28 % of the kernel's stall samples), so the NEXT pivot's reciprocal is computed one step
// ahead from three extra shuffles of the pre-update state: a'[t+1][t+1] = fma(-a[t+1][t], a[t][t+1] * p, a[t+1][t+1])
// is bit-identical to what the update below writes, and the per-pivot chain shrinks from shuffle -> divide -> shuffle ->
// multiply -> fma to multiply -> fma -> reciprocal. Result written to Ps (row-major 16x16).
__device__ __forceinline__ void invert16_warp(const float* __restrict__ Akk /* smem 16x16 */, float* __restrict__ Ps,
                                              int lane) {
  const int i = lane >> 1, h = lane & 1;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = Akk[i * kNB + h * 8 + j];
  float p = rcp_rn_pos(__shfl_sync(kFullMask, a[0], 0));
#pragma unroll
  for (int t = 0; t < kNB; ++t) {
    // pivot row t lives in lanes 2t (cols 0-7) and 2t+1 (cols 8-15); column t lives in register t&7 of half t>>3
    const float cit = __shfl_sync(kFullMask, a[t & 7], (lane & ~1) | (t >> 3));  // a[i][t]
    float row[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) row[j] = __shfl_sync(kFullMask, a[j], 2 * t + h);  // a[t][j]
    float pnext = 0.f;
    if (t + 1 < kNB) {
      const int u = t + 1;
      const float a11 = __shfl_sync(kFullMask, a[u & 7], 2 * u + (u >> 3));
      const float a10 = __shfl_sync(kFullMask, a[t & 7], 2 * u + (t >> 3));
      const float a01 = __shfl_sync(kFullMask, a[u & 7], 2 * t + (u >> 3));
      pnext = rcp_rn_pos(fmaf(-a10, a01 * p, a11));
    }
    const bool prow = (i == t);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool pcol = (h * 8 + j == t);
      const float rp = row[j] * p;
      float v;
      if (prow) v = pcol ? p : rp;
      else v = pcol ? -cit * p : fmaf(-cit, rp, a[j]);
      a[j] = v;
    }
    p = pnext;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) Ps[i * kNB + h * 8 + j] = a[j];
}

template <int RM, int RN>
__global__ void __launch_bounds__(kInvThreads, 1)
spd_inverse_kernel(const float* __restrict__ A_in, int D, float one_minus_eps, float eps,
                   float* __restrict__ rowpanel /* [2][16][D] */, float* __restrict__ colpanel /* [2][D][16] */,
                   unsigned* __restrict__ counter, __half* __restrict__ out_h, float* __restrict__ out_f) {
  constexpr int TM = kNB * RM, TN = kNB * RN;
  __shared__ __align__(16) float Cs[TM * kNB];    // -A[I, k] (pivot rows zeroed)
  __shared__ __align__(16) float Ro[kNB * TN];    // old A[k, J]
  __shared__ __align__(16) float Rs[kNB * TN];    // R = P * A[k, J]  (pivot columns: P)
  __shared__ __align__(16) float Akk[kNB * kNB];
  __shared__ __align__(16) float Ps[kNB * kNB];

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, lane = tid & 31, warp = tid >> 5;
  const int I0 = blockIdx.y * TM, J0 = blockIdx.x * TN;
  const unsigned G = gridDim.x * gridDim.y;
  const int steps = D / kNB;
  const size_t ld = (size_t)D;

  // ---- load the tile: a = (1-eps)*overall + eps*I, the arithmetic of dota.py:67 in fp32 -------------------------
  float a[RM][RN];
#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int c = 0; c < RN; ++c) {
      const int gi = I0 + ty + kNB * r, gj = J0 + tx + kNB * c;
      float v = 0.f;
      if (gi < D && gj < D) v = __fadd_rn(__fmul_rn(one_minus_eps, A_in[gi * ld + gj]), gi == gj ? eps : 0.f);
      a[r][c] = v;
    }

  // publish the slices of this tile that lie in pivot row / column block `k` into panel buffer `par`
  auto publish = [&](int k, int par) {
    const int k0 = k * kNB;
    float* rp = rowpanel + (size_t)par * kNB * ld;
    float* cp = colpanel + (size_t)par * ld * kNB;
    if (k0 >= I0 && k0 < I0 + TM) {
      const int rs = (k0 - I0) / kNB;
#pragma unroll
      for (int r = 0; r < RM; ++r)
        if (r == rs) {
#pragma unroll
          for (int c = 0; c < RN; ++c) {
            const int gj = J0 + tx + kNB * c;
            if (gj < D) rp[(size_t)ty * ld + gj] = a[r][c];
          }
        }
    }
    if (k0 >= J0 && k0 < J0 + TN) {
      const int cs = (k0 - J0) / kNB;
#pragma unroll
      for (int c = 0; c < RN; ++c)
        if (c == cs) {
#pragma unroll
          for (int r = 0; r < RM; ++r) {
            const int gi = I0 + ty + kNB * r;
            if (gi < D) cp[(size_t)gi * kNB + tx] = a[r][c];
          }
        }
    }
  };
  // the CTA's panel stores happen-before the barrier; thread 0's release (cumulative) orders them before the count
  auto arrive = [&]() {
    __syncthreads();
    if (tid == 0) red_release_add(counter, 1u);
  };

  publish(0, 0);
  arrive();

  for (int k = 0; k < steps; ++k) {
    const int k0 = k * kNB, par = k & 1;
    // ---- grid barrier: every CTA has arrived k+1 times => the panels of step k are complete, and nobody still reads
    //      the buffer (k+1)&1 that this step's publish will overwrite
    if (tid == 0) {
      const unsigned target = (unsigned)(k + 1) * G;
      while (ld_acquire_u32(counter) < target) {
      }
    }
    __syncthreads();

    const float* rp = rowpanel + (size_t)par * kNB * ld;
    const float* cp = colpanel + (size_t)par * ld * kNB;
    const bool row_in = (k0 >= I0 && k0 < I0 + TM), col_in = (k0 >= J0 && k0 < J0 + TN);
    const int rstar = row_in ? (k0 - I0) / kNB : -1, cstar = col_in ? (k0 - J0) / kNB : -1;

    // ---- fetch the panels (L2; the lines were written by other SMs: bypass L1) ---------------------------------
    for (int e = tid; e < TM * kNB / 4; e += kInvThreads) {       // column panel: TM x 16 contiguous
      const int row = e >> 2;                                      // 4 float4 per row
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (I0 + row < D && (row / kNB) != rstar) {
        v = __ldcg(reinterpret_cast<const float4*>(cp + (size_t)(I0 + row) * kNB) + (e & 3));
        v.x = -v.x; v.y = -v.y; v.z = -v.z; v.w = -v.w;
      }
      reinterpret_cast<float4*>(Cs)[e] = v;
    }
    for (int e = tid; e < kNB * TN / 4; e += kInvThreads) {       // row panel: 16 rows of TN
      const int t = e / (TN / 4), q = e - t * (TN / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (J0 + q * 4 < D) v = __ldcg(reinterpret_cast<const float4*>(rp + (size_t)t * ld + J0) + q);
      reinterpret_cast<float4*>(Ro)[e] = v;
    }
    if (tid < kNB * kNB / 4) {                                     // pivot block: columns k0.. of the row panel
      const int t = tid >> 2, q = tid & 3;
      reinterpret_cast<float4*>(Akk)[tid] = __ldcg(reinterpret_cast<const float4*>(rp + (size_t)t * ld + k0) + q);
    }
    __syncthreads();

    // ---- P = inv(A_kk) (warp 0), then R = P * A[k, J] --------------------------------------------------------
    if (warp == 0) invert16_warp(Akk, Ps, lane);
    __syncthreads();
    {
      float pr[kNB];
#pragma unroll
      for (int s = 0; s < kNB; s += 4) {
        const float4 v = *reinterpret_cast<const float4*>(Ps + ty * kNB + s);
        pr[s] = v.x; pr[s + 1] = v.y; pr[s + 2] = v.z; pr[s + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < RN; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int s = 0; s < kNB; ++s) acc = fmaf(pr[s], Ro[s * TN + tx + kNB * c], acc);
        Rs[ty * TN + tx + kNB * c] = (c == cstar) ? Ps[ty * kNB + tx] : acc;
      }
    }
    __syncthreads();

    // ---- rank-16 update of the register tile -------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
      for (int c = 0; c < RN; ++c)
        if (c == cstar) a[r][c] = 0.f;
#pragma unroll
    for (int t4 = 0; t4 < kNB; t4 += 4) {
      float4 cv[RM];
#pragma unroll
      for (int r = 0; r < RM; ++r) cv[r] = *reinterpret_cast<const float4*>(Cs + (ty + kNB * r) * kNB + t4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float rv[RN];
#pragma unroll
        for (int c = 0; c < RN; ++c) rv[c] = Rs[(t4 + u) * TN + tx + kNB * c];
#pragma unroll
        for (int r = 0; r < RM; ++r) {
          const float cval = u == 0 ? cv[r].x : u == 1 ? cv[r].y : u == 2 ? cv[r].z : cv[r].w;
#pragma unroll
          for (int c = 0; c < RN; ++c) a[r][c] = fmaf(cval, rv[c], a[r][c]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RM; ++r)
      if (r == rstar) {
#pragma unroll
        for (int c = 0; c < RN; ++c) a[r][c] = Rs[ty * TN + tx + kNB * c];
      }

    if (k + 1 < steps) {
      publish(k + 1, (k + 1) & 1);
      arrive();
    }
  }

#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int c = 0; c < RN; ++c) {
      const int gi = I0 + ty + kNB * r, gj = J0 + tx + kNB * c;
      if (gi < D && gj < D) {
        if (out_h) out_h[gi * ld + gj] = __float2half_rn(a[r][c]);
        if (out_f) out_f[gi * ld + gj] = a[r][c];
      }
    }
}

struct TileChoice {
  int rm, rn;
};
// smallest tile whose grid fits on the 148 SMs (one CTA per SM: the kernel is a chain of grid barriers)
inline bool choose_tile(int D, TileChoice* out) {
  static const TileChoice cand[] = {{2, 2}, {2, 4}, {4, 4}, {4, 8}, {8, 8}};
  for (const TileChoice& t : cand) {
    const int gy = (D + 16 * t.rm - 1) / (16 * t.rm), gx = (D + 16 * t.rn - 1) / (16 * t.rn);
    if (gx * gy <= kNumSMs) {
      *out = t;
      return true;
    }
  }
  return false;
}

template <int RM, int RN>
int launch_inverse(const float* A, int D, float one_minus, float eps, float* rowpanel, float* colpanel,
                   unsigned* counter, __half* out_h, float* out_f, cudaStream_t st) {
  dim3 grid((D + 16 * RN - 1) / (16 * RN), (D + 16 * RM - 1) / (16 * RM)), block(kInvThreads);
  void* args[] = {(void*)&A, (void*)&D, (void*)&one_minus, (void*)&eps, (void*)&rowpanel,
                  (void*)&colpanel, (void*)&counter, (void*)&out_h, (void*)&out_f};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)spd_inverse_kernel<RM, RN>, grid, block, args, 0, st);
  if (e != cudaSuccess) {
    set_error("ua_dota_update_f32: cooperative launch failed: %s", cudaGetErrorString(e));
    return UA_ERR_CUDA;
  }
  return check_launch("ua_dota_update_f32");
}

}  // namespace
}  // namespace ua

extern "C" long long ua_dota_update_workspace_bytes(int D) {
  if (D < 16) return -1;
  // two row panels [16, D] + two column panels [D, 16] + the barrier counter (own 128-byte line)
  return (long long)4 * 16 * D * (long long)sizeof(float) + 128;
}

extern "C" int ua_dota_update_f32(const float* overall, int D, float eps, void* workspace, void* out_lambda_h,
                                  float* out_lambda_f32, void* stream) {
  using namespace ua;
  UA_REQUIRE(overall && workspace && (out_lambda_h || out_lambda_f32), "ua_dota_update_f32: NULL pointer");
  UA_REQUIRE(D >= 16, "ua_dota_update_f32: bad size D=%d", D);
  UA_UNSUPPORTED((D % 16) != 0, "ua_dota_update_f32: D=%d must be a multiple of 16", D);
  TileChoice tc;
  UA_UNSUPPORTED(!choose_tile(D, &tc), "ua_dota_update_f32: D=%d does not fit one register-resident wave (max 1536)", D);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* counter = reinterpret_cast<unsigned*>(workspace);
  float* rowpanel = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 128);
  float* colpanel = rowpanel + (size_t)2 * 16 * D;
  cudaError_t e = cudaMemsetAsync(counter, 0, 128, st);
  if (e != cudaSuccess) {
    set_error("ua_dota_update_f32: memset failed: %s", cudaGetErrorString(e));
    return UA_ERR_CUDA;
  }
  const float one_minus = (float)(1.0 - (double)eps);
  __half* oh = reinterpret_cast<__half*>(out_lambda_h);
  if (tc.rm == 2 && tc.rn == 2) return launch_inverse<2, 2>(overall, D, one_minus, eps, rowpanel, colpanel, counter, oh, out_lambda_f32, st);
  if (tc.rm == 2 && tc.rn == 4) return launch_inverse<2, 4>(overall, D, one_minus, eps, rowpanel, colpanel, counter, oh, out_lambda_f32, st);
  if (tc.rm == 4 && tc.rn == 4) return launch_inverse<4, 4>(overall, D, one_minus, eps, rowpanel, colpanel, counter, oh, out_lambda_f32, st);
  if (tc.rm == 4 && tc.rn == 8) return launch_inverse<4, 8>(overall, D, one_minus, eps, rowpanel, colpanel, counter, oh, out_lambda_f32, st);
  return launch_inverse<8, 8>(overall, D, one_minus, eps, rowpanel, colpanel, counter, oh, out_lambda_f32, st);
}
