// fp32-accurate multi-head self-attention on tcgen05 (3xTF32) for sm_100a: O = softmax(Q K^T / sqrt(d)) V, head dim 64.
//
// The transformer blocks of the point encoder are not part of the adaptation path proper, but after the group encoder
// and the Linear layers moved to the tensor cores, the fp32 SIMT attention torch dispatches (fmha_cutlassF, 182 us per
// layer at 15 x 6 heads x 513 tokens) was a third of the encoder. This kernel keeps fp32-class accuracy with the same
// (hi, lo) operand split as gemm_tf32x3.cu.
//
//   ua_attn_prepare_f32   qkv [B*N, 3*H*64] fp32 -> Q, K as (hi, lo) [B*H, N, 64] (Q pre-scaled by log2(e)/8) and V^T as
//                         (hi, lo) [B*H, 64, Npad] (kv contiguous, zero padded to a multiple of 32)
//   ua_attention_f32      one CTA per (128 query rows, batch*head); kv blocks of 128:
//       warp 0    TMA producer: Q tile once, then K_j (B operand of S = Q K^T) and V^T_j (B operand of O += P V)
//       warp 1    TMEM allocation + one lane issuing tcgen05.mma: S_j = Q K_j^T (A, B from smem) and, once the softmax
//                 warps have published P_j, O += P_j V_j with the A operand read from TENSOR MEMORY (P never touches smem)
//       warps 2-5 one query row per thread: tcgen05.ld of S_j, online softmax in base 2 (running max / sum, one MUFU.EX2 per weight), rescale of
//                 the O accumulator in TMEM, tcgen05.st of P_j as a (hi, lo) tf32 pair; at the end O / l as the (hi, lo)
//                 pair the projection GEMM consumes
//   TMEM columns: S 128 | P_hi 128 | P_lo 128 | O 64.
#include <cuda.h>
#include <string.h>
#include "common.cuh"

namespace ua {
namespace {

constexpr int kHd = 64;          // head dimension
constexpr int kQ = 128;          // query rows per CTA (UMMA M)
constexpr int kKv = 128;         // kv rows per block
constexpr int kAttnThreads = 192;
// 1/sqrt(64) * log2(e): the scores leave the tensor core in base-2 units, so the softmax needs one MUFU.EX2 per weight
constexpr float kQScale = 0.125f * 1.4426950408889634f;

// ---- PTX wrappers (same conventions as gemm_tf32x3.cu) ----------------------------------------------------------
__device__ __forceinline__ uint32_t a_elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ void a_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void a_tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void a_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void a_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void a_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void a_umma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T   (A: lane = row, 8 consecutive 32-bit columns = the K slice)
__device__ __forceinline__ void a_umma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d),
      "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void a_tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// tcgen05.ld without the wait, and a wait that is tied to the destination registers (so that no use of them can be
// scheduled above it): two 32-column loads can be in flight before the first value is consumed.
__device__ __forceinline__ void a_tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void a_tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void a_tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void a_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t a_smem_desc(uint32_t smem_addr) {   // K-major, 128-byte rows, 128B swizzle
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t a_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// 2^x for x <= 0 (MUFU.EX2, 2 ulp; results below 2^-126 flush to zero, which a softmax weight may do)
__device__ __forceinline__ float a_exp2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float a_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---- operand preparation -----------------------------------------------------------------------------------------
// 1-D grid of B*H*ceil(Npad/64) slabs, block 256: a 64-token slab of one head, 16-byte accesses throughout.
// Q (scaled), K: straight (hi, lo) copies; V: transposed through shared memory so that kv becomes the contiguous
// dimension (256-byte runs per head-dimension row).
constexpr int kPrepTokens = 64;

__global__ void __launch_bounds__(256)
    attn_prepare_kernel(const float* __restrict__ qkv, int B, int N, int H, int Npad, float* __restrict__ q_hi,
                        float* __restrict__ q_lo, float* __restrict__ k_hi, float* __restrict__ k_lo,
                        float* __restrict__ vt_hi, float* __restrict__ vt_lo) {
  __shared__ float s_v[kPrepTokens][kHd + 1];
  const int nslab = (Npad + kPrepTokens - 1) / kPrepTokens;
  const int bh = (int)blockIdx.x / nslab, b = bh / H, h = bh - b * H;
  const int n0 = ((int)blockIdx.x - bh * nslab) * kPrepTokens;
  const int C3 = 3 * H * kHd;
  auto split4 = [](const float4 x, float4& hi, float4& lo) {
    hi.x = a_tf32(x.x), hi.y = a_tf32(x.y), hi.z = a_tf32(x.z), hi.w = a_tf32(x.w);
    lo.x = x.x - hi.x, lo.y = x.y - hi.y, lo.z = x.z - hi.z, lo.w = x.w - hi.w;
  };
  for (int e = threadIdx.x; e < kPrepTokens * (kHd / 4); e += 256) {
    const int r = e / (kHd / 4), d = (e - r * (kHd / 4)) * 4, n = n0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
      const float* row = qkv + ((size_t)b * N + n) * C3 + h * kHd + d;
      float4 q = __ldg(reinterpret_cast<const float4*>(row));
      const float4 k = __ldg(reinterpret_cast<const float4*>(row + H * kHd));
      v = __ldg(reinterpret_cast<const float4*>(row + 2 * H * kHd));
      q.x *= kQScale, q.y *= kQScale, q.z *= kQScale, q.w *= kQScale;
      const size_t o = ((size_t)bh * N + n) * kHd + d;
      float4 hi, lo;
      split4(q, hi, lo);
      *reinterpret_cast<float4*>(q_hi + o) = hi, *reinterpret_cast<float4*>(q_lo + o) = lo;
      split4(k, hi, lo);
      *reinterpret_cast<float4*>(k_hi + o) = hi, *reinterpret_cast<float4*>(k_lo + o) = lo;
    }
    s_v[r][d] = v.x, s_v[r][d + 1] = v.y, s_v[r][d + 2] = v.z, s_v[r][d + 3] = v.w;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kHd * (kPrepTokens / 4); e += 256) {
    const int d = e / (kPrepTokens / 4), r = (e - d * (kPrepTokens / 4)) * 4;
    if (n0 + r < Npad) {                     // Npad is a multiple of 32: whole float4 groups are inside or outside
      const float4 v = make_float4(s_v[r][d], s_v[r + 1][d], s_v[r + 2][d], s_v[r + 3][d]);
      float4 hi, lo;
      split4(v, hi, lo);
      const size_t o = ((size_t)bh * kHd + d) * Npad + n0 + r;
      *reinterpret_cast<float4*>(vt_hi + o) = hi, *reinterpret_cast<float4*>(vt_lo + o) = lo;
    }
  }
}

struct AttnParams {
  int N, H, BH;
  float* out_hi;     // [B*N, H*64]
  float* out_lo;
};

struct AttnSmem {
  static constexpr int kTile = kQ * 128;                 // 128 rows x 128 bytes = 16 KB: one k-block of Q or K
  static constexpr int kQBytes = 4 * kTile;              // 2 k-blocks x (hi, lo)
  static constexpr int kKBytes = 4 * kTile;
  static constexpr int kVTile = kHd * 128;               // 64 rows x 128 bytes = 8 KB: 32 kv of V^T
  static constexpr int kVBytes = 8 * kVTile;             // 4 k-blocks x (hi, lo)
  static constexpr int kTotal = kQBytes + kKBytes + kVBytes + 1024 + 256;
};

__global__ void __launch_bounds__(kAttnThreads, 1)
    attention_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                     const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo,
                     const __grid_constant__ CUtensorMap map_v_hi, const __grid_constant__ CUtensorMap map_v_lo,
                     const AttnParams p) {
  using L = AttnSmem;
  extern __shared__ unsigned char s_raw[];
  unsigned char* s_q = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(s_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* s_k = s_q + L::kQBytes;
  unsigned char* s_v = s_k + L::kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_v + L::kVBytes);
  uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 2, *v_full = bars + 3, *v_empty = bars + 4,
           *s_full = bars + 5 /* [2]: one per S buffer */, *p_ready = bars + 7, *o_done = bars + 8;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, q0 = blockIdx.x * kQ;
  const int nkv = (p.N + kKv - 1) / kKv;

  if (warp == 0 && a_elect_one()) {
    mbar_init(q_full, 1), mbar_init(k_full, 1), mbar_init(k_empty, 1), mbar_init(v_full, 1), mbar_init(v_empty, 1);
    mbar_init(s_full, 1), mbar_init(s_full + 1, 1), mbar_init(p_ready, 4), mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  a_fence_before();
  __syncthreads();
  a_fence_after();
  const uint32_t tmem = *s_tmem;
  // S is double-buffered (columns 0-127 / 128-255) and P_hi is written IN PLACE over the S block it came from, so that
  // S_{j+1} = Q K_{j+1}^T runs on the tensor pipe while the softmax warps work on block j (before: S, softmax and PV of a
  // block ran strictly one after the other, tensor pipe 30 % active). P_lo 256-383, O 384-447.
  const uint32_t t_s = tmem, t_plo = tmem + 256, t_o = tmem + 384;

  if (warp == 0) {
    // ===== TMA producers: lane 0 streams Q and K, lane 1 streams V^T =====
    // Two lanes because the two streams wait on different consumers: K_{j+1} may load as soon as S_j has been
    // computed (early: S runs one block ahead of the softmax), V_{j+1} only when PV_j is done. A single producer
    // thread that alternates between them holds K_{j+1} back behind the wait for PV_{j-1}, and the softmax warps then
    // wait for S_{j+1} (22 % of the kernel's stall samples before the split).
    if (lane == 0) {
      const int qrow = bh * p.N + q0;
      mbar_expect_tx(q_full, (uint32_t)L::kQBytes);
      for (int kb = 0; kb < 2; ++kb) {
        a_tma_load_2d(s_q + (2 * kb) * L::kTile, &map_q_hi, q_full, kb * 32, qrow);
        a_tma_load_2d(s_q + (2 * kb + 1) * L::kTile, &map_q_lo, q_full, kb * 32, qrow);
      }
      for (int j = 0; j < nkv; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        mbar_wait(k_empty, ph ^ 1u);
        mbar_expect_tx(k_full, (uint32_t)L::kKBytes);
        const int krow = bh * p.N + j * kKv;
        for (int kb = 0; kb < 2; ++kb) {
          a_tma_load_2d(s_k + (2 * kb) * L::kTile, &map_k_hi, k_full, kb * 32, krow);
          a_tma_load_2d(s_k + (2 * kb + 1) * L::kTile, &map_k_lo, k_full, kb * 32, krow);
        }
      }
    } else if (lane == 1) {
      for (int j = 0; j < nkv; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        mbar_wait(v_empty, ph ^ 1u);
        mbar_expect_tx(v_full, (uint32_t)L::kVBytes);
        for (int kb = 0; kb < 4; ++kb) {
          a_tma_load_2d(s_v + (2 * kb) * L::kVTile, &map_v_hi, v_full, j * kKv + kb * 32, bh * kHd);
          a_tma_load_2d(s_v + (2 * kb + 1) * L::kVTile, &map_v_lo, v_full, j * kKv + kb * 32, bh * kHd);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (a_elect_one()) {
      const uint32_t id_s = a_idesc(kQ, kKv), id_o = a_idesc(kQ, kHd);
      const uint32_t q_base = smem_u32(s_q), k_base = smem_u32(s_k), v_base = smem_u32(s_v);
      mbar_wait(q_full, 0);
      // S_j = Q K_j^T into S buffer j & 1: 2 k-blocks x 4 k-steps x 3 products; a short last block only computes the
      // columns it has. In-order execution of the MMAs of this thread is what makes the buffer reuse safe: S_{j+1}
      // overwrites the buffer of block j-1 and is issued after PV_{j-1}, the last reader of P_hi_{j-1}.
      auto issue_s = [&](int j) {
        mbar_wait(k_full, (uint32_t)j & 1u);
        a_fence_after();
        const int kv_valid = min(kKv, p.N - j * kKv);
        const uint32_t id_sj = kv_valid == kKv ? id_s : a_idesc(kQ, (kv_valid + 15) & ~15);
        const uint32_t t_sj = t_s + (uint32_t)(j & 1) * 128u;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t qh = a_smem_desc(q_base + (2 * kb) * L::kTile), ql = a_smem_desc(q_base + (2 * kb + 1) * L::kTile);
          const uint64_t kh = a_smem_desc(k_base + (2 * kb) * L::kTile), kl = a_smem_desc(k_base + (2 * kb + 1) * L::kTile);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t off = (uint64_t)(ks * 2);
            a_umma_ss(t_sj, ql + off, kh + off, id_sj, (kb | ks) != 0);
            a_umma_ss(t_sj, qh + off, kl + off, id_sj, 1u);
            a_umma_ss(t_sj, qh + off, kh + off, id_sj, 1u);
          }
        }
        a_commit(k_empty);
        a_commit(s_full + (j & 1));
      };
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        if (j + 1 < nkv) issue_s(j + 1);             // runs while the softmax warps are busy with block j
        // O += P_j V_j once P_j is in tensor memory: 4 k-blocks x 4 k-steps x 3 products, A from TMEM
        const int kv_valid = min(kKv, p.N - j * kKv);
        const uint32_t t_phi = t_s + ph * 128u;
        mbar_wait(v_full, ph);
        mbar_wait(p_ready, ph);
        a_fence_after();
        const int ksteps = (kv_valid + 7) >> 3;      // 8 kv per MMA; a short last block stops early
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const uint64_t vh = a_smem_desc(v_base + (2 * kb) * L::kVTile), vl = a_smem_desc(v_base + (2 * kb + 1) * L::kVTile);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (kb * 4 + ks < ksteps) {
              const uint64_t off = (uint64_t)(ks * 2);
              const uint32_t col = (uint32_t)(kb * 32 + ks * 8);
              a_umma_ts(t_o, t_plo + col, vh + off, id_o, (uint32_t)((j | kb | ks) != 0));
              a_umma_ts(t_o, t_phi + col, vl + off, id_o, 1u);
              a_umma_ts(t_o, t_phi + col, vh + off, id_o, 1u);
            }
          }
        }
        a_commit(v_empty);
        a_commit(o_done);
      }
    }
  } else {
    // ===== softmax / epilogue warps: one query row per thread =====
    const int quarter = warp & 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int qi = q0 + quarter * 32 + lane;           // token index of this row
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const uint32_t ph = (uint32_t)j & 1u;
      const int kv_valid = min(kKv, p.N - j * kKv);    // columns beyond the sequence are masked
      const uint32_t t_sj = t_s + ph * 128u;           // S_j, and P_hi_j in place once pass 2 has rewritten it
      mbar_wait(s_full + (j & 1), (uint32_t)(j >> 1) & 1u);
      a_fence_after();
      // pass 1: row maximum of the block (columns beyond the sequence count as -inf; branch-free so that the 32
      // per-column chains of a thread interleave)
      float bmax = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < kKv; c0 += 64) {           // two 32-column loads in flight per round trip
        if (c0 >= kv_valid) break;
        uint32_t ra[32], rb[32];
        const bool two = c0 + 32 < kv_valid;
        a_tmem_ld32_issue(t_sj + lane_off + (uint32_t)c0, ra);
        if (two) a_tmem_ld32_issue(t_sj + lane_off + (uint32_t)(c0 + 32), rb);
        a_tmem_ld_wait(ra);
        const int na = kv_valid - c0;
        if (na >= 32) {                                // full chunk (warp-uniform): no masking selects
#pragma unroll
          for (int i = 0; i < 32; ++i) bmax = fmaxf(bmax, __uint_as_float(ra[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) bmax = fmaxf(bmax, i < na ? __uint_as_float(ra[i]) : -INFINITY);
        }
        if (two) {
          a_tmem_ld_wait(rb);
          const int nb = na - 32;
          if (nb >= 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) bmax = fmaxf(bmax, __uint_as_float(rb[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) bmax = fmaxf(bmax, i < nb ? __uint_as_float(rb[i]) : -INFINITY);
          }
        }
      }
      const float m_new = fmaxf(m_run, bmax);
      const float alpha = a_exp2(m_run - m_new);       // 0 at the first block (m_run = -inf)
      // pass 2: P = 2^(S - m_new) as a (hi, lo) pair into tensor memory, running sum. P_hi goes in place over S_j (a
      // buffer PV_{j-1} does not touch); P_lo and the O accumulator are still being read / written by PV_{j-1}, so the
      // first 64 columns are exponentiated and their P_hi stored BEFORE the wait for PV_{j-1}, their P_lo kept in
      // registers until it is over (that wait was 9 % of the kernel's stall samples).
      float psum = 0.f;
      auto weights = [&](uint32_t (&r)[32], int c0, float (&lo)[32]) {   // one 32-column chunk; P_hi stored, P_lo returned
        float hi[32];
        const int nv = kv_valid - c0;
        // (hi, lo) split of a weight by TRUNCATION (one LOP3 instead of a cvt.rna on the conversion pipe, which the
        // MUFU.EX2 of the same weight already loads): hi keeps 10 mantissa bits, lo = e - hi is exact (13 bits); the
        // tensor core truncates lo to tf32 itself, leaving 2^-21 relative per weight instead of 2^-22
        if (nv >= 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = a_exp2(__uint_as_float(r[i]) - m_new);
            psum += e;
            hi[i] = __uint_as_float(__float_as_uint(e) & 0xffffe000u);
            lo[i] = e - hi[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = a_exp2((i < nv ? __uint_as_float(r[i]) : -INFINITY) - m_new);
            psum += e;
            hi[i] = __uint_as_float(__float_as_uint(e) & 0xffffe000u);
            lo[i] = e - hi[i];
          }
        }
        a_tmem_st32(t_sj + lane_off + (uint32_t)c0, hi);     // in place: this thread's row, columns it has just read
      };
      float lo0[32], lo1[32];
      const bool has1 = 32 < kv_valid;
      {
        uint32_t ra[32], rb[32];
        a_tmem_ld32_issue(t_sj + lane_off, ra);
        if (has1) a_tmem_ld32_issue(t_sj + lane_off + 32u, rb);
        a_tmem_ld_wait(ra);
        weights(ra, 0, lo0);
        if (has1) {
          a_tmem_ld_wait(rb);
          weights(rb, 32, lo1);
        }
      }
      // the previous PV must have finished before O is rescaled and before P_lo is overwritten
      if (j > 0) {
        mbar_wait(o_done, ph ^ 1u);
        a_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < kHd; c0 += 32) {
          float o[32];
          a_tmem_ld32(t_o + lane_off + (uint32_t)c0, o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] *= alpha;
          a_tmem_st32(t_o + lane_off + (uint32_t)c0, o);
        }
      }
      a_tmem_st32(t_plo + lane_off, lo0);
      if (has1) a_tmem_st32(t_plo + lane_off + 32u, lo1);
#pragma unroll 1
      for (int c0 = 64; c0 < kKv; c0 += 64) {
        if (c0 >= kv_valid) break;                     // the PV MMAs stop at ceil(kv_valid / 8) * 8 columns
        uint32_t ra[32], rb[32];
        const bool two = c0 + 32 < kv_valid;
        a_tmem_ld32_issue(t_sj + lane_off + (uint32_t)c0, ra);
        if (two) a_tmem_ld32_issue(t_sj + lane_off + (uint32_t)(c0 + 32), rb);
        a_tmem_ld_wait(ra);
        weights(ra, c0, lo0);
        a_tmem_st32(t_plo + lane_off + (uint32_t)c0, lo0);
        if (two) {
          a_tmem_ld_wait(rb);
          weights(rb, c0 + 32, lo1);
          a_tmem_st32(t_plo + lane_off + (uint32_t)(c0 + 32), lo1);
        }
      }
      a_tmem_st_wait();
      l_run = l_run * alpha + psum;
      m_run = m_new;
      a_fence_before();
      __syncwarp();
      if (lane == 0) a_mbar_arrive(p_ready);     // P_j (and the rescaled O) are in tensor memory
    }
    // epilogue: O / l as the (hi, lo) pair of the projection GEMM's A operand
    mbar_wait(o_done, (uint32_t)(nkv - 1) & 1u);
    a_fence_after();
    const float inv_l = __fdiv_rn(1.0f, l_run);
    const int b = bh / p.H, h = bh - b * p.H;
#pragma unroll 1
    for (int c0 = 0; c0 < kHd; c0 += 32) {
      float o[32];
      a_tmem_ld32(t_o + lane_off + (uint32_t)c0, o);
      if (qi < p.N) {
        const size_t off = ((size_t)b * p.N + qi) * (size_t)(p.H * kHd) + h * kHd + c0;
        float4* oh = reinterpret_cast<float4*>(p.out_hi + off);
        float4* ol = reinterpret_cast<float4*>(p.out_lo + off);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 hv, lv;
          float x;
          x = o[4 * i] * inv_l, hv.x = a_tf32(x), lv.x = x - hv.x;
          x = o[4 * i + 1] * inv_l, hv.y = a_tf32(x), lv.y = x - hv.y;
          x = o[4 * i + 2] * inv_l, hv.z = a_tf32(x), lv.z = x - hv.z;
          x = o[4 * i + 3] * inv_l, hv.w = a_tf32(x), lv.w = x - hv.w;
          oh[i] = hv, ol[i] = lv;
        }
      }
    }
  }
  a_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int attn_map(CUtensorMap* map, const float* base, long long rows, long long cols, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  if (!fn) {
    set_error("ua_attention_f32: cuTensorMapEncodeTiled is not available from the driver");
    return UA_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("ua_attention_f32: cuTensorMapEncodeTiled failed with %d", (int)r);
    return UA_ERR_CUDA;
  }
  return UA_OK;
}

}  // namespace
}  // namespace ua

extern "C" long long ua_attn_padded_tokens(int N) { return ((long long)N + 31) / 32 * 32; }

extern "C" int ua_attn_prepare_f32(const float* qkv, int B, int N, int H, float* q_hi, float* q_lo, float* k_hi,
                                   float* k_lo, float* vt_hi, float* vt_lo, void* stream) {
  using namespace ua;
  UA_REQUIRE(qkv && q_hi && q_lo && k_hi && k_lo && vt_hi && vt_lo, "ua_attn_prepare_f32: NULL pointer");
  UA_REQUIRE(B >= 1 && N >= 1 && H >= 1 && (long long)B * H <= 65535, "ua_attn_prepare_f32: bad sizes B=%d N=%d H=%d", B,
             N, H);
  const int Npad = (int)ua_attn_padded_tokens(N);
  UA_REQUIRE(((uintptr_t)qkv | (uintptr_t)q_hi | (uintptr_t)q_lo | (uintptr_t)k_hi | (uintptr_t)k_lo | (uintptr_t)vt_hi |
              (uintptr_t)vt_lo) % 16 == 0, "ua_attn_prepare_f32: pointers must be 16-byte aligned");
  const unsigned grid = (unsigned)((long long)B * H * ((Npad + kPrepTokens - 1) / kPrepTokens));
  attn_prepare_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(qkv, B, N, H, Npad, q_hi, q_lo, k_hi, k_lo, vt_hi, vt_lo);
  return check_launch("ua_attn_prepare_f32");
}

extern "C" int ua_attention_f32(const float* q_hi, const float* q_lo, const float* k_hi, const float* k_lo,
                                const float* vt_hi, const float* vt_lo, int B, int N, int H, float* out_hi,
                                float* out_lo, void* stream) {
  using namespace ua;
  UA_REQUIRE(q_hi && q_lo && k_hi && k_lo && vt_hi && vt_lo && out_hi && out_lo, "ua_attention_f32: NULL pointer");
  UA_REQUIRE(B >= 1 && N >= 1 && H >= 1 && (long long)B * H <= 65535, "ua_attention_f32: bad sizes B=%d N=%d H=%d", B, N, H);
  UA_REQUIRE((uintptr_t)out_hi % 16 == 0 && (uintptr_t)out_lo % 16 == 0, "ua_attention_f32: outputs must be 16-byte aligned");
  const int BH = B * H;
  const long long Npad = ua_attn_padded_tokens(N);
  CUtensorMap mq_hi, mq_lo, mk_hi, mk_lo, mv_hi, mv_lo;
  int rc;
  if ((rc = attn_map(&mq_hi, q_hi, (long long)BH * N, kHd, kQ)) != UA_OK) return rc;
  if ((rc = attn_map(&mq_lo, q_lo, (long long)BH * N, kHd, kQ)) != UA_OK) return rc;
  if ((rc = attn_map(&mk_hi, k_hi, (long long)BH * N, kHd, kKv)) != UA_OK) return rc;
  if ((rc = attn_map(&mk_lo, k_lo, (long long)BH * N, kHd, kKv)) != UA_OK) return rc;
  if ((rc = attn_map(&mv_hi, vt_hi, (long long)BH * kHd, Npad, kHd)) != UA_OK) return rc;
  if ((rc = attn_map(&mv_lo, vt_lo, (long long)BH * kHd, Npad, kHd)) != UA_OK) return rc;
  AttnParams p;
  p.N = N, p.H = H, p.BH = BH, p.out_hi = out_hi, p.out_lo = out_lo;
  cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal);
  if (e != cudaSuccess) {
    set_error("ua_attention_f32: cudaFuncSetAttribute(%d B): %s", AttnSmem::kTotal, cudaGetErrorString(e));
    return UA_ERR_CUDA;
  }
  // (A second softmax warpgroup -- 8 softmax warps, each thread one row and half of the columns, row maxima exchanged
  // through shared memory -- was built and measured after the S double buffer: 123.5 us, no change. The per-block cycle
  // is p_ready -> PV (tensor pipe) -> o_done -> rescale + pass 2 -> p_ready, not the exponentials.)
  // One CTA per 128 query rows. (Handing the one leftover row of 512 + 1 tokens to a SIMT side kernel, so that every
  // (batch, head) needs 4 CTAs instead of 5, was tried: the latency-bound side kernel cost what the saved wave gained.
  // So was a SIMT branch for the leftover row INSIDE this kernel: one row still has to read the (batch, head)'s whole K
  // and V^T, 540 KB through one SM's L2 port = 14 us at best against 26 us for the tensor-core CTA, and the plain
  // version measured 100 us per tail CTA: attention 159 -> 250 us per launch. Removed.)
  const int q_tiles = (N + kQ - 1) / kQ;
  dim3 grid(q_tiles, BH);
  attention_kernel<<<grid, kAttnThreads, AttnSmem::kTotal, (cudaStream_t)stream>>>(mq_hi, mq_lo, mk_hi, mk_lo, mv_hi,
                                                                                   mv_lo, p);
  return check_launch("ua_attention_f32");
}
