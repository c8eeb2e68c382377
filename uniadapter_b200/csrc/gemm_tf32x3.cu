// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05, kind::tf32) for sm_100a.
//
//   C[M,N] = epilogue(A[M,K] . W[N,K]^T)          A, W row-major with K contiguous ("K-major", nn.Linear / 1x1-conv layout)
//
// fp32 accuracy comes from the 3xTF32 split: every operand is stored as a pair (hi, lo) with hi = tf32(x) (low 13
// mantissa bits zero, so the tensor core reads it exactly) and lo = x - hi (exact in fp32); the kernel accumulates
//   A_lo.W_hi + A_hi.W_lo + A_hi.W_hi   in the fp32 TMEM accumulator (the dropped lo.lo term is 2^-22 relative).
// Producers write their outputs directly as (hi, lo) pairs (OUT_SPLIT epilogue), so a chain of GEMMs never runs a
// separate split pass.
//
// Structure (one CTA per SM, persistent over 128 x BN output tiles, warp-specialised):
//   warp 0     TMA producer: cp.async.bulk.tensor (128B swizzle) of A_hi, A_lo [128 x 32] and W_hi, W_lo [BN x 32]
//              per k-block into a multi-stage shared-memory ring (full/empty mbarriers)
//   warp 1     TMEM allocation; one elected lane issues tcgen05.mma (M=128, N=BN, K=8 per instruction, 12 per
//              k-block), tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2-5  epilogue: tcgen05.ld of the accumulator (one row per thread, 32 columns per load), bias / per-row-group
//              bias / residual / ReLU / GELU, then any of: fp32 tile, (hi, lo) tiles — staged in 128B-swizzled shared
//              memory and written by TMA stores (full lines instead of 32 partial lines per instruction) — and the max
//              over each group of 32 consecutive rows (= one warp's TMEM lanes: the 32 points of a point group)
//   two accumulator stages in TMEM (2 x BN <= 512 columns): the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <string.h>
#include "common.cuh"

namespace ua {

int g_gemm_bn = 0;   // tuning override: force the N tile (128, 192 or 256)

namespace {

constexpr int kBM = 128;         // rows per tile (UMMA M)
constexpr int kBK = 32;          // fp32 elements per k-block = 128 bytes = one swizzle-128B row
constexpr int kUmmaK = 8;        // tf32 elements per tcgen05.mma
constexpr int kGemmThreads = 192;

struct GemmParams {
  int M, N, K;
  const float* bias;        // [N] or null
  const float* group_bias;  // [ceil(M/32), N] or null: added to every row of a 32-row group
  int act;                  // 0 none, 1 ReLU, 2 GELU (erf form, torch.nn.functional.gelu default)
  const float* residual;    // [M, ldo] or null: added after the bias (transformer skip connection)
  float* out;               // [M, ldo] fp32 or null
  float* out_hi;            // [M, ldo] or null (with out_lo)
  float* out_lo;
  long long ldo;
  float* gmax;              // [M/32, N] fp32 or null: max over each group of 32 consecutive rows
  float* gmax_hi;           // optional (hi, lo) copy of gmax
  float* gmax_lo;
};

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t elect_one() {
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// shared -> global tile store through the TMA unit (bulk async-group completion); rows beyond the tensor are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the mbarrier receives one arrival when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
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

// Shared-memory matrix descriptor of a K-major tile whose rows are 128 bytes, 128B-swizzled (what the TMA box
// [32 fp32 x rows] with CU_TENSOR_MAP_SWIZZLE_128B writes): 8-row atoms of 1024 bytes (stride byte offset),
// descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B). Addresses and offsets are in units of 16 bytes.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);       // start address        bits [0,14)
  d |= (uint64_t)1 << 16;                            // leading byte offset  bits [16,30) (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset   bits [32,46)
  d |= (uint64_t)1 << 46;                            // version              bits [46,48)
  d |= (uint64_t)2 << 61;                            // layout type          bits [61,64)
  return d;
}
// Instruction descriptor: D fp32, A/B tf32, both K-major, dense; N >> 3 at [17,23), M >> 4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// order-preserving float <-> int map for REDUX max
__device__ __forceinline__ int float_to_sortable(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float sortable_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = kBM * kBK * 4;       // 16 KB
  static constexpr int kBBytes = BN * kBK * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kStoreBytes = 4 * 4096;        // one 32 x 32 fp32 staging tile per epilogue warp (TMA store)
  static constexpr int kBiasBytes = 4 * BN * 4;       // one effective bias row per epilogue warp
  static constexpr int kTotal =
      STAGES * kStageBytes + kStoreBytes + kBiasBytes + 1024 /* alignment slack */ + 256 /* barriers */;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
    gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                       const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                       const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_out_hi,
                       const __grid_constant__ CUtensorMap map_out_lo, const GemmParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ unsigned char s_raw[];
  // 128B-swizzled tiles need 1024-byte alignment
  unsigned char* s_tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(s_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* s_store = s_tiles + STAGES * L::kStageBytes;                             // [4 warps][4096], 1024-aligned
  float* s_bias = reinterpret_cast<float*>(s_store + L::kStoreBytes);                     // [4 warps][BN]
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_store + L::kStoreBytes + L::kBiasBytes);   // [STAGES]
  uint64_t* s_empty = s_full + STAGES;                                                   // [STAGES]
  uint64_t* s_tfull = s_empty + STAGES;                                                  // [2]
  uint64_t* s_tempty = s_tfull + 2;                                                      // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + kBM - 1) / kBM, n_tiles = p.N / BN, k_blocks = p.K / kBK;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&map_a_hi), prefetch_tmap(&map_a_lo), prefetch_tmap(&map_w_hi), prefetch_tmap(&map_w_lo);
    if (p.out) prefetch_tmap(&map_out);
    if (p.out_hi) prefetch_tmap(&map_out_hi), prefetch_tmap(&map_out_lo);
    for (int i = 0; i < STAGES; ++i) mbar_init(&s_full[i], 1), mbar_init(&s_empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&s_tfull[i], 1), mbar_init(&s_tempty[i], 4);
    fence_mbar_init();
  }
  constexpr uint32_t kTmemCols = 2 * BN <= 256 ? 256u : 512u;   // 2 accumulator stages of BN fp32 columns, power of two
  if (warp == 1) tmem_alloc(s_tmem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m_idx = (t / n_tiles) * kBM, n_idx = (t % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&s_empty[stage], phase ^ 1u);
          unsigned char* st = s_tiles + stage * L::kStageBytes;
          mbar_expect_tx(&s_full[stage], (uint32_t)L::kStageBytes);
          tma_load_2d(st, &map_a_hi, &s_full[stage], kb * kBK, m_idx);
          tma_load_2d(st + L::kABytes, &map_a_lo, &s_full[stage], kb * kBK, m_idx);
          tma_load_2d(st + 2 * L::kABytes, &map_w_hi, &s_full[stage], kb * kBK, n_idx);
          tma_load_2d(st + 2 * L::kABytes + L::kBBytes, &map_w_lo, &s_full[stage], kb * kBK, n_idx);
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = make_idesc(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&s_tempty[acc], acc_phase ^ 1u);       // the epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&s_full[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(s_tiles + stage * L::kStageBytes);
          const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + L::kABytes);
          const uint64_t w_hi = make_smem_desc(st + 2 * L::kABytes);
          const uint64_t w_lo = make_smem_desc(st + 2 * L::kABytes + L::kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint64_t koff = (uint64_t)((k * kUmmaK * 4) >> 4);   // 32 bytes per step inside the swizzled row
            umma_tf32(d_tmem, a_lo + koff, w_hi + koff, idesc, (kb | k) != 0);
            umma_tf32(d_tmem, a_hi + koff, w_lo + koff, idesc, 1u);
            umma_tf32(d_tmem, a_hi + koff, w_hi + koff, idesc, 1u);
          }
          umma_commit(&s_empty[stage]);                  // frees the smem stage once these MMAs have read it
          if (kb == k_blocks - 1) umma_commit(&s_tfull[acc]);
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
        if (++acc == 2) acc = 0, acc_phase ^= 1u;
      }
    }
  } else {
    // ===== epilogue warps (2..5): TMEM lanes 32*(warp % 4) .. +31 =====
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int m_idx = (t / n_tiles) * kBM, n_idx = (t % n_tiles) * BN;
      const int row = m_idx + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const int group = (m_idx + quarter * 32) >> 5;       // 32-row group of this warp
      // effective bias row of this warp (bias[N] + group_bias[group, N]) into shared memory while the MMAs of the
      // tile are still running: the chunk loop below then reads it with broadcast loads instead of stalling on HBM
      const bool has_bias = p.bias != nullptr || p.group_bias != nullptr;
      float* my_bias = s_bias + (warp - 2) * BN;
      unsigned char* my_store = s_store + (warp - 2) * 4096;
      if (has_bias) {
        const bool grp_ok = p.group_bias != nullptr && m_idx + quarter * 32 < p.M;
#pragma unroll
        for (int c = lane * 4; c < BN; c += 128) {
          float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n_idx + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (grp_ok) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.group_bias + (size_t)group * p.N + n_idx + c));
            b.x += g.x, b.y += g.y, b.z += g.z, b.w += g.w;
          }
          *reinterpret_cast<float4*>(my_bias + c) = b;
        }
        __syncwarp();
      }
      mbar_wait(&s_tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld_32x32(t_row + (uint32_t)c0, v);
        const int col = n_idx + c0;
        if (has_bias) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(my_bias + c0 + i);
            v[i] += b.x, v[i + 1] += b.y, v[i + 2] += b.z, v[i + 3] += b.w;
          }
        }
        if (p.residual && row_ok) {
          const float4* rp = reinterpret_cast<const float4*>(p.residual + (size_t)row * p.ldo + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 r4 = __ldg(rp + i);
            v[4 * i] += r4.x, v[4 * i + 1] += r4.y, v[4 * i + 2] += r4.z, v[4 * i + 3] += r4.w;
          }
        }
        if (p.act == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        } else if (p.act == 2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.5f * v[i] * (1.0f + erff(v[i] * 0.70710678118654752f));
        }
        // outputs leave through shared memory and the TMA unit: a thread owns one row, so direct stores would touch 32
        // different lines per instruction; the staged tile (128B-swizzled, conflict-free quarter-warp writes) goes out
        // as full lines. One 4 KB staging tile per warp, reused once the previous store has read it.
        const int row0 = m_idx + quarter * 32;
        auto stage_and_store = [&](const CUtensorMap* map, const float (&t)[32]) {
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
          const uint32_t base = smem_u32(my_store) + (uint32_t)lane * 128u;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t addr = base + ((uint32_t)(c ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(t[4 * c]), "f"(t[4 * c + 1]),
                         "f"(t[4 * c + 2]), "f"(t[4 * c + 3])
                         : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(map, my_store, col, row0);
            bulk_commit();
          }
        };
        if (p.out && row0 < p.M) stage_and_store(&map_out, v);
        if (p.out_hi && row0 < p.M) {
          float h[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) h[i] = tf32_round(v[i]);
          stage_and_store(&map_out_hi, h);
#pragma unroll
          for (int i = 0; i < 32; ++i) h[i] = v[i] - h[i];
          stage_and_store(&map_out_lo, h);
        }
        if (p.gmax) {
          // max over the warp's 32 rows (one point group), one REDUX per column; lane i keeps column i
          float mine = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int m = __reduce_max_sync(kFullMask, float_to_sortable(row_ok ? v[i] : -INFINITY));
            if (lane == i) mine = sortable_to_float(m);
          }
          if (m_idx + quarter * 32 < p.M) {
            const size_t o = (size_t)group * p.N + col + lane;
            p.gmax[o] = mine;
            if (p.gmax_hi) {
              const float h = tf32_round(mine);
              p.gmax_hi[o] = h, p.gmax_lo[o] = mine - h;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();                                        // also orders this tile's bias reads before the next staging
      if (lane == 0) mbar_arrive(&s_tempty[acc]);
      if (++acc == 2) acc = 0, acc_phase ^= 1u;
    }
    if (lane == 0) bulk_wait<0>();     // this warp's tile stores are complete before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// hi = tf32(x) (round to nearest, low 13 bits zero), lo = x - hi
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                         float* __restrict__ lo, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    float4 h, l;
    h.x = tf32_round(v.x), l.x = v.x - h.x;
    h.y = tf32_round(v.y), l.y = v.y - h.y;
    h.z = tf32_round(v.z), l.z = v.z - h.z;
    h.w = tf32_round(v.w), l.w = v.w - h.w;
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
}

// Point-wise first layer of the group encoder: y = relu?(x . W^T + b) with a tiny inner dimension (C = 3 or 6 input
// channels), written as the (hi, lo) pair the next GEMM consumes. One thread per (row, 4 output channels).
template <int C>
__global__ void __launch_bounds__(256) pointwise_linear_split_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ b, int relu, long long M,
                                                                   int N, float* __restrict__ hi, float* __restrict__ lo) {
  // A thread keeps its 4 output channels (weights + bias in registers) and walks rows: per row three broadcast loads,
  // 4*C FMAs and two 16-byte stores, so the kernel is the HBM write stream of the (hi, lo) pair (8*N bytes per row).
  const int n4 = N / 4;
  const long long tglobal = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long tstride = (long long)gridDim.x * blockDim.x;     // a multiple of n4 (host guarantees it)
  const int c0 = (int)(tglobal % n4) * 4;
  float wr[4][C], br[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    br[j] = __ldg(b + c0 + j);
#pragma unroll
    for (int c = 0; c < C; ++c) wr[j][c] = __ldg(w + (size_t)(c0 + j) * C + c);
  }
  const long long row_stride = tstride / n4;
#pragma unroll 2
  for (long long row = tglobal / n4; row < M; row += row_stride) {
    float xin[C];
#pragma unroll
    for (int c = 0; c < C; ++c) xin[c] = __ldg(x + row * C + c);
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) acc = fmaf(xin[c], wr[j][c], acc);
      acc += br[j];
      o[j] = relu ? fmaxf(acc, 0.f) : acc;
    }
    float4 h, l;
    h.x = tf32_round(o[0]), l.x = o[0] - h.x;
    h.y = tf32_round(o[1]), l.y = o[1] - h.y;
    h.z = tf32_round(o[2]), l.z = o[2] - h.z;
    h.w = tf32_round(o[3]), l.w = o[3] - h.w;
    __stcs(reinterpret_cast<float4*>(hi + row * N + c0), h);
    __stcs(reinterpret_cast<float4*>(lo + row * N + c0), l);
  }
}

// LayerNorm over the last dimension fused with the optional position-embedding add in front of it and the (hi, lo)
// split behind it: s = x (+ pos); y = (s - mean) * rsqrt(var + eps) * gamma + beta. One warp per row, the row in
// registers (C <= 32 * 4 * kLnVec), two-pass variance. out_sum (the skip-connection input) is optional.
constexpr int kLnVec = 8;   // float4 per lane: rows up to 1024 floats
__global__ void __launch_bounds__(256) layernorm_split_kernel(const float* __restrict__ x, const float* __restrict__ pos,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, long long rows, int C, float* __restrict__ out_sum,
                                                              float* __restrict__ out_hi, float* __restrict__ out_lo) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = C / 128;           // float4 per lane (C % 128 == 0)
  float4 v[kLnVec];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kLnVec; ++j) {
    if (j < nvec) {
      const size_t o = (size_t)row * C + j * 128 + lane * 4;
      v[j] = __ldg(reinterpret_cast<const float4*>(x + o));
      if (pos) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(pos + o));
        v[j].x += q.x, v[j].y += q.y, v[j].z += q.z, v[j].w += q.w;
      }
      if (out_sum) *reinterpret_cast<float4*>(out_sum + o) = v[j];
      sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  }
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < kLnVec; ++j) {
    if (j < nvec) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
#pragma unroll
  for (int j = 0; j < kLnVec; ++j) {
    if (j < nvec) {
      const int c0 = j * 128 + lane * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c0));
      float y[4] = {(v[j].x - mean) * rstd * g.x + b.x, (v[j].y - mean) * rstd * g.y + b.y,
                    (v[j].z - mean) * rstd * g.z + b.z, (v[j].w - mean) * rstd * g.w + b.w};
      float4 h, l;
      h.x = tf32_round(y[0]), l.x = y[0] - h.x;
      h.y = tf32_round(y[1]), l.y = y[1] - h.y;
      h.z = tf32_round(y[2]), l.z = y[2] - h.z;
      h.w = tf32_round(y[3]), l.w = y[3] - h.w;
      const size_t o = (size_t)row * C + c0;
      *reinterpret_cast<float4*>(out_hi + o) = h;
      *reinterpret_cast<float4*>(out_lo + o) = l;
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row pitch ld (elements), box [box_rows x 32 columns], 128B swizzle, zero fill
int make_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows) {
  memset(map, 0, sizeof(*map));
  if (!base) return UA_OK;             // unused output: the kernel never touches this map
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("ua_gemm_tf32x3_f32: cuTensorMapEncodeTiled is not available from the driver");
    return UA_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("ua_gemm_tf32x3_f32: cuTensorMapEncodeTiled failed with %d (base %p rows %lld cols %lld ld %lld)", (int)r,
              (const void*)base, rows, cols, ld);
    return UA_ERR_CUDA;
  }
  return UA_OK;
}

template <int BN, int STAGES>
int launch_gemm(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo,
                const CUtensorMap& o, const CUtensorMap& o_hi, const CUtensorMap& o_lo, const GemmParams& p,
                cudaStream_t st) {
  using L = SmemLayout<BN, STAGES>;
  auto kern = gemm_tf32x3_kernel<BN, STAGES>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
  if (e != cudaSuccess) {
    set_error("ua_gemm_tf32x3_f32: cudaFuncSetAttribute(%d B): %s", L::kTotal, cudaGetErrorString(e));
    return UA_ERR_CUDA;
  }
  const int tiles = ((p.M + kBM - 1) / kBM) * (p.N / BN);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  kern<<<grid, kGemmThreads, L::kTotal, st>>>(a_hi, a_lo, w_hi, w_lo, o, o_hi, o_lo, p);
  return check_launch("ua_gemm_tf32x3_f32");
}

}  // namespace
}  // namespace ua

extern "C" int ua_split_tf32_f32(const float* x, float* hi, float* lo, long long n, void* stream) {
  using namespace ua;
  UA_REQUIRE(x && hi && lo && n >= 0, "ua_split_tf32_f32: NULL pointer");
  UA_REQUIRE(n % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)hi % 16 == 0 && (uintptr_t)lo % 16 == 0,
             "ua_split_tf32_f32: n must be a multiple of 4 and the pointers 16-byte aligned");
  if (n == 0) return UA_OK;
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n4);
  return check_launch("ua_split_tf32_f32");
}

extern "C" int ua_layernorm_split_f32(const float* x, const float* pos, const float* gamma, const float* beta, float eps,
                                     long long rows, int C, float* out_sum, float* out_hi, float* out_lo, void* stream) {
  using namespace ua;
  UA_REQUIRE(x && gamma && beta && out_hi && out_lo && rows >= 1, "ua_layernorm_split_f32: NULL pointer / no rows");
  UA_UNSUPPORTED(C % 128 != 0 || C > 128 * kLnVec, "ua_layernorm_split_f32: C=%d must be a multiple of 128, <= %d", C,
                 128 * kLnVec);
  UA_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)out_hi % 16 == 0 && (uintptr_t)out_lo % 16 == 0 &&
                 (uintptr_t)gamma % 16 == 0 && (uintptr_t)beta % 16 == 0 && (!pos || (uintptr_t)pos % 16 == 0) &&
                 (!out_sum || (uintptr_t)out_sum % 16 == 0),
             "ua_layernorm_split_f32: pointers must be 16-byte aligned");
  const long long blocks = (rows + 7) / 8;
  layernorm_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, pos, gamma, beta, eps, rows, C, out_sum,
                                                                           out_hi, out_lo);
  return check_launch("ua_layernorm_split_f32");
}

extern "C" int ua_pointwise_linear_split_f32(const float* x, const float* w, const float* b, int relu, long long M, int C,
                                            int N, float* out_hi, float* out_lo, void* stream) {
  using namespace ua;
  UA_REQUIRE(x && w && b && out_hi && out_lo, "ua_pointwise_linear_split_f32: NULL pointer");
  UA_REQUIRE(M >= 1 && N >= 4 && N % 4 == 0, "ua_pointwise_linear_split_f32: bad sizes M=%lld N=%d", M, N);
  UA_REQUIRE((uintptr_t)out_hi % 16 == 0 && (uintptr_t)out_lo % 16 == 0, "ua_pointwise_linear_split_f32: unaligned output");
  UA_UNSUPPORTED(C != 3 && C != 6, "ua_pointwise_linear_split_f32: C=%d input channels (3 or 6 supported)", C);
  const long long total = M * (N / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  // the kernel keeps a thread on fixed output channels: the grid size in threads must be a multiple of N/4
  UA_UNSUPPORTED(256 % (N / 4) != 0 && (N / 4) % 256 != 0, "ua_pointwise_linear_split_f32: N=%d (N/4 must divide 256 or be a multiple of it)", N);
  if ((N / 4) > 256) blocks = (blocks + (N / 4) / 256 - 1) / ((N / 4) / 256) * ((N / 4) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 3)
    pointwise_linear_split_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(x, w, b, relu, M, N, out_hi, out_lo);
  else
    pointwise_linear_split_kernel<6><<<(unsigned)blocks, 256, 0, st>>>(x, w, b, relu, M, N, out_hi, out_lo);
  return check_launch("ua_pointwise_linear_split_f32");
}

extern "C" int ua_gemm_tf32x3_f32(const float* a_hi, const float* a_lo, long long lda, const float* w_hi,
                                  const float* w_lo, long long ldw, int M, int N, int K, const float* bias,
                                  const float* group_bias, const float* residual, int act, float* out, float* out_hi,
                                  float* out_lo, long long ldo, float* gmax, float* gmax_hi, float* gmax_lo,
                                  void* stream) {
  using namespace ua;
  UA_REQUIRE(a_hi && a_lo && w_hi && w_lo, "ua_gemm_tf32x3_f32: NULL operand");
  UA_REQUIRE(M >= 1 && N >= 1 && K >= 1, "ua_gemm_tf32x3_f32: bad sizes M=%d N=%d K=%d", M, N, K);
  UA_UNSUPPORTED(N % 128 != 0, "ua_gemm_tf32x3_f32: N=%d must be a multiple of 128", N);
  UA_UNSUPPORTED(K % kBK != 0, "ua_gemm_tf32x3_f32: K=%d must be a multiple of %d", K, kBK);
  UA_REQUIRE(lda >= K && ldw >= K && lda % 4 == 0 && ldw % 4 == 0, "ua_gemm_tf32x3_f32: bad leading dimensions");
  UA_REQUIRE((uintptr_t)a_hi % 16 == 0 && (uintptr_t)a_lo % 16 == 0 && (uintptr_t)w_hi % 16 == 0 &&
                 (uintptr_t)w_lo % 16 == 0,
             "ua_gemm_tf32x3_f32: operands must be 16-byte aligned");
  UA_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "ua_gemm_tf32x3_f32: out_hi and out_lo come as a pair");
  UA_REQUIRE((gmax_hi == nullptr) == (gmax_lo == nullptr) && (!gmax_hi || gmax),
             "ua_gemm_tf32x3_f32: gmax_hi / gmax_lo come as a pair, with gmax");
  UA_REQUIRE(out || out_hi || gmax, "ua_gemm_tf32x3_f32: no output requested");
  UA_REQUIRE(!(out || out_hi) || (ldo >= N && ldo % 4 == 0), "ua_gemm_tf32x3_f32: bad ldo");
  UA_REQUIRE((uintptr_t)out % 16 == 0 && (uintptr_t)out_hi % 16 == 0 && (uintptr_t)out_lo % 16 == 0,
             "ua_gemm_tf32x3_f32: outputs must be 16-byte aligned");
  UA_REQUIRE(!gmax || M % 32 == 0, "ua_gemm_tf32x3_f32: the group max needs M %% 32 == 0");
  GemmParams p;
  UA_REQUIRE(act >= 0 && act <= 2, "ua_gemm_tf32x3_f32: act=%d (0 none, 1 relu, 2 gelu)", act);
  UA_REQUIRE(!residual || ((uintptr_t)residual % 16 == 0 && ldo >= N && ldo % 4 == 0),
             "ua_gemm_tf32x3_f32: residual must be 16-byte aligned with the outputs' leading dimension");
  p.M = M, p.N = N, p.K = K, p.bias = bias, p.group_bias = group_bias, p.act = act, p.residual = residual;
  p.out = out, p.out_hi = out_hi, p.out_lo = out_lo, p.ldo = ldo, p.gmax = gmax, p.gmax_hi = gmax_hi, p.gmax_lo = gmax_lo;
  // N tile: among the sizes that divide N, the one with the least (waves over the 148 SMs) x (tile width); wider wins
  // ties (fewer, larger MMAs and less A traffic per flop). The kernel is bound by L2 -> shared-memory operand traffic
  // (3xTF32 streams four operand tiles per k-block), so a single-wave N=192 tiling of N=384 does not beat two waves of
  // N=128 (measured: 60.8 vs 58.4 us at 7695 x 384 x 1536).
  int bn = 0;
  {
    const long long m_tiles = (M + kBM - 1) / kBM;
    long long best = 0;
    for (int cand : {256, 192, 128}) {
      if (N % cand) continue;
      if (g_gemm_bn > 0 ? cand != g_gemm_bn : cand == 192) continue;   // 192 only on request: never faster when measured
      const long long waves = (m_tiles * (N / cand) + kNumSMs - 1) / kNumSMs;
      const long long cost = waves * cand;
      if (!bn || cost < best) bn = cand, best = cost;
    }
    UA_UNSUPPORTED(!bn, "ua_gemm_tf32x3_f32: no N tile for N=%d (gemm_bn=%d)", N, g_gemm_bn);
  }
  CUtensorMap m_a_hi, m_a_lo, m_w_hi, m_w_lo;
  int rc;
  if ((rc = make_map(&m_a_hi, a_hi, M, K, lda, kBM)) != UA_OK) return rc;
  if ((rc = make_map(&m_a_lo, a_lo, M, K, lda, kBM)) != UA_OK) return rc;
  if ((rc = make_map(&m_w_hi, w_hi, N, K, ldw, bn)) != UA_OK) return rc;
  if ((rc = make_map(&m_w_lo, w_lo, N, K, ldw, bn)) != UA_OK) return rc;
  // output tiles of 32 rows x 32 columns (128-byte rows, 128B swizzle) for the staged TMA stores
  CUtensorMap m_o, m_o_hi, m_o_lo;
  if ((rc = make_map(&m_o, out, M, N, ldo, 32)) != UA_OK) return rc;
  if ((rc = make_map(&m_o_hi, out_hi, M, N, ldo, 32)) != UA_OK) return rc;
  if ((rc = make_map(&m_o_lo, out_lo, M, N, ldo, 32)) != UA_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (bn == 256) return launch_gemm<256, 2>(m_a_hi, m_a_lo, m_w_hi, m_w_lo, m_o, m_o_hi, m_o_lo, p, st);
  if (bn == 192) return launch_gemm<192, 2>(m_a_hi, m_a_lo, m_w_hi, m_w_lo, m_o, m_o_hi, m_o_lo, p, st);
  return launch_gemm<128, 3>(m_a_hi, m_a_lo, m_w_hi, m_w_lo, m_o, m_o_hi, m_o_lo, p, st);
}
