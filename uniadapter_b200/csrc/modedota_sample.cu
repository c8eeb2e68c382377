// MODE-DOTA cache pass of one whole batch-1 sample step, sm_100a -- and its class-sharded form with the logit exchange
// over NVLink peer memory fused into the same persistent kernel.
//
// The reference adapts every sample with three cache operations (Uni_Adapter.py:416-430):
//     dota_logits = predict(x.half())            dota_mixture.py:236-267
//     fit(x, prob_map)                            dota_mixture.py:162-234
//     fit(x_aug, prob_map)                        (the jittered view, the ORIGINAL prob_map)
// Every one of them is a pass over the (K,M,D) cache; per class they only touch that class's (M,D) tile. So ONE pass
// does all three: the tile is pulled from the TMA ring into registers once (a warp owns a (mode, 128*V-float chunk)),
// the predict logit is taken on the old state, fit #1 updates the registers, fit #2 evaluates its likelihood on the
// updated registers and updates them again, and the tile goes back to HBM once: 16*K*M*D bytes per sample step
// instead of 32 (ua_modedota_step_f32 twice). The arithmetic of each operation is the one of modedota.cu's
// single-sample kernel, operation by operation, so the result is bit-identical to the two-launch sequence.
//
// Class-sharded form (BASELINE cfg 4, SURVEY 8e: Objaverse-LVIS, K = 1156 split over P GPUs by class). Every rank
// holds the state and text rows of its class range and sees the same sample. One launch per rank and step:
//   A. CTA 0 stores the rank's zero-shot logits into slot [parity][rank][0] of EVERY peer's symmetric receive buffer
//      (plain stores to mapped peer memory = NVLink), fences at system scope and releases flag [0][rank] on every peer;
//      meanwhile every CTA has its first class tiles in flight;
//   B. every CTA acquires the P flags of exchange 0 in its OWN memory (bounded wait), copies the gathered zero-shot row
//      into shared memory and reduces its softmax statistics (gamma_class = prob_map needs max and sum over all K);
//   C. the class loop above over the rank's classes; each cache logit is stored into slot [parity][rank][1] of every
//      peer as soon as the class's predict is done (the "epilogue" of predict writes straight into the peers);
//   D. a grid-wide counter elects the last CTA: it releases flag [1][rank] on every peer, acquires the peers' flags,
//      and fuses the two gathered rows (entropy-weighted blend, Uni_Adapter.py:491-521) into the final logits + argmax,
//      replicated on every rank. sum(c) of the fusion weight comes from the closed form K + fits.
// No collective-library call, no second launch, CUDA-graph capturable; a peer that never arrives raises `err`, leaves
// the cache untouched and poisons the outputs with NaN instead of fusing stale logits.
// Single-GPU tests emulate the P ranks as ONE cooperative launch (gridDim.y = P, all CTAs co-resident), never as P
// launches that wait on one another.
#include <cooperative_groups.h>
#include "fuse_dev.cuh"
#include "modedota_params.cuh"

namespace ua {

int g_p2p_timeout_ms = 2000;   // tuning: bound of every in-kernel wait for a peer
int g_sample_v = 0;            // tuning override: float4s per lane (0 = heuristic)
int g_sample_g = 0;            // tuning override: warp groups per CTA (0 = heuristic)
int g_sample_per = 0;          // tuning: CTAs per SM of the unsharded launch (0 = what the occupancy calculator allows)
int g_sample_skip = 0;         // diagnosis only: bit 0 skips the class loop, bit 1 the fusion (timing of the phases)

// diagnosis: phase time stamps (globaltimer, ns) of rank 0's CTA 0 (slots 0-7) and of its last CTA (slots 8-15)
__device__ long long g_sample_trace[16];
int g_sample_trace_on = 0;

namespace {

__device__ __forceinline__ void trace(int on, int slot) {
  if (on && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_sample_trace[slot] = t;
  }
}

constexpr int kMaxLaunchRanks = 8;   // ranks emulated by one launch (tests); a real rank launches exactly one

// Everything one rank (or the plain single-GPU form: rank 0 of 1) needs. Lives in the kernel parameter space
// (__grid_constant__): the fields are read through the constant bank where they are used instead of occupying
// registers across the class loop.
struct RankParams {
  const float* x_fit;    // [S,D] normalised sample (predict uses its fp16 rounding, Uni_Adapter.py:416)
  const float* x_fit2;   // [S,D] normalised jittered view, or null (one fit only)
  const float* gamma;    // [S,ldg] prob_map (null in the sharded form: computed from the gathered zero-shot row)
  float *mu, *var, *pi, *c, *class_counts;
  float* out_logits;     // [S,ldo] or null
  // sharded form only
  const float* text_local;   // [K_local, D] text rows of this rank: the kernel normalises the rows and forms the zero-shot logits
  const float* clip_local;   // or: zero-shot logits computed by the caller (x_fit / x_fit2 then hold NORMALISED rows)
  float* const* peer_recv;
  int* const* peer_flag;
  int* seq;
  int* err;
  unsigned* done;
  float* c_sum;
  float* out_final;
  int* out_argmax;
  float* out_clip;
  float* out_dota;
  int K, rank, k_lo, ldg, kg_off, ldo, ko_off, pad;
};

struct LaunchParams {
  RankParams r[kMaxLaunchRanks];
  int S, M, D, stages, want_pred, sharded, P, Ktot, K_pad, skip, trace;
  float eps, rho, eta;
  long long timeout_cycles;
};

// threads per CTA the register budget of a (V, G) variant allows (64 registers at 1024 threads, 128 at 512)
constexpr int s_max_threads(int V, int G) { return V <= 2 ? 1024 : (V <= 5 ? 512 : 256 * G); }

template <bool SHORT>
__device__ __forceinline__ float modes_max(float v) {
  if (!SHORT) {
    v = fmaxf(v, __shfl_xor_sync(kFullMask, v, 16));
    v = fmaxf(v, __shfl_xor_sync(kFullMask, v, 8));
  }
  v = fmaxf(v, __shfl_xor_sync(kFullMask, v, 4));
  v = fmaxf(v, __shfl_xor_sync(kFullMask, v, 2));
  return fmaxf(v, __shfl_xor_sync(kFullMask, v, 1));
}
template <bool SHORT>
__device__ __forceinline__ float modes_sum(float v) {
  if (!SHORT) {
    v += __shfl_xor_sync(kFullMask, v, 16);
    v += __shfl_xor_sync(kFullMask, v, 8);
  }
  v += __shfl_xor_sync(kFullMask, v, 4);
  v += __shfl_xor_sync(kFullMask, v, 2);
  return v + __shfl_xor_sync(kFullMask, v, 1);
}

__device__ __forceinline__ void group_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ int ld_volatile_shared(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_shared(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// data written by a peer into THIS GPU's memory, read after the acquire of the peer's flag: L2 is the point of coherence,
// so an L1-bypassing load is enough (and, unlike ld.volatile, several of them may be in flight)
__device__ __forceinline__ float ld_peer_written(const float* p) { return __ldcg(p); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// class k of the full range -> (owner rank, index inside the owner's shard); the first K mod P ranks own one class more
__device__ __forceinline__ void owner_of(int k, int K, int P, int& r, int& i) {
  const int base = K / P, extra = K % P, cut = extra * (base + 1);
  if (k < cut) {
    r = k / (base + 1), i = k - r * (base + 1);
  } else {
    const int q = (k - cut) / max(base, 1);
    r = extra + q, i = (k - cut) - q * base;
  }
}

// Bounded acquire of P flags (threads 0..P-1 poll one flag each); returns false when a peer did not arrive.
__device__ __forceinline__ bool wait_flags(const int* flags, int P, int seq, long long timeout, int* s_flag) {
  if (threadIdx.x == 0) *s_flag = 0;
  __syncthreads();
  if ((int)threadIdx.x < P) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + threadIdx.x) < seq) {
      if (clock64() - t0 > timeout) {
        atomicExch(s_flag, 1 + (int)threadIdx.x);
        break;
      }
    }
  }
  __syncthreads();
  return *s_flag == 0;
}

template <int V, int G>
__global__ void __launch_bounds__(s_max_threads(V, G), 1)
    modedota_sample_kernel(const __grid_constant__ LaunchParams lp) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ int s_flag;
  __shared__ float s_stat[2];
  __shared__ float s_tmp[64];
  __shared__ float s_bestv[32];
  __shared__ unsigned s_besti[32];
  const int tid = threadIdx.x, lane = tid & 31, T = blockDim.x;
  const bool sharded = lp.sharded != 0;
  const RankParams& R = lp.r[blockIdx.y];
  const RankParams& p = R;
  struct { int P, K, K_pad; float rho, eta; long long timeout_cycles; } sh = {lp.P, lp.Ktot, lp.K_pad, lp.rho, lp.eta,
                                                                           lp.timeout_cycles};
  int seq = 0, par = 0;
  const int k_lo = R.k_lo;
  if (sharded) {
    seq = *reinterpret_cast<volatile int*>(R.seq) + 1;
    par = seq & 1;
  }
  const bool fit2 = p.x_fit2 != nullptr;
  const bool pred = lp.want_pred != 0;

  const int M = lp.M, D = lp.D, MD = M * D, NS = lp.stages;
  const int chunks = D / (128 * V);
  const int gwarps = M * chunks;                               // warps per group
  const int warp = (tid >> 5) % gwarps, grp = (tid >> 5) / gwarps;
  const int wm = warp / chunks, wch = warp - wm * chunks;      // this warp's mode and chunk
  const int d0 = wch * 128 * V + lane * 4;                    // first float of this lane (then every 128 floats)
  const bool SHORTM = M <= 8;
  const uint32_t tile_bytes = (uint32_t)MD * sizeof(float);

  float* s_tiles = reinterpret_cast<float*>(s_raw);                        // [NS][2][MD]
  float* s_part = s_tiles + (size_t)NS * 2 * MD;                           // [G][3][32 warps][4]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_part + G * 3 * 32 * 4);  // [NS]
  int* s_issued = reinterpret_cast<int*>(s_bar + 8);                       // [NS] highest class index armed per stage
  float* s_clip = reinterpret_cast<float*>(s_issued + 8);                  // [K] gathered zero-shot row (sharded)
  // [D] the fp16-rounded sample predict() sees (Uni_Adapter.py:416), rounded once per CTA when there is one stream
  // [3][D] rows of the sample when there is one stream: its fp16 rounding (what predict() sees, Uni_Adapter.py:416), the
  // sample, the jittered view -- staged once per CTA, read by every class (shared-memory loads instead of global ones)
  float* s_rows = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_clip + (sharded ? lp.Ktot : 0)) + 15) & ~uintptr_t(15));
  const bool rows_smem = lp.S == 1;

  const int total = lp.S * p.K;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_mine = (first < total && !(lp.skip & 1)) ? (total - first + stride - 1) / stride : 0;

  auto item_of = [&](int j) { return first + j * stride; };   // j-th class of this CTA
  auto issue_load = [&](int j) {   // -> stage j % NS
    const int stage = j % NS;
    const size_t item = (size_t)item_of(j);
    float* dst = s_tiles + (size_t)stage * 2 * MD;
    mbar_expect_tx(&s_bar[stage], 2 * tile_bytes);
    bulk_g2s(dst, p.mu + item * MD, tile_bytes, &s_bar[stage]);
    bulk_g2s(dst + MD, p.var + item * MD, tile_bytes, &s_bar[stage]);
  };
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&s_bar[i], 1);
      s_issued[i] = i < n_mine ? i : -1;
    }
    fence_mbar_init();
    for (int j = 0; j < NS && j < n_mine; ++j) issue_load(j);   // the first tiles fly while the exchange below runs
  }
  const int tr0 = lp.trace && blockIdx.x == 0 && blockIdx.y == 0;
  trace(tr0, 0);
  const bool own_head = sharded && R.text_local != nullptr;
  float* s_xs = s_rows + 3 * D;        // [D] scale * xnorm: the left operand of the zero-shot logits (own_head)
  if (own_head) {
    // L2-normalise the raw rows exactly as ua_head_f32's l2norm_kernel does (256 threads, strided FMA sums, warp sums,
    // serial sum over the warps, one division per element), so that the sharded step sees the unsharded step's features
    const int NT = min(T, 256), NW = NT >> 5;
    // both raw rows land in shared memory first (all loads in flight together: one DRAM round trip), then the two sums
    for (int d = tid; d < D; d += T) {
      s_rows[D + d] = __ldg(p.x_fit + d);
      s_rows[2 * D + d] = fit2 ? __ldg(p.x_fit2 + d) : 0.f;
    }
    __syncthreads();
    float ss0 = 0.f, ss1 = 0.f;
    if (tid < NT)
      for (int d = tid; d < D; d += NT) {
        const float v0 = s_rows[D + d], v1 = s_rows[2 * D + d];
        ss0 = fmaf(v0, v0, ss0);
        ss1 = fmaf(v1, v1, ss1);
      }
    ss0 = warp_sum(ss0);
    ss1 = warp_sum(ss1);
    if (tid < NT && lane == 0) s_tmp[tid >> 5] = ss0, s_tmp[16 + (tid >> 5)] = ss1;
    __syncthreads();
    float tot0 = 0.f, tot1 = 0.f;
    for (int w = 0; w < NW; ++w) tot0 += s_tmp[w], tot1 += s_tmp[16 + w];
    const float nrm0 = sqrtf(tot0), nrm1 = sqrtf(tot1);
    for (int d = tid; d < D; d += T) {
      const float xn = __fdiv_rn(s_rows[D + d], nrm0);
      s_rows[d] = __half2float(__float2half_rn(xn));
      s_rows[D + d] = xn;
      s_xs[d] = __fmul_rn(100.0f, xn);            // the reference scales x before the contraction (Uni_Adapter.py:57)
      if (fit2) s_rows[2 * D + d] = __fdiv_rn(s_rows[2 * D + d], nrm1);
    }
  } else if (rows_smem) {
    for (int d = tid; d < D; d += T) {
      const float xv = __ldg(p.x_fit + d);
      s_rows[d] = __half2float(__float2half_rn(xv));
      s_rows[D + d] = xv;
      s_rows[2 * D + d] = fit2 ? __ldg(p.x_fit2 + d) : 0.f;
    }
  }
  __syncthreads();

  trace(tr0, 1);
  bool ok = true;
  if (sharded) {
    // ---- A: this rank's zero-shot logits -> every peer ----------------------------------------------------------
    if (own_head) {
      // every CTA forms the logits of ITS classes (one warp per class: lane-strided float4 FMAs + butterfly sum, the
      // arithmetic of ua_head_f32's logits_kernel) and stores them straight into every peer; a grid-wide counter
      // elects the CTA that releases the flags
      const int nwarps_cta = T >> 5, wid = tid >> 5;
      for (int jj = wid; jj < n_mine; jj += nwarps_cta) {
        const int k = item_of(jj);
        const float* trow = R.text_local + (size_t)k * D;
        float acc = 0.f;
        for (int d = lane * 4; d < D; d += 128) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(trow + d));
          const float4 xv = *reinterpret_cast<const float4*>(s_xs + d);
          acc = fmaf(xv.x, t4.x, acc);
          acc = fmaf(xv.y, t4.y, acc);
          acc = fmaf(xv.z, t4.z, acc);
          acc = fmaf(xv.w, t4.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
          const size_t off = ((size_t)(par * sh.P + R.rank) * 2 + 0) * sh.K_pad + k;
          for (int r = 0; r < sh.P; ++r) R.peer_recv[r][off] = acc;
        }
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();          // device scope is enough here: the electing CTA below fences at system scope (cumulative)
        if (atomicAdd(R.done + 1, 1u) == gridDim.x - 1) {
          R.done[1] = 0u;
          __threadfence_system();
          for (int r = 0; r < sh.P; ++r) st_release_sys(R.peer_flag[r] + (0 * sh.P + R.rank), seq);
        }
      }
    } else if (blockIdx.x == 0) {
      for (int r = 0; r < sh.P; ++r) {
        float* dst = R.peer_recv[r] + ((size_t)(par * sh.P + R.rank) * 2 + 0) * sh.K_pad;
        for (int i = tid; i < p.K; i += T) dst[i] = R.clip_local[i];
      }
      __threadfence_system();
      __syncthreads();
      if (tid < sh.P) st_release_sys(R.peer_flag[tid] + (0 * sh.P + R.rank), seq);
    }
    trace(tr0, 2);
    // ---- B: gathered zero-shot row + softmax statistics (every CTA) ---------------------------------------------
    ok = wait_flags(R.peer_flag[R.rank] + 0 * sh.P, sh.P, seq, sh.timeout_cycles, &s_flag);
    if (ok) {
      const float* recv = R.peer_recv[R.rank] + (size_t)par * sh.P * 2 * sh.K_pad;
      for (int k = tid; k < sh.K; k += T) {
        int r, i;
        owner_of(k, sh.K, sh.P, r, i);
        s_clip[k] = ld_peer_written(recv + ((size_t)r * 2 + 0) * sh.K_pad + i);
      }
      __syncthreads();
      float mx = -INFINITY;
      for (int k = tid; k < sh.K; k += T) mx = fmaxf(mx, s_clip[k]);
      mx = block_max(mx, s_tmp);
      float se = 0.f;
      for (int k = tid; k < sh.K; k += T) se += expf(s_clip[k] - mx);
      se = block_sum(se, s_tmp);
      if (tid == 0) s_stat[0] = mx, s_stat[1] = se;
      __syncthreads();
    } else if (tid == 0) {
      atomicExch(R.err, 1);
    }
  }
  trace(tr0, 3);
  const float sm_max = sharded ? s_stat[0] : 0.f, sm_sum = sharded ? s_stat[1] : 1.f;

  if (ok) {
    // per-class scalars, fetched one class (of this group) ahead: pi/c of mode `lane`, gamma_class, class_counts
    float nx_pi = 0.f, nx_c = 0.f, nx_g = 0.f, nx_cc = 0.f;
    auto fetch_small = [&](int j) {
      const int item = item_of(j);
      if (lane < M) {
        nx_pi = __ldg(p.pi + (size_t)item * M + lane);
        nx_c = __ldg(p.c + (size_t)item * M + lane);
      }
      const int s = item / p.K, k = item - s * p.K;
      // prob_map[k] = exp(l_k - max) / sum: the softmax of the zero-shot row (Uni_Adapter.py:70)
      nx_g = sharded ? __fdiv_rn(expf(s_clip[k_lo + k] - sm_max), sm_sum)
                     : __ldg(p.gamma + (size_t)s * p.ldg + p.kg_off + k);
      if (warp == 0) nx_cc = __ldg(p.class_counts + item);
    };
    if (grp < n_mine) fetch_small(grp);
    const int nfit = fit2 ? 2 : 1;

    int it = 0;   // group-local iteration (parity of the partial-sum buffer)
    for (int j = grp; j < n_mine; j += G, ++it) {
      const int stage = j % NS;
      const uint32_t parity = (uint32_t)(j / NS) & 1u;
      const int item = item_of(j);
      const int s = item / p.K, k = item - s * p.K;
      const float* t_mu = s_tiles + (size_t)stage * 2 * MD + (size_t)wm * D + d0;
      const float* t_var = t_mu + MD;
      const float cur_g = nx_g;
      float c_cur = nx_c, pi_cur = nx_pi, cc_cur = nx_cc;      // soft counts / mixing weights / class count so far
      if (j + G < n_mine) fetch_small(j + G);
      // rows of this class's stream: staged in shared memory when there is one stream, else L1-resident global rows
      const float* row_pred = rows_smem ? s_rows + d0 : nullptr;
      const float* row_fit0 = rows_smem ? s_rows + D + d0 : p.x_fit + (size_t)s * D + d0;
      const float* row_fit1 = rows_smem ? s_rows + 2 * D + d0 : (fit2 ? p.x_fit2 + (size_t)s * D + d0 : row_fit0);

      // The ring is shared by the groups and an mbarrier parity only tells adjacent phases apart: a group may look at a
      // stage only once the load of ITS class has been armed (by the group that drained the stage's previous class).
      if (lane == 0) {
        while (ld_volatile_shared(&s_issued[stage]) < j) {
        }
      }
      __syncwarp();
      mbar_wait(&s_bar[stage], parity);

      float* o_mu = p.mu + (size_t)item * MD + (size_t)wm * D + d0;
      float* o_var = p.var + (size_t)item * MD + (size_t)wm * D + d0;
      float4 mu4[V], var4[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        mu4[v] = *reinterpret_cast<const float4*>(t_mu + 128 * v);
        var4[v] = *reinterpret_cast<const float4*>(t_var + 128 * v);
      }
      // One body for both fits (fit #2 = the same code on the registers fit #1 left): the loop is NOT unrolled, so the
      // instruction footprint stays inside the instruction cache (the two-copy version spent 21 % of its warp stalls
      // waiting for instructions: profiles/r2_ncu_sample_step.txt).
#pragma unroll 1
      for (int f = 0; f < nfit; ++f) {
        const float* xrow = f == 0 ? row_fit0 : row_fit1;
        const bool do_pred = pred && f == 0;
        // ---- likelihood pass: log-determinant and Mahalanobis partial sums (fit row; predict row on the first pass) ---
        float accp = 0.f, accf = 0.f, mprod = 1.f;
        int esum = 0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float mm[4] = {mu4[v].x, mu4[v].y, mu4[v].z, mu4[v].w};
          const float vv[4] = {var4[v].x, var4[v].y, var4[v].z, var4[v].w};
          const float4 t = rows_smem ? *reinterpret_cast<const float4*>(xrow + 128 * v)
                                     : __ldg(reinterpret_cast<const float4*>(xrow + 128 * v));
          const float ff[4] = {t.x, t.y, t.z, t.w};
          float pp[4] = {0.f, 0.f, 0.f, 0.f};
          if (do_pred) {
            if (rows_smem) {
              const float4 u = *reinterpret_cast<const float4*>(row_pred + 128 * v);
              pp[0] = u.x, pp[1] = u.y, pp[2] = u.z, pp[3] = u.w;
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) pp[q] = __half2float(__float2half_rn(ff[q]));   // Uni_Adapter.py:416
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float vq = fmaxf(__fadd_rn(vv[q], lp.eps), 1e-8f);
            const float inv = rcp_rn_normal(vq);   // one correctly-rounded reciprocal shared by both rows
            const float dp = __fsub_rn(pp[q], mm[q]);
            accp = fmaf(dp * dp, inv, accp);       // (unused on the second pass: three instructions, no branch)
            const float dq = __fsub_rn(ff[q], mm[q]);
            accf = fmaf(dq * dq, inv, accf);
            // log-determinant in product form: sum_d log v = ln2 * sum_d e_d + log prod_d m_d, v = m * 2^e, m in [1,2)
            const uint32_t bits = __float_as_uint(vq);
            esum += (int)(bits >> 23);
            mprod *= __uint_as_float((bits & 0x007fffffu) | 0x3f800000u);
          }
        }
        float ld = fmaf((float)(esum - 127 * 4 * V), 0.693147182f, logf(mprod));
        accp = warp_sum(accp);
        accf = warp_sum(accf);
        ld = warp_sum(ld);
        const float logpi = logf(__fadd_rn(pi_cur, 1e-10f));   // off the critical path: before the barrier
        float* part = s_part + ((grp * 3 + (f == 0 ? (it & 1) : 2)) * 32) * 4;
        if (lane == 0) *reinterpret_cast<float4*>(part + warp * 4) = make_float4(accp, accf, ld, 0.f);
        if (G == 1) __syncthreads(); else group_barrier(1 + grp, gwarps * 32);
        // every warp of the group has pulled the stage into registers: re-arm it NS classes ahead
        if (f == 0 && warp == 0 && lane == 0 && j + NS < n_mine) {
          issue_load(j + NS);
          st_volatile_shared(&s_issued[stage], j + NS);
        }

        // ---- responsibilities: every warp evaluates all M modes (lane = mode; lanes 8.. idle when M <= 8) ----------
        float mp = 0.f, mf = 0.f, ldet = 0.f;
        if (lane < M) {
          for (int ch = 0; ch < chunks; ++ch) {
            const float4 q = *reinterpret_cast<const float4*>(part + (lane * chunks + ch) * 4);
            mp += q.x, mf += q.y, ldet += q.z;
          }
        }
        {   // (warp collectives stay outside of run-time conditions: the compiler then keeps them in line)
          const float lj = lane < M ? __fadd_rn(logpi, __fmul_rn(-0.5f, __fadd_rn(ldet, mp))) : -INFINITY;
          const float mx = SHORTM ? modes_max<true>(lj) : modes_max<false>(lj);
          const float ex = lane < M ? expf(__fsub_rn(lj, mx)) : 0.f;
          const float se = SHORTM ? modes_sum<true>(ex) : modes_sum<false>(ex);
          if (do_pred && warp == 0 && lane == 0) {
            const float lse = __fadd_rn(logf(se), mx);
            if (p.out_logits) p.out_logits[(size_t)s * p.ldo + p.ko_off + k] = lse;
            if (sharded) {   // the predict epilogue writes the class's cache logit straight into every peer
              const size_t off = ((size_t)(par * sh.P + R.rank) * 2 + 1) * sh.K_pad + k;
              for (int r = 0; r < sh.P; ++r) R.peer_recv[r][off] = lse;
            }
          }
        }
        float rd, g0, cold;
        {
          const float lj = lane < M ? __fadd_rn(logpi, __fmul_rn(-0.5f, __fadd_rn(ldet, mf))) : -INFINITY;
          const float mx = SHORTM ? modes_max<true>(lj) : modes_max<false>(lj);
          const float ex = lane < M ? expf(__fsub_rn(lj, mx)) : 0.f;
          const float se = SHORTM ? modes_sum<true>(ex) : modes_sum<false>(ex);
          const float lse = __fadd_rn(logf(se), mx);
          // gamma[b=0, mode] = gamma_class * exp(log_joint - logsumexp), the reference's form (dota_mixture.py:182-186);
          // both fits use the ORIGINAL prob_map (Uni_Adapter.py:417,430)
          const float gam = lane < M ? __fmul_rn(cur_g, expf(__fsub_rn(lj, lse))) : 0.f;
          const float cnew = __fadd_rn(c_cur, gam);
          const float rden = rcp_rn_normal(__fadd_rn(cnew, 1e-10f));
          const float cm = lane < M ? cnew : 0.f;
          const float ck = SHORTM ? modes_sum<true>(cm) : modes_sum<false>(cm);
          cold = __shfl_sync(kFullMask, c_cur, wm);
          g0 = __shfl_sync(kFullMask, gam, wm);
          rd = __shfl_sync(kFullMask, rden, wm);
          c_cur = cnew;
          pi_cur = __fdiv_rn(cnew, __fadd_rn(ck, 1e-10f));
          cc_cur = cc_cur + cur_g;
        }
        // ---- M-step on the registers of this warp's (mode, chunk) ---------------------------------------------------
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float mm[4] = {mu4[v].x, mu4[v].y, mu4[v].z, mu4[v].w};
          float vv[4] = {var4[v].x, var4[v].y, var4[v].z, var4[v].w};
          const float4 t = rows_smem ? *reinterpret_cast<const float4*>(xrow + 128 * v)
                                     : __ldg(reinterpret_cast<const float4*>(xrow + 128 * v));
          const float ff[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float wx = __fmul_rn(g0, ff[q]);
            const float wxsq = __fmul_rn(g0, __fmul_rn(ff[q], ff[q]));
            const float mu_ = mm[q];
            mm[q] = __fmul_rn(__fadd_rn(__fmul_rn(cold, mu_), wx), rd);
            const float term2 = __fmul_rn(__fmul_rn(-2.0f, mu_), wx);
            const float term3 = __fmul_rn(g0, __fmul_rn(mu_, mu_));
            const float wsd = __fadd_rn(__fadd_rn(wxsq, term2), term3);
            vv[q] = fmaxf(__fmul_rn(__fadd_rn(__fmul_rn(cold, vv[q]), wsd), rd), 1e-8f);
          }
          mu4[v] = make_float4(mm[0], mm[1], mm[2], mm[3]);
          var4[v] = make_float4(vv[0], vv[1], vv[2], vv[3]);
          if (f == nfit - 1) {   // the class's new state: straight from the registers to HBM, interleaved with the math
            *reinterpret_cast<float4*>(o_mu + 128 * v) = mu4[v];
            *reinterpret_cast<float4*>(o_var + 128 * v) = var4[v];
          }
        }
      }
      if (warp == 0) {
        if (lane < M) {
          p.c[(size_t)item * M + lane] = c_cur;
          p.pi[(size_t)item * M + lane] = pi_cur;
        }
        if (lane == 0) p.class_counts[item] = cc_cur;
      }
    }
  } else if (tid == 0) {
    // a peer is missing: the cache stays untouched; drain the tiles that are already in flight before leaving
    for (int j = 0; j < NS && j < n_mine; ++j) mbar_wait(&s_bar[j % NS], 0u);
  }
  if (!sharded) return;
  trace(tr0, 4);

  // ---- D: the last CTA of the rank closes exchange 1 and fuses the gathered rows ------------------------------------
  __syncthreads();
  if (tid == 0) {
    __threadfence();                 // this CTA's peer stores are ordered before its arrival (the last CTA fences at system scope)
    s_flag = atomicAdd(R.done, 1u) == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (!s_flag) return;
  const int tr1 = lp.trace && blockIdx.y == 0;
  trace(tr1, 8);
  if (tid == 0) *R.done = 0u;
  __threadfence_system();
  if (tid < sh.P) st_release_sys(R.peer_flag[tid] + (1 * sh.P + R.rank), ok ? seq : -seq);
  bool ok2 = ok;
  if (ok2) {
    // a peer that aborted publishes -seq: wait for |flag| >= seq, then look at the sign
    if (tid == 0) s_flag = 0;
    __syncthreads();
    if (tid < sh.P) {
      const int* fl = R.peer_flag[R.rank] + 1 * sh.P + tid;
      const long long t0 = clock64();
      for (;;) {
        const int v = ld_acquire_sys(fl);
        if (v >= seq) break;
        if (v <= -seq || clock64() - t0 > sh.timeout_cycles) {
          atomicExch(&s_flag, 1 + tid);
          break;
        }
      }
    }
    __syncthreads();
    ok2 = s_flag == 0;
  }
  const int K = sh.K;
  trace(tr1, 9);
  if (ok2 && !(lp.skip & 2)) {
    float* s_dota = s_tiles;         // the ring is idle: every load of this CTA has been consumed
    const float* recv = R.peer_recv[R.rank] + (size_t)par * sh.P * 2 * sh.K_pad;
    for (int k = tid; k < K; k += T) {
      int r, i;
      owner_of(k, K, sh.P, r, i);
      s_dota[k] = ld_peer_written(recv + ((size_t)r * 2 + 1) * sh.K_pad + i);
    }
    __syncthreads();
    trace(tr1, 10);
    // sum(c) in closed form: every fit of a batch-1 sample adds exactly 1
    const float csum = *R.c_sum + (fit2 ? 2.f : 1.f);
    const float w = cache_weight(csum, (float)K * (float)M, sh.rho, 1.f, sh.eta);
    // the two softmax entropies (zero-shot row, scaled cache row) share their passes: three block reductions of a
    // pair each instead of six (the arithmetic per row is fuse_dev.cuh's softmax_entropy)
    float mc = -INFINITY, md = -INFINITY;
    for (int k = tid; k < K; k += T) mc = fmaxf(mc, s_clip[k]), md = fmaxf(md, __fmul_rn(w, s_dota[k]));
    block_max2(mc, md, s_tmp);
    float sc = 0.f, sd = 0.f;
    for (int k = tid; k < K; k += T) sc += expf(s_clip[k] - mc), sd += expf(__fmul_rn(w, s_dota[k]) - md);
    block_sum2(sc, sd, s_tmp);
    float ec = 0.f, ed = 0.f;
    for (int k = tid; k < K; k += T) {
      const float pc = __fdiv_rn(expf(s_clip[k] - mc), sc), pd = __fdiv_rn(expf(__fmul_rn(w, s_dota[k]) - md), sd);
      ec += pc * logf(pc + 1e-10f);
      ed += pd * logf(pd + 1e-10f);
    }
    block_sum2(ec, ed, s_tmp);
    const float hc = -ec, hd = -ed;
    float wc, wd;
    entropy_weights(hc, hd, wc, wd);
    trace(tr1, 11);
    float best = -INFINITY;
    unsigned besti = 0xffffffffu;
    for (int k = tid; k < K; k += T) {
      const float f = __fadd_rn(__fmul_rn(wc, s_clip[k]), __fmul_rn(wd, __fmul_rn(w, s_dota[k])));
      R.out_final[k] = f;
      if (R.out_clip) R.out_clip[k] = s_clip[k];
      if (R.out_dota) R.out_dota[k] = s_dota[k];
      if (f > best) best = f, besti = k;
    }
    const float wmx = warp_max(best);
    const unsigned wmi = __reduce_min_sync(kFullMask, best == wmx ? besti : 0xffffffffu);
    __syncthreads();
    if (lane == 0) s_bestv[tid >> 5] = wmx, s_besti[tid >> 5] = wmi;
    __syncthreads();
    if (tid == 0) {
      float m = s_bestv[0];
      unsigned a = s_besti[0];
      for (int q = 1; q < (T >> 5); ++q)
        if (s_bestv[q] > m || (s_bestv[q] == m && s_besti[q] < a)) m = s_bestv[q], a = s_besti[q];
      *R.out_argmax = (int)a;
      *R.c_sum = csum;
    }
  } else {
    for (int k = tid; k < K; k += T) R.out_final[k] = __int_as_float(0x7fc00000);   // never fuse stale logits
    if (tid == 0) {
      *R.out_argmax = -1;
      atomicExch(R.err, 2);
    }
  }
  trace(tr1, 12);
  if (tid == 0) *R.seq = seq;
}

struct Plan {
  int V = 0, G = 1, NS = 0, threads = 0;
  size_t smem = 0;
};

// widest lanes first for two groups (more ILP per thread, fewer warps per barrier), else the narrowest fit; the ring
// depth is a multiple of the group count where it can be (a stage then always serves the same group)
Plan make_plan(int M, int D, size_t extra_smem, long long items = 1 << 30) {
  static const int kV[] = {1, 2, 4, 5, 8, 10};
  Plan pl;
  for (int gtry = 2; gtry >= 1 && !pl.V; --gtry) {
    if (g_sample_g > 0 && gtry != g_sample_g) continue;
    for (int i = 5; i >= 0 && !pl.V; --i) {
      const int cand = kV[i];
      if (g_sample_v > 0 && cand != g_sample_v) continue;
      if (D % (128 * cand)) continue;
      const int thr = gtry * 32 * M * (D / (128 * cand));
      if (thr > s_max_threads(cand, gtry) || thr > 1024) continue;
      if (gtry == 2 && thr < 256 && !g_sample_g) continue;   // too few warps to hide anything
      pl.V = cand, pl.G = gtry, pl.threads = thr;
    }
  }
  if (!pl.V) return pl;
  // few (stream, class) tiles per SM: nothing for a second warp group to overlap with -> one group, SAME lane width (the
  // lane width fixes the summation order over D, the group count does not: results stay bit-identical across tilings,
  // ranks and the two-launch sequence). Measured (tools/dbg/sample_sweep_cfg2.py): cfg 3 shape 30.7 -> 26.6 us, cfg 2 33.5 -> 31.5 us.
  if (!g_sample_g && pl.G == 2 && items < 4 * kNumSMs && pl.threads / 2 >= 128) pl.G = 1, pl.threads /= 2;
  const size_t tile_bytes = (size_t)M * D * sizeof(float);
  const size_t small = (size_t)pl.G * 3 * 32 * 4 * sizeof(float) + 8 * sizeof(uint64_t) + 8 * sizeof(int) + extra_smem + 128;
  const size_t budget = 226 * 1024;
  for (int cand = 4; cand >= 2 && !pl.NS; --cand)
    if ((size_t)cand * 2 * tile_bytes + small <= budget && (cand <= 3 || 4 * 2 * tile_bytes <= 100 * 1024)) pl.NS = cand;
  if (!pl.NS) {
    pl.V = 0;
    return pl;
  }
  pl.smem = (size_t)pl.NS * 2 * tile_bytes + small;
  return pl;
}

// grid.x == 0 asks for the persistent grid: one CTA per resident slot as the occupancy calculator sees it (the register
// budget of the wide variants allows ONE 512-thread CTA per SM), capped by the number of (stream, class) tiles
template <int V, int G>
cudaError_t launch_vg(const LaunchParams& lp, dim3 grid, int threads, size_t smem, bool coop, cudaStream_t st) {
  auto kern = modedota_sample_kernel<V, G>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (!coop) {
    if (grid.x == 0) {
      int per_sm = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
      if (e != cudaSuccess) return e;
      if (per_sm < 1) per_sm = 1;
      if (per_sm > 4) per_sm = 4;
      if (g_sample_per > 0) per_sm = g_sample_per;
      long long g = (long long)kNumSMs * per_sm;
      const long long total = (long long)lp.S * lp.r[0].K;
      if (g > total) g = total;
      grid.x = (unsigned)g;
    }
    kern<<<grid, threads, smem, st>>>(lp);
    return cudaSuccess;
  }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if ((long long)grid.x * grid.y > (long long)per_sm * sms) return cudaErrorCooperativeLaunchTooLarge;
  void* args[] = {(void*)&lp};
  return cudaLaunchCooperativeKernel((void*)kern, grid, dim3(threads), args, smem, st);
}

template <int V>
cudaError_t launch_v(const LaunchParams& lp, int G, dim3 grid, int threads, size_t smem, bool coop, cudaStream_t st) {
  return G == 2 ? launch_vg<V, 2>(lp, grid, threads, smem, coop, st) : launch_vg<V, 1>(lp, grid, threads, smem, coop, st);
}

cudaError_t launch_plan(const Plan& pl, const LaunchParams& lp, dim3 grid, bool coop, cudaStream_t st) {
  switch (pl.V) {
    case 1: return launch_v<1>(lp, pl.G, grid, pl.threads, pl.smem, coop, st);
    case 2: return launch_v<2>(lp, pl.G, grid, pl.threads, pl.smem, coop, st);
    case 4: return launch_v<4>(lp, pl.G, grid, pl.threads, pl.smem, coop, st);
    case 5: return launch_v<5>(lp, pl.G, grid, pl.threads, pl.smem, coop, st);
    case 8: return launch_v<8>(lp, pl.G, grid, pl.threads, pl.smem, coop, st);
    default: return launch_v<10>(lp, pl.G, grid, pl.threads, pl.smem, coop, st);
  }
}

}  // namespace
}  // namespace ua

extern "C" int ua_modedota_sample_step_f32(const float* x_fit, const float* x_fit2, const float* gamma_class, int ldg,
                                           int k_gamma_offset, float* mu, float* var, float* pi, float* c,
                                           float* class_counts, int S, int K, int M, int D, float eps,
                                           float* out_logits, int ldo, int k_out_offset, void* stream) {
  using namespace ua;
  UA_REQUIRE(x_fit && gamma_class && mu && var && pi && c && class_counts,
             "ua_modedota_sample_step_f32: x_fit / gamma_class / state pointers must be non-NULL");
  UA_REQUIRE(S >= 1 && K >= 1 && M >= 1 && D >= 1, "ua_modedota_sample_step_f32: bad sizes S=%d K=%d M=%d D=%d", S, K, M, D);
  UA_REQUIRE(ldg >= k_gamma_offset + K, "ua_modedota_sample_step_f32: ldg=%d < offset+K", ldg);
  UA_REQUIRE(!out_logits || ldo >= k_out_offset + K, "ua_modedota_sample_step_f32: ldo=%d < offset+K", ldo);
  UA_UNSUPPORTED(M > kMaxM || D % 128 != 0, "ua_modedota_sample_step_f32: needs M <= %d and D %% 128 == 0 (M=%d D=%d)", kMaxM,
                 M, D);
  UA_UNSUPPORTED(((uintptr_t)mu | (uintptr_t)var | (uintptr_t)x_fit | (uintptr_t)x_fit2) & 15,
                 "ua_modedota_sample_step_f32: pointers must be 16-byte aligned");
  UA_UNSUPPORTED((long long)S * K > 0x3fffffffLL, "ua_modedota_sample_step_f32: S*K too large");
  const Plan pl = make_plan(M, D, (size_t)3 * D * sizeof(float) + 16, (long long)S * K);
  UA_UNSUPPORTED(!pl.V, "ua_modedota_sample_step_f32: no register tiling for M=%d D=%d", M, D);
  LaunchParams lp = {};
  RankParams& r = lp.r[0];
  r.x_fit = x_fit, r.x_fit2 = x_fit2, r.gamma = gamma_class;
  r.mu = mu, r.var = var, r.pi = pi, r.c = c, r.class_counts = class_counts, r.out_logits = out_logits;
  r.K = K, r.ldg = ldg, r.kg_off = k_gamma_offset, r.ldo = ldo, r.ko_off = k_out_offset;
  lp.S = S, lp.M = M, lp.D = D, lp.eps = eps, lp.stages = pl.NS, lp.want_pred = out_logits != nullptr;
  cudaError_t e = launch_plan(pl, lp, dim3(0u), false, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("ua_modedota_sample_step_f32: launch setup failed (%zu B smem): %s", pl.smem, cudaGetErrorString(e));
    return UA_ERR_CUDA;
  }
  return check_launch("ua_modedota_sample_step_f32");
}

extern "C" int ua_modedota_sharded_step_f32(const ua_shard_rank* ranks, int n_ranks, int P, int K, int K_pad, int M,
                                            int D, float eps, float rho, float eta, void* stream) {
  using namespace ua;
  UA_REQUIRE(ranks, "ua_modedota_sharded_step_f32: ranks is NULL");
  UA_REQUIRE(P >= 1 && P <= 64 && (n_ranks == 1 || n_ranks == P) && n_ranks <= kMaxLaunchRanks,
             "ua_modedota_sharded_step_f32: P=%d n_ranks=%d (n_ranks is 1, or P <= %d for the single-GPU emulation)", P, n_ranks,
             kMaxLaunchRanks);
  UA_REQUIRE(K >= P && M >= 1 && D >= 1 && K_pad >= (K + P - 1) / P, "ua_modedota_sharded_step_f32: bad sizes K=%d K_pad=%d", K,
             K_pad);
  UA_UNSUPPORTED(M > kMaxM || D % 128 != 0, "ua_modedota_sharded_step_f32: needs M <= %d and D %% 128 == 0 (M=%d D=%d)", kMaxM,
                 M, D);
  // the same (V, G) plan for every P: the in-kernel softmax and fusion reductions run over the CTA's threads, so the group
  // count is part of the summation order there and the result must not depend on the partition (P = 1, 2, 4, 8 give the
  // same bits, tests/test_gpu_adapters.py)
  const Plan pl = make_plan(M, D, (size_t)(K + 4 * D) * sizeof(float) + 16);
  UA_UNSUPPORTED(!pl.V || pl.threads < P, "ua_modedota_sharded_step_f32: no register tiling for M=%d D=%d K=%d", M, D, K);
  UA_UNSUPPORTED((size_t)K * sizeof(float) > (size_t)pl.NS * 2 * M * D * sizeof(float),
                 "ua_modedota_sharded_step_f32: K=%d too large for the fusion scratch", K);
  LaunchParams lp = {};
  const int base = K / P, extra = K % P;
  for (int i = 0; i < n_ranks; ++i) {
    const ua_shard_rank& h = ranks[i];
    UA_REQUIRE(h.rank >= 0 && h.rank < P, "ua_modedota_sharded_step_f32: rank %d out of range", h.rank);
    UA_REQUIRE(h.x_fit && (h.clip_local || h.text_local) && h.mu && h.var && h.pi && h.c && h.class_counts && h.peer_recv && h.peer_flag && h.seq &&
                   h.err && h.done && h.c_sum && h.out_final && h.out_argmax,
               "ua_modedota_sharded_step_f32: NULL pointer in rank struct %d", i);
    UA_UNSUPPORTED(((uintptr_t)h.mu | (uintptr_t)h.var | (uintptr_t)h.x_fit | (uintptr_t)h.x_fit2 | (uintptr_t)h.text_local) & 15,
                   "ua_modedota_sharded_step_f32: pointers must be 16-byte aligned");
    RankParams& r = lp.r[i];
    r.x_fit = h.x_fit, r.x_fit2 = h.x_fit2, r.gamma = nullptr;
    r.mu = h.mu, r.var = h.var, r.pi = h.pi, r.c = h.c, r.class_counts = h.class_counts, r.out_logits = nullptr;
    r.text_local = h.text_local, r.clip_local = h.clip_local, r.peer_recv = h.peer_recv, r.peer_flag = h.peer_flag;
    r.seq = h.seq, r.err = h.err, r.done = h.done, r.c_sum = h.c_sum;
    r.out_final = h.out_final, r.out_argmax = h.out_argmax, r.out_clip = h.out_clip, r.out_dota = h.out_dota;
    r.rank = h.rank;
    r.k_lo = h.rank * base + (h.rank < extra ? h.rank : extra);
    r.K = base + (h.rank < extra ? 1 : 0);
  }
  lp.S = 1, lp.M = M, lp.D = D, lp.eps = eps, lp.stages = pl.NS, lp.want_pred = 1;
  lp.skip = g_sample_skip, lp.trace = g_sample_trace_on;
  lp.sharded = 1, lp.P = P, lp.Ktot = K, lp.K_pad = K_pad, lp.rho = rho, lp.eta = eta;
  lp.timeout_cycles = (long long)g_p2p_timeout_ms * 2000000LL;      // ~2 GHz
  int gx = kNumSMs / n_ranks;
  const int k_local_max = (K + P - 1) / P;
  if (gx > k_local_max) gx = k_local_max;
  if (gx < 1) gx = 1;
  cudaError_t e = launch_plan(pl, lp, dim3((unsigned)gx, (unsigned)n_ranks), true, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("ua_modedota_sharded_step_f32: cooperative launch failed (%d x %d CTAs, %zu B smem): %s", gx, n_ranks, pl.smem,
              cudaGetErrorString(e));
    return UA_ERR_CUDA;
  }
  return check_launch("ua_modedota_sharded_step_f32");
}

extern "C" int ua_debug_sample_trace(int64_t* host_out16) {
  return cudaMemcpyFromSymbol(host_out16, ua::g_sample_trace, sizeof(long long) * 16) == cudaSuccess ? UA_OK : UA_ERR_CUDA;
}
