// MODE-DOTA cache step for sm_100a: predict (on the current state) + streaming-EM fit, one pass over the cache.
// Replaces dota_mixture.py:117-156 (_get_var/_log_likelihood), :162-234 (fit), :236-267 (predict).
//
// Data layout: mu,var [S,K,M,D] fp32 (S independent streams), pi,c [S,K,M], class_counts [S,K]; the (M,D) tile of
// one (stream,class) is contiguous, so the TMA unit moves it with one bulk copy per tensor. A CTA walks classes
// (stride gridDim.x) with a two-stage shared-memory ring: while class i is evaluated and updated in place, the
// tiles of class i+1 are in flight, and the updated tiles of class i-1 drain back to HBM with a bulk store. Every
// state byte crosses HBM exactly once in and once out per fit (16*K*M*D bytes), which is the roofline of the op.
//
// The M-step keeps the reference's operation order for the expanded-form variance (the part that cancels,
// SURVEY H5); the two divisions by (c_new + 1e-10) become one correctly-rounded reciprocal per mode.
#include "common.cuh"
#include "modedota_params.cuh"

namespace ua {

int g_modedota_threads = 0;  // tuning override
int g_modedota_logprod = 1;  // tuning: 1 = product-form log-determinant in the single-sample path
int g_modedota_groups = 0;   // tuning: warp groups per CTA of the single-sample path (0 = heuristic)
int g_modedota_v = 0;        // tuning override: floats4 per lane of the single-sample path (-1 disables that path)
int g_modedota_batch = 0;    // tuning: -1 disables the batched (cluster) path, N > 0 forces N D-splits

namespace {

template <int MM>
__global__ void __launch_bounds__(MM > 8 ? 512 : 1024, 1) modedota_step_kernel(const StepParams p) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const int M = p.M, D = p.D, MD = M * D;
  const int rows = p.Bp + p.B;
  const size_t tile_bytes = (size_t)MD * sizeof(float);
  const int wpm = max(1, nwarps / M);        // warps per mode
  const int slots = M * wpm;                 // (mode, d-range) work slots, dealt round-robin to the warps
  const bool vec4 = p.vec_ok != 0;           // rows are 16-byte tileable: float4 shared-memory / global accesses
  const int Cd = vec4 ? ((D / 4 + wpm - 1) / wpm) * 4 : (D + wpm - 1) / wpm;   // d-range per slot

  // shared-memory carve-up
  float* s_tiles = reinterpret_cast<float*>(s_raw);                       // [stages][2][MD]
  float* s_ll = s_tiles + (size_t)p.stages * 2 * MD;                      // [rows][M]   maha sums -> log joint
  float* s_gamma = s_ll + (size_t)kMaxRows * kMaxM;                       // [B][M]
  float* s_part = s_gamma + (size_t)kMaxRows * kMaxM;                     // [<=32 slots][3] (+ scratch), 32*3*kMaxM floats
  float* s_small = s_part + 32 * 3 * kMaxM;                               // logdet[M], cold[M], sumg[M], denom[M], logpi[M]
  float* s_gc = s_small + 8 * kMaxM;                                      // [kMaxRows] gamma_class column of this class
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_gc + kMaxRows);         // [2]

  float* s_logdet = s_small;
  float* s_cold = s_small + kMaxM;
  float* s_sumg = s_small + 2 * kMaxM;
  float* s_denom = s_small + 3 * kMaxM;
  float* s_logpi = s_small + 4 * kMaxM;

  const int total = p.S * p.K;
  if (p.use_bulk && tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue_load = [&](int item, int stage) {
    float* dst_mu = s_tiles + (size_t)stage * 2 * MD;
    float* dst_var = dst_mu + MD;
    mbar_expect_tx(&s_bar[stage], (uint32_t)(2 * tile_bytes));
    bulk_g2s(dst_mu, p.mu + (size_t)item * MD, (uint32_t)tile_bytes, &s_bar[stage]);
    bulk_g2s(dst_var, p.var + (size_t)item * MD, (uint32_t)tile_bytes, &s_bar[stage]);
  };

  auto phys = [&](int logical) { return logical; };
  int li = blockIdx.x;
  if (p.use_bulk && tid == 0 && li < total) issue_load(phys(li), 0);

  // Small per-class operands (pi, c, class_counts, gamma_class column) are fetched one class ahead into
  // registers, so their DRAM latency overlaps the previous class instead of sitting on the critical path.
  // Holders: threads [0,M) pi and c, thread 32 class_counts, threads 64+b the gamma of fit row b.
  float nx_pi = 0.f, nx_c = 0.f, nx_cc = 0.f, nx_g = 0.f;
  auto fetch_small = [&](int it_item) {
    const int fs = it_item / p.K, fk = it_item - fs * p.K;
    if (tid < M) {
      nx_pi = p.pi[(size_t)it_item * M + tid];
      nx_c = p.c[(size_t)it_item * M + tid];
    }
    if (p.B > 0) {
      if (tid == 32) nx_cc = p.class_counts[it_item];
      if (tid >= 64 && tid < 64 + p.B) nx_g = __ldg(p.gamma + ((size_t)fs * p.B + (tid - 64)) * p.ldg + p.kg_off + fk);
    }
  };
  if (li < total) fetch_small(phys(li));

  uint32_t phase_bits = 0;  // per-stage mbarrier parity
  int it = 0;
  for (; li < total; li += gridDim.x, ++it) {
    const int item = phys(li);
    const int stage = p.stages == 2 ? (it & 1) : 0;
    const int s = item / p.K, k = item - s * p.K;
    float* t_mu = s_tiles + (size_t)stage * 2 * MD;
    float* t_var = t_mu + MD;
    const float cur_pi = nx_pi, cur_c = nx_c, cur_cc = nx_cc;
    if (tid >= 64 && tid < 64 + p.B) s_gc[tid - 64] = nx_g;   // read after the barriers below
    if (li + gridDim.x < total) fetch_small(phys(li + gridDim.x));

    if (p.use_bulk) {
      if (tid == 0) {
        const int nxt = li + gridDim.x;
        if (p.stages == 2 && nxt < total) {
          bulk_wait_read<0>();  // the store that last read the other stage has drained its shared-memory reads
          issue_load(phys(nxt), stage ^ 1);
        }
        mbar_wait(&s_bar[stage], (phase_bits >> stage) & 1u);  // one poller; the barrier below releases the CTA
      }
      phase_bits ^= 1u << stage;
      __syncthreads();
    } else {
      for (int i = tid; i < MD; i += T) {
        t_mu[i] = p.mu[(size_t)item * MD + i];
        t_var[i] = p.var[(size_t)item * MD + i];
      }
      __syncthreads();
    }

    // ---- phase 1: per-mode log-determinant and Mahalanobis sums for every row (2 rows per sweep) ----------
    // A warp owns one (mode, d-range) slot, so a sweep ends in three warp reductions per warp instead of 3*M.
    for (int r0 = 0; r0 < max(rows, 1); r0 += 2) {
      const bool has0 = r0 < rows, has1 = r0 + 1 < rows;
      const float* xa = nullptr;
      const float* xb = nullptr;
      if (has0) xa = r0 < p.Bp ? p.x_pred + ((size_t)s * p.Bp + r0) * D : p.x_fit + ((size_t)s * p.B + (r0 - p.Bp)) * D;
      if (has1)
        xb = (r0 + 1) < p.Bp ? p.x_pred + ((size_t)s * p.Bp + r0 + 1) * D
                             : p.x_fit + ((size_t)s * p.B + (r0 + 1 - p.Bp)) * D;
      for (int slot = warp; slot < slots; slot += nwarps) {
        const int m = slot / wpm, part = slot - m * wpm;
        const int dlo = part * Cd, dhi = min(D, dlo + Cd);
        const float* mrow = t_mu + m * D;
        const float* vrow = t_var + m * D;
        float acc0 = 0.f, acc1 = 0.f, ld = 0.f;
        // One correctly-rounded reciprocal is shared by both rows instead of a division per row (<= 1.5 ulp per
        // term, unbiased). The log-determinant keeps one 1-ulp logf per element, exactly like the reference: the
        // initial variance is constant along d, so the rounding error of log(v) adds up coherently over D and
        // only the same per-element rounding reproduces the reference's mode responsibilities.
        if (vec4) {
#pragma unroll 2
          for (int d = dlo + 4 * lane; d < dhi; d += 128) {
            const float4 m4 = *reinterpret_cast<const float4*>(mrow + d);
            const float4 v4 = *reinterpret_cast<const float4*>(vrow + d);
            const float4 a4 = has0 ? __ldg(reinterpret_cast<const float4*>(xa + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 b4 = has1 ? __ldg(reinterpret_cast<const float4*>(xb + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
            const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float v = fmaxf(__fadd_rn(vv[q], p.eps), 1e-8f);
              const float inv = rcp_rn_normal(v);
              const float da = __fsub_rn(aa[q], mm[q]), db = __fsub_rn(bb[q], mm[q]);
              acc0 = fmaf(da * da, inv, acc0);
              acc1 = fmaf(db * db, inv, acc1);
              if (r0 == 0) ld += logf(v);
            }
          }
        } else {
          for (int d = dlo + lane; d < dhi; d += 32) {
            const float xva = has0 ? __ldg(xa + d) : 0.f;
            const float xvb = has1 ? __ldg(xb + d) : 0.f;
            const float mu_ = mrow[d];
            const float v = fmaxf(__fadd_rn(vrow[d], p.eps), 1e-8f);
            const float inv = rcp_rn_normal(v);
            const float da = __fsub_rn(xva, mu_), db = __fsub_rn(xvb, mu_);
            acc0 = fmaf(da * da, inv, acc0);
            acc1 = fmaf(db * db, inv, acc1);
            if (r0 == 0) ld += logf(v);
          }
        }
        acc0 = warp_sum(acc0), acc1 = warp_sum(acc1);
        if (r0 == 0) ld = warp_sum(ld);
        if (lane == 0) {
          s_part[slot * 3 + 0] = acc0;
          s_part[slot * 3 + 1] = acc1;
          s_part[slot * 3 + 2] = ld;
        }
      }
      __syncthreads();
      if (tid < 3 * M) {
        const int which = tid / M, m = tid - which * M;
        float t = 0.f;
        for (int part = 0; part < wpm; ++part) t += s_part[(m * wpm + part) * 3 + which];
        if (which == 0 && has0) s_ll[r0 * kMaxM + m] = t;
        if (which == 1 && has1) s_ll[(r0 + 1) * kMaxM + m] = t;
        if (which == 2 && r0 == 0) s_logdet[m] = t;
      }
      __syncthreads();
    }

    // ---- phase 2: log joint, predict logits, responsibilities ----------------------------------------------
    if (tid < M) s_logpi[tid] = logf(__fadd_rn(cur_pi, 1e-10f));
    __syncthreads();
    for (int r = tid; r < rows; r += T) {
      float lj[MM];
      float mx = -INFINITY;
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        if (m < M) {
          const float ll = __fmul_rn(-0.5f, __fadd_rn(s_logdet[m], s_ll[r * kMaxM + m]));
          lj[m] = __fadd_rn(s_logpi[m], ll);
          mx = fmaxf(mx, lj[m]);
        }
      }
      float se = 0.f;
#pragma unroll
      for (int m = 0; m < MM; ++m)
        if (m < M) se += expf(lj[m] - mx);
      const float lse = __fadd_rn(logf(se), mx);
      if (r < p.Bp) {
        p.out_logits[((size_t)s * p.Bp + r) * p.ldo + p.ko_off + k] = lse;
      } else {
        // r = exp(log_joint - logsumexp(log_joint)) exactly as dota_mixture.py:182-183 writes it (NOT the softmax
        // quotient): the log joints are O(1e3), so the rounding of lse is part of the reference's responsibilities
        const int b = r - p.Bp;
        const float gc = s_gc[b];
#pragma unroll
        for (int m = 0; m < MM; ++m)
          if (m < M) s_gamma[b * kMaxM + m] = __fmul_rn(gc, expf(__fsub_rn(lj[m], lse)));
      }
    }
    __syncthreads();

    if (p.B > 0) {
      // ---- phase 3: soft counts ----------------------------------------------------------------------------
      if (tid < M) {
        float sg = 0.f;
        for (int b = 0; b < p.B; ++b) sg += s_gamma[b * kMaxM + tid];
        const float cold = cur_c;
        const float cnew = __fadd_rn(cold, sg);
        s_sumg[tid] = sg;
        s_cold[tid] = cold;
        s_denom[tid] = rcp_rn_normal(__fadd_rn(cnew, 1e-10f));  // reciprocal of (c_new + 1e-10)
        s_part[tid] = cnew;
      }
      __syncthreads();
      if (tid < M) {
        float ck = 0.f;
        for (int m = 0; m < M; ++m) ck += s_part[m];
        p.c[(size_t)item * M + tid] = s_part[tid];
        p.pi[(size_t)item * M + tid] = __fdiv_rn(s_part[tid], __fadd_rn(ck, 1e-10f));
      }
      if (tid == 32) {
        float gsum = 0.f;
        for (int b = 0; b < p.B; ++b) gsum += s_gc[b];
        p.class_counts[item] = cur_cc + gsum;
      }

      // ---- phase 4: M-step, in place in shared memory (same warp -> (mode, d-range) slots) -----------------
      const float* xf = p.x_fit + (size_t)s * p.B * D;
      for (int slot = warp; slot < slots; slot += nwarps) {
        const int m = slot / wpm, part = slot - m * wpm;
        const int dlo = part * Cd, dhi = min(D, dlo + Cd);
        const float cold = s_cold[m], sg = s_sumg[m], rden = s_denom[m], g0 = s_gamma[m];
        float* mrow = t_mu + m * D;
        float* vrow = t_var + m * D;
        auto update = [&](float mu_, float var_, float wx, float wxsq, float& mu_new, float& var_new) {
          mu_new = __fmul_rn(__fadd_rn(__fmul_rn(cold, mu_), wx), rden);
          const float term2 = __fmul_rn(__fmul_rn(-2.0f, mu_), wx);
          const float term3 = __fmul_rn(sg, __fmul_rn(mu_, mu_));
          const float wsd = __fadd_rn(__fadd_rn(wxsq, term2), term3);
          var_new = fmaxf(__fmul_rn(__fadd_rn(__fmul_rn(cold, var_), wsd), rden), 1e-8f);
        };
        if (vec4 && p.B == 1) {
#pragma unroll 2
          for (int d = dlo + 4 * lane; d < dhi; d += 128) {
            float4 m4 = *reinterpret_cast<const float4*>(mrow + d);
            float4 v4 = *reinterpret_cast<const float4*>(vrow + d);
            const float4 x4 = __ldg(reinterpret_cast<const float4*>(xf + d));
            update(m4.x, v4.x, __fmul_rn(g0, x4.x), __fmul_rn(g0, __fmul_rn(x4.x, x4.x)), m4.x, v4.x);
            update(m4.y, v4.y, __fmul_rn(g0, x4.y), __fmul_rn(g0, __fmul_rn(x4.y, x4.y)), m4.y, v4.y);
            update(m4.z, v4.z, __fmul_rn(g0, x4.z), __fmul_rn(g0, __fmul_rn(x4.z, x4.z)), m4.z, v4.z);
            update(m4.w, v4.w, __fmul_rn(g0, x4.w), __fmul_rn(g0, __fmul_rn(x4.w, x4.w)), m4.w, v4.w);
            *reinterpret_cast<float4*>(mrow + d) = m4;
            *reinterpret_cast<float4*>(vrow + d) = v4;
          }
        } else {
          for (int d = dlo + lane; d < dhi; d += 32) {
            const float x0 = __ldg(xf + d);
            float wx = __fmul_rn(g0, x0);
            float wxsq = __fmul_rn(g0, __fmul_rn(x0, x0));
            for (int b = 1; b < p.B; ++b) {
              const float xv = __ldg(xf + (size_t)b * D + d), gb = s_gamma[b * kMaxM + m];
              wx = __fmaf_rn(gb, xv, wx);
              wxsq = __fmaf_rn(gb, __fmul_rn(xv, xv), wxsq);
            }
            float mu_new, var_new;
            update(mrow[d], vrow[d], wx, wxsq, mu_new, var_new);
            mrow[d] = mu_new;
            vrow[d] = var_new;
          }
        }
      }
      // ---- write back --------------------------------------------------------------------------------------
      if (p.use_bulk) {
        fence_proxy_async();  // generic-proxy writes above -> visible to the bulk-copy (async) proxy
        __syncthreads();
        if (tid == 0) {
          bulk_s2g(p.mu + (size_t)item * MD, t_mu, (uint32_t)tile_bytes);
          bulk_s2g(p.var + (size_t)item * MD, t_var, (uint32_t)tile_bytes);
          bulk_commit();
          if (p.stages == 1) bulk_wait_read<0>();
        }
        if (p.stages == 1) __syncthreads();
      } else {
        __syncthreads();
        for (int i = tid; i < MD; i += T) {
          p.mu[(size_t)item * MD + i] = t_mu[i];
          p.var[(size_t)item * MD + i] = t_var[i];
        }
        __syncthreads();
      }
    } else {
      __syncthreads();  // predict only: tile may be overwritten by the next prefetch
    }
    if (p.use_bulk && p.stages == 1) {
      const int nxt = li + gridDim.x;
      if (tid == 0 && nxt < total) issue_load(phys(nxt), 0);
    }
  }
  if (p.use_bulk && tid == 0) bulk_wait<0>();  // stores complete before the CTA retires
}


// ---------------------------------------------------------------------------------------------------------
// Single-sample fast path (Bp <= 1, B <= 1, D % (128*V) == 0): the per-sample step of the reference loop and the
// only shape that is HBM-bound (Objaverse-LVIS: 75.8 MB of state in, 75.8 MB out per fit).
//
// * A warp owns a (mode, 128*V-float chunk) of a class and keeps its mu/var in registers from the likelihood
//   pass to the M-step: shared memory is only the landing zone of the TMA loads (read once), and the updated
//   state goes back with coalesced 512-byte warp stores straight from the registers.
// * The warps of a CTA form G independent groups that work on different classes (class j of the CTA belongs to
//   group j mod G, each group has its own named barrier), so the reduction / responsibility bubble of one group
//   is filled by the streaming phase of the other.
// * The class tiles land in an NS-stage ring shared by the groups; a stage is released as soon as its group has
//   pulled it into registers (one group barrier per class), and the group leader re-arms it NS classes ahead.
// * The mode responsibilities (a handful of scalars) are evaluated redundantly by every warp of the group.
// ---------------------------------------------------------------------------------------------------------
// reductions over the first 8 (M <= 8) or all 32 lanes; lanes beyond M hold the neutral element
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

// threads per CTA the register budget of a (V, G) variant allows
constexpr int b1_max_threads(int V, int G) { return (V <= 2 ? 1024 : (V <= 5 ? 512 : 256)) * (V <= 2 ? 1 : G); }

template <int V, int G, bool PRED, bool FIT, bool LOGP>
__global__ void __launch_bounds__(b1_max_threads(V, G), 1) modedota_b1_kernel(const StepParams p) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int M = p.M, D = p.D, MD = M * D, NS = p.stages;
  const int chunks = D / (128 * V);
  const int gwarps = M * chunks;                               // warps per group
  const int warp = (tid >> 5) % gwarps, grp = (tid >> 5) / gwarps;
  const int wm = warp / chunks, wch = warp - wm * chunks;      // this warp's mode and chunk
  const int d0 = wch * 128 * V + lane * 4;                    // first float of this lane (then every 128 floats)
  const bool SHORTM = M <= 8;
  const uint32_t tile_bytes = (uint32_t)MD * sizeof(float);

  float* s_tiles = reinterpret_cast<float*>(s_raw);                        // [NS][2][MD]
  float* s_part = s_tiles + (size_t)NS * 2 * MD;                           // [G][2][32 warps][4]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_part + G * 2 * 32 * 4);  // [NS]
  int* s_issued = reinterpret_cast<int*>(s_bar + 8);                       // [NS] highest class index armed per stage

  const int total = p.S * p.K;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_mine = first < total ? (total - first + stride - 1) / stride : 0;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&s_bar[i], 1);
      s_issued[i] = i < n_mine ? i : -1;
    }
    fence_mbar_init();
  }
  __syncthreads();

  auto item_of = [&](int j) { return first + j * stride; };   // j-th class of this CTA
  auto issue_load = [&](int j) {   // -> stage j % NS
    const int stage = j % NS;
    const size_t item = (size_t)item_of(j);
    float* dst = s_tiles + (size_t)stage * 2 * MD;
    mbar_expect_tx(&s_bar[stage], 2 * tile_bytes);
    bulk_g2s(dst, p.mu + item * MD, tile_bytes, &s_bar[stage]);
    bulk_g2s(dst + MD, p.var + item * MD, tile_bytes, &s_bar[stage]);
  };
  if (tid == 0)
    for (int j = 0; j < NS && j < n_mine; ++j) issue_load(j);

  // per-class scalars, fetched one class (of this group) ahead: pi/c of mode `lane`, gamma_class, class_counts
  float nx_pi = 0.f, nx_c = 0.f, nx_g = 0.f, nx_cc = 0.f;
  auto fetch_small = [&](int j) {
    const int item = item_of(j);
    if (lane < M) {
      nx_pi = __ldg(p.pi + (size_t)item * M + lane);
      if (FIT) nx_c = __ldg(p.c + (size_t)item * M + lane);
    }
    if (FIT) {
      const int s = item / p.K, k = item - s * p.K;
      nx_g = __ldg(p.gamma + (size_t)s * p.ldg + p.kg_off + k);
      if (warp == 0) nx_cc = __ldg(p.class_counts + item);
    }
  };
  if (grp < n_mine) fetch_small(grp);

  int it = 0;   // group-local iteration (parity of the partial-sum buffer)
  for (int j = grp; j < n_mine; j += G, ++it) {
    const int stage = j % NS;
    const uint32_t parity = (uint32_t)(j / NS) & 1u;
    const int item = item_of(j);
    const int s = item / p.K, k = item - s * p.K;
    const float* t_mu = s_tiles + (size_t)stage * 2 * MD + (size_t)wm * D + d0;
    const float* t_var = t_mu + MD;
    const float cur_pi = nx_pi, cur_c = nx_c, cur_g = nx_g, cur_cc = nx_cc;
    if (j + G < n_mine) fetch_small(j + G);
    const float* xp_row = PRED ? p.x_pred + (size_t)s * D + d0 : nullptr;   // the sample (L1-resident)
    const float* xf_row = FIT ? p.x_fit + (size_t)s * D + d0 : nullptr;

    // The ring is shared by the groups and an mbarrier parity only tells adjacent phases apart: a group may look at a
    // stage only once the load of ITS class has been armed (by the group that drained the stage's previous class);
    // otherwise a group two phases ahead would pass the parity test on the previous class's tile.
    if (lane == 0) {
      while (*reinterpret_cast<volatile int*>(&s_issued[stage]) < j) {
      }
    }
    __syncwarp();
    mbar_wait(&s_bar[stage], parity);

    // ---- likelihood pass: log-determinant and Mahalanobis partial sums of this warp's chunk ----------------
    float4 mu4[V], var4[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      mu4[v] = *reinterpret_cast<const float4*>(t_mu + 128 * v);
      var4[v] = *reinterpret_cast<const float4*>(t_var + 128 * v);
    }
    float accp = 0.f, accf = 0.f, ld = 0.f, mprod = 1.f;
    int esum = 0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float mm[4] = {mu4[v].x, mu4[v].y, mu4[v].z, mu4[v].w};
      const float vv[4] = {var4[v].x, var4[v].y, var4[v].z, var4[v].w};
      float pp[4] = {0.f, 0.f, 0.f, 0.f}, ff[4] = {0.f, 0.f, 0.f, 0.f};
      if (PRED) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(xp_row + 128 * v));
        pp[0] = t.x, pp[1] = t.y, pp[2] = t.z, pp[3] = t.w;
      }
      if (FIT) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(xf_row + 128 * v));
        ff[0] = t.x, ff[1] = t.y, ff[2] = t.z, ff[3] = t.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float vq = fmaxf(__fadd_rn(vv[q], p.eps), 1e-8f);
        const float inv = rcp_rn_normal(vq);   // one correctly-rounded reciprocal shared by both rows
        if (PRED) {
          const float dq = __fsub_rn(pp[q], mm[q]);
          accp = fmaf(dq * dq, inv, accp);
        }
        if (FIT) {
          const float dq = __fsub_rn(ff[q], mm[q]);
          accf = fmaf(dq * dq, inv, accf);
        }
        if (LOGP) {
          // log-determinant in product form: sum_d log v = ln2 * sum_d e_d + log prod_d m_d with v = m * 2^e,
          // m in [1,2). At most 4*V <= 40 mantissas are multiplied (< 2^40), then one logf per lane.
          const uint32_t bits = __float_as_uint(vq);
          esum += (int)(bits >> 23);
          mprod *= __uint_as_float((bits & 0x007fffffu) | 0x3f800000u);
        } else {
          ld += logf(vq);   // 1-ulp logf per element, like the reference (see the note in the general kernel)
        }
      }
    }
    if (LOGP) ld = fmaf((float)(esum - 127 * 4 * V), 0.693147182f, logf(mprod));
    if (PRED) accp = warp_sum(accp);
    if (FIT) accf = warp_sum(accf);
    ld = warp_sum(ld);
    const float logpi = logf(__fadd_rn(cur_pi, 1e-10f));   // off the critical path: before the barrier
    float* part = s_part + ((grp * 2 + (it & 1)) * 32) * 4;
    if (lane == 0) *reinterpret_cast<float4*>(part + warp * 4) = make_float4(accp, accf, ld, 0.f);
    if (G == 1) __syncthreads(); else group_barrier(1 + grp, gwarps * 32);
    // every warp of the group has pulled the stage into registers: re-arm it NS classes ahead
    if (warp == 0 && lane == 0 && j + NS < n_mine) {
      issue_load(j + NS);
      *reinterpret_cast<volatile int*>(&s_issued[stage]) = j + NS;
    }

    // ---- responsibilities: every warp evaluates all M modes (lane = mode; lanes 8.. idle when M <= 8) ------
    float mp = 0.f, mf = 0.f, ldet = 0.f;
    if (lane < M) {
      for (int ch = 0; ch < chunks; ++ch) {
        const float4 q = *reinterpret_cast<const float4*>(part + (lane * chunks + ch) * 4);
        mp += q.x, mf += q.y, ldet += q.z;
      }
    }
    if (PRED) {
      const float lj = lane < M ? __fadd_rn(logpi, __fmul_rn(-0.5f, __fadd_rn(ldet, mp))) : -INFINITY;
      const float mx = SHORTM ? modes_max<true>(lj) : modes_max<false>(lj);
      const float ex = lane < M ? expf(__fsub_rn(lj, mx)) : 0.f;
      const float se = SHORTM ? modes_sum<true>(ex) : modes_sum<false>(ex);
      if (warp == 0 && lane == 0) p.out_logits[(size_t)s * p.ldo + p.ko_off + k] = __fadd_rn(logf(se), mx);
    }
    if (FIT) {
      const float lj = lane < M ? __fadd_rn(logpi, __fmul_rn(-0.5f, __fadd_rn(ldet, mf))) : -INFINITY;
      const float mx = SHORTM ? modes_max<true>(lj) : modes_max<false>(lj);
      const float ex = lane < M ? expf(__fsub_rn(lj, mx)) : 0.f;
      const float se = SHORTM ? modes_sum<true>(ex) : modes_sum<false>(ex);
      const float lse = __fadd_rn(logf(se), mx);
      // gamma[b=0, mode] = gamma_class * exp(log_joint - logsumexp), the reference's form (dota_mixture.py:182-186)
      const float gam = lane < M ? __fmul_rn(cur_g, expf(__fsub_rn(lj, lse))) : 0.f;
      const float cnew = __fadd_rn(cur_c, gam);
      const float rden = rcp_rn_normal(__fadd_rn(cnew, 1e-10f));
      const float cm = lane < M ? cnew : 0.f;
      const float ck = SHORTM ? modes_sum<true>(cm) : modes_sum<false>(cm);
      if (warp == 0) {
        if (lane < M) {
          p.c[(size_t)item * M + lane] = cnew;
          p.pi[(size_t)item * M + lane] = __fdiv_rn(cnew, __fadd_rn(ck, 1e-10f));
        }
        if (lane == 0) p.class_counts[item] = cur_cc + cur_g;
      }
      // ---- M-step on the registers of this warp's (mode, chunk), written straight back to HBM ---------------
      const float cold = __shfl_sync(kFullMask, cur_c, wm);
      const float g0 = __shfl_sync(kFullMask, gam, wm);
      const float rd = __shfl_sync(kFullMask, rden, wm);
      float* o_mu = p.mu + (size_t)item * MD + (size_t)wm * D + d0;
      float* o_var = p.var + (size_t)item * MD + (size_t)wm * D + d0;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float mm[4] = {mu4[v].x, mu4[v].y, mu4[v].z, mu4[v].w};
        float vv[4] = {var4[v].x, var4[v].y, var4[v].z, var4[v].w};
        const float4 t = __ldg(reinterpret_cast<const float4*>(xf_row + 128 * v));
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
        *reinterpret_cast<float4*>(o_mu + 128 * v) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        *reinterpret_cast<float4*>(o_var + 128 * v) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      }
    }
  }
}

template <int V, int G, bool LP>
cudaError_t launch_b1(const StepParams& p, unsigned grid, int threads, size_t smem, cudaStream_t st) {
  cudaError_t e;
#define UA_B1_CASE(PR, FI)                                                                        \
  {                                                                                               \
    auto kern = modedota_b1_kernel<V, G, PR, FI, LP>;                                             \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
    if (e == cudaSuccess) kern<<<grid, threads, smem, st>>>(p);                                   \
  }
  if (p.Bp && p.B) UA_B1_CASE(true, true)
  else if (p.B) UA_B1_CASE(false, true)
  else UA_B1_CASE(true, false)
#undef UA_B1_CASE
  return e;
}

template <int V>
cudaError_t launch_b1_v(const StepParams& p, int groups, bool lp, unsigned grid, int threads, size_t smem,
                        cudaStream_t st) {
  if (groups == 2)
    return lp ? launch_b1<V, 2, true>(p, grid, threads, smem, st) : launch_b1<V, 2, false>(p, grid, threads, smem, st);
  return lp ? launch_b1<V, 1, true>(p, grid, threads, smem, st) : launch_b1<V, 1, false>(p, grid, threads, smem, st);
}

}  // namespace
}  // namespace ua

extern "C" int ua_modedota_step_f32(const float* x_pred, int Bp, const float* x_fit, const float* gamma_class, int B,
                                    int ldg, int k_gamma_offset, float* mu, float* var, float* pi, float* c,
                                    float* class_counts, int S, int K, int M, int D, float eps, float* out_logits,
                                    int ldo, int k_out_offset, void* stream) {
  using namespace ua;
  UA_REQUIRE(mu && var && pi && c, "ua_modedota_step_f32: state pointers must be non-NULL");
  UA_REQUIRE(S >= 1 && K >= 1 && M >= 1 && D >= 1, "ua_modedota_step_f32: bad sizes S=%d K=%d M=%d D=%d", S, K, M, D);
  if (!x_pred) Bp = 0;
  if (!x_fit) B = 0;
  UA_REQUIRE(Bp >= 0 && B >= 0 && Bp + B >= 1, "ua_modedota_step_f32: nothing to do (Bp=%d, B=%d)", Bp, B);
  UA_REQUIRE(Bp == 0 || out_logits, "ua_modedota_step_f32: out_logits is NULL");
  UA_REQUIRE(B == 0 || (gamma_class && class_counts), "ua_modedota_step_f32: gamma_class/class_counts is NULL");
  UA_REQUIRE(B == 0 || ldg >= k_gamma_offset + K, "ua_modedota_step_f32: ldg=%d < offset+K", ldg);
  UA_REQUIRE(Bp == 0 || ldo >= k_out_offset + K, "ua_modedota_step_f32: ldo=%d < offset+K", ldo);
  UA_UNSUPPORTED(M > kMaxM, "ua_modedota_step_f32: M=%d > %d", M, kMaxM);
  UA_UNSUPPORTED((long long)S * K > 0x3fffffffLL, "ua_modedota_step_f32: S*K too large");
  UA_UNSUPPORTED(Bp + B > kMaxRows, "ua_modedota_step_f32: Bp+B=%d > %d rows per launch", Bp + B, kMaxRows);

  StepParams p;
  p.x_pred = x_pred, p.x_fit = x_fit, p.gamma = gamma_class;
  p.mu = mu, p.var = var, p.pi = pi, p.c = c, p.class_counts = class_counts, p.out_logits = out_logits;
  p.S = S, p.Bp = Bp, p.B = B, p.K = K, p.M = M, p.D = D;
  p.ldg = ldg, p.kg_off = k_gamma_offset, p.ldo = ldo, p.ko_off = k_out_offset, p.eps = eps;

  const size_t tile_bytes = (size_t)M * D * sizeof(float);
  const size_t fixed = ((size_t)2 * kMaxRows * kMaxM + 32 * 3 * kMaxM + 8 * kMaxM + kMaxRows) * sizeof(float) + 16;
  const size_t budget = 227 * 1024;
  p.use_bulk = (tile_bytes % 16 == 0) && ((uintptr_t)mu % 16 == 0) && ((uintptr_t)var % 16 == 0);
  p.vec_ok = (D % 4 == 0) && ((uintptr_t)x_pred % 16 == 0) && ((uintptr_t)x_fit % 16 == 0);
  p.stages = (4 * tile_bytes + fixed <= budget) ? 2 : 1;
  const size_t smem = (size_t)p.stages * 2 * tile_bytes + fixed;
  UA_UNSUPPORTED(smem > budget, "ua_modedota_step_f32: M*D=%d does not fit in shared memory", M * D);
  if (!p.use_bulk) p.stages = 1;

  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)S * K;

  // ---- single-sample fast path -----------------------------------------------------------------------------
  if (Bp <= 1 && B <= 1 && p.use_bulk && p.vec_ok && D % 128 == 0 && g_modedota_v >= 0) {
    static const int kV[] = {1, 2, 4, 5, 8, 10};
    // widest lanes first for two groups (more ILP per thread, fewer warps per barrier), else the narrowest fit
    int V = 0, groups = 1;
    const int want_g = g_modedota_groups > 0 ? g_modedota_groups : 2;
    for (int gtry = want_g; gtry >= 1 && !V; --gtry) {
      for (int i = 5; i >= 0 && !V; --i) {
        const int cand = kV[i];
        if (g_modedota_v > 0 && cand != g_modedota_v) continue;
        if (D % (128 * cand)) continue;
        const int thr = gtry * 32 * M * (D / (128 * cand));
        if (thr > b1_max_threads(cand, gtry) || thr > 1024) continue;
        if (g_modedota_v == 0 && gtry == 2 && thr < 256) continue;   // too few warps to hide anything
        V = cand, groups = gtry;
      }
    }
    const size_t small = (size_t)2 * 2 * 32 * 4 * sizeof(float) + 8 * sizeof(uint64_t) + 8 * sizeof(int);
    int ns = 0;
    for (int cand = 4; cand >= 2 && !ns; --cand)
      if ((size_t)cand * 2 * tile_bytes + small <= budget && (cand <= 3 || 4 * 2 * tile_bytes <= 100 * 1024)) ns = cand;
    if (V && ns) {
      const int thr = groups * 32 * M * (D / (128 * V));
      const size_t smem_b1 = (size_t)ns * 2 * tile_bytes + small;
      int per = (int)(budget / (smem_b1 + 1024));
      if (per > 2048 / thr) per = 2048 / thr;
      if (per < 1) per = 1;
      if (per > 4) per = 4;
      long long g = (long long)kNumSMs * per;
      if (g > total) g = total;
      p.stages = ns;
      cudaError_t e = cudaSuccess;
      const bool lp = g_modedota_logprod != 0;
      switch (V) {
        case 1: e = launch_b1_v<1>(p, groups, lp, (unsigned)g, thr, smem_b1, st); break;
        case 2: e = launch_b1_v<2>(p, groups, lp, (unsigned)g, thr, smem_b1, st); break;
        case 4: e = launch_b1_v<4>(p, groups, lp, (unsigned)g, thr, smem_b1, st); break;
        case 5: e = launch_b1_v<5>(p, groups, lp, (unsigned)g, thr, smem_b1, st); break;
        case 8: e = launch_b1_v<8>(p, groups, lp, (unsigned)g, thr, smem_b1, st); break;
        default: e = launch_b1_v<10>(p, groups, lp, (unsigned)g, thr, smem_b1, st); break;
      }
      if (e != cudaSuccess) {
        set_error("ua_modedota_step_f32: cudaFuncSetAttribute(%zu B): %s", smem_b1, cudaGetErrorString(e));
        return UA_ERR_CUDA;
      }
      return check_launch("ua_modedota_step_f32(b1)");
    }
  }

  // ---- batched path: a cluster per class, feature axis split across its CTAs (modedota_batch.cu) ---------------
  {
    const int rc = modedota_batch_launch(p, st);
    if (rc != 1) return rc;
  }

  int threads = g_modedota_threads > 0 ? g_modedota_threads : (D >= 1024 ? 1024 : 512);
  if (threads > (M > 8 ? 512 : 1024)) threads = M > 8 ? 512 : 1024;
  if (threads < 256) threads = 256;
  // CTAs per SM allowed by shared memory; persistent grid = that many waves' worth of CTAs at most
  int per_sm = (int)(budget / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  long long grid = (long long)kNumSMs * per_sm;
  if (grid > total) grid = total;

  cudaError_t e;
  if (M <= 4) {
    auto kern = modedota_step_kernel<4>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) kern<<<(unsigned)grid, threads, smem, st>>>(p);
  } else if (M <= 8) {
    auto kern = modedota_step_kernel<8>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) kern<<<(unsigned)grid, threads, smem, st>>>(p);
  } else {
    auto kern = modedota_step_kernel<16>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) kern<<<(unsigned)grid, threads, smem, st>>>(p);
  }
  if (e != cudaSuccess) {
    set_error("ua_modedota_step_f32: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    return UA_ERR_CUDA;
  }
  return check_launch("ua_modedota_step_f32");
}
