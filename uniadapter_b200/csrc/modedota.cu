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

namespace ua {

int g_modedota_threads = 0;  // tuning override

namespace {

constexpr int kMaxM = 16;
constexpr int kMaxRows = 160;  // Bp + B rows of log-likelihoods kept in shared memory

struct StepParams {
  const float* x_pred;  // [S,Bp,D] or null
  const float* x_fit;   // [S,B,D] or null
  const float* gamma;   // [S,B,ldg]
  float* mu;
  float* var;
  float* pi;
  float* c;
  float* class_counts;
  float* out_logits;  // [S,Bp,ldo]
  int S, Bp, B, K, M, D;
  int ldg, kg_off, ldo, ko_off;
  float eps;
  int use_bulk, stages, vec_ok;
};

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

  int item = blockIdx.x;
  if (p.use_bulk && tid == 0 && item < total) issue_load(item, 0);

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
  if (item < total) fetch_small(item);

  uint32_t phase_bits = 0;  // per-stage mbarrier parity
  int it = 0;
  for (; item < total; item += gridDim.x, ++it) {
    const int stage = p.stages == 2 ? (it & 1) : 0;
    const int s = item / p.K, k = item - s * p.K;
    float* t_mu = s_tiles + (size_t)stage * 2 * MD;
    float* t_var = t_mu + MD;
    const float cur_pi = nx_pi, cur_c = nx_c, cur_cc = nx_cc;
    if (tid >= 64 && tid < 64 + p.B) s_gc[tid - 64] = nx_g;   // read after the barriers below
    if (item + gridDim.x < total) fetch_small(item + gridDim.x);

    if (p.use_bulk) {
      if (tid == 0) {
        const int nxt = item + gridDim.x;
        if (p.stages == 2 && nxt < total) {
          bulk_wait_read<0>();  // the store that last read the other stage has drained its shared-memory reads
          issue_load(nxt, stage ^ 1);
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
              const float inv = __frcp_rn(v);
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
            const float inv = __frcp_rn(v);
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
        s_denom[tid] = __frcp_rn(__fadd_rn(cnew, 1e-10f));  // reciprocal of (c_new + 1e-10)
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
      const int nxt = item + gridDim.x;
      if (tid == 0 && nxt < total) issue_load(nxt, 0);
    }
  }
  if (p.use_bulk && tid == 0) bulk_wait<0>();  // stores complete before the CTA retires
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

  const long long total = (long long)S * K;
  int threads = g_modedota_threads > 0 ? g_modedota_threads : (D >= 1024 ? 1024 : 512);
  if (threads > (M > 8 ? 512 : 1024)) threads = M > 8 ? 512 : 1024;
  if (threads < 256) threads = 256;
  // CTAs per SM allowed by shared memory; persistent grid = that many waves' worth of CTAs at most
  int per_sm = (int)(budget / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  long long grid = (long long)kNumSMs * per_sm;
  if (grid > total) grid = total;

  cudaStream_t st = (cudaStream_t)stream;
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
