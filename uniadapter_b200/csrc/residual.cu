// Residual text-feature learning for sm_100a: the inner loop of Uni_Adapter.py:443-476 (10 Adam steps on the per-class
// text residuals against the MODE-DOTA likelihood matrix, compute_text_alignment_loss, Uni_Adapter.py:191-270) with a
// hand-derived backward instead of autograd. S independent streams per launch.
//
//   E = text0 + R;  n_i = |E_i|;  X_i = E_i / n_i
//   lj[i,k,m] = log(pi_km + 1e-10) - 0.5 * (sum_d log v_kmd + sum_d (X_id - mu_kmd)^2 / v_kmd),  v = max(var + eps, 1e-8)
//   LM[i,k]   = logsumexp_m lj;   Z = LM / max(LM);   P = exp(exp(Z))
//   loss      = -mean_i P_ii / sum_k P_ik - mean_i P_ii / sum_j P_ji
//   backward:  dP -> dZ = dP * P * exp(Z) -> dLM = dZ / max  (and -sum(dZ * LM) / max^2 at the arg-max, as autograd does)
//              W[i,k,m] = dLM[i,k] * softmax_m(lj)[i,k,m];   dX[i,d] = -sum_km W[i,km] (X_id - mu_kmd) / v_kmd
//              dE_i = (dX_i - X_i <X_i, dX_i>) / n_i;   Adam (torch.optim.Adam defaults) on R.
//
// Four kernels per Adam step, all fp32 SIMT (the contraction is 40 x 320 x 512 per stream: far too small for the
// tensor pipe to matter, and the direct (x - mu)^2 / v form is the reference's arithmetic):
//   resid_forward   grid (class blocks, S): a block of classes' mu and 1/v tiles live in shared memory, every warp
//                   accumulates a 4-row x 5-column register tile over D (lanes = D), then logsumexp over the modes.
//   resid_loss      grid (S): the K x K matrix in shared memory; loss, dLM, W.
//   resid_backward  grid (D blocks, S): all K*M columns of a 64-wide D slice in shared memory, 5-row register tiles.
//   resid_embed     grid (S*K): normalisation backward + Adam + the next normalised embedding.
// The state (mu, var) of 15 streams is 19.7 MB: it stays in L2 across the 40 launches of a step.
#include "common.cuh"

namespace ua {

int g_resid_cb = 0;    // tuning: classes per CTA of the forward kernel (0 = heuristic)
int g_resid_dbl = 0;   // tuning: floats per lane of the backward kernel (1 or 2; 0 = heuristic)

namespace {

constexpr int kRI = 5;    // rows per register tile (forward)
constexpr int kCJ = 4;    // columns per register tile (forward)
constexpr int kBR = 5;    // rows per register tile (backward)
constexpr int kThreads = 256;
constexpr int kLossThreads = 1024;

__device__ __forceinline__ float block_sum_256(float v, float* s_tmp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) s_tmp[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < W; ++w) t += s_tmp[w];
  return t;
}

// ---- constants of the cache that do not change during the 10 steps: log-determinant and log-weight per (k,m) ----
__global__ void __launch_bounds__(kThreads)
    resid_prep_kernel(const float* __restrict__ var, const float* __restrict__ pi, int rows, int D, float eps,
                      float* __restrict__ consts, float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* v = var + (size_t)row * D;
  float ld = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float vv = fmaxf(__fadd_rn(v[d], eps), 1e-8f);
    ld += logf(vv);
    inv[(size_t)row * D + d] = __fdiv_rn(1.0f, vv);   // the cache does not change during the 10 steps: one division per
  }                                                    // element and call instead of one per element and launch
  ld = warp_sum(ld);
  if (lane == 0) {
    consts[2 * row + 0] = ld;
    consts[2 * row + 1] = logf(__fadd_rn(pi[row], 1e-10f));
  }
}

// ---- normalisation backward + Adam + next embedding ------------------------------------------------------------
struct EmbedParams {
  const float* text0;
  long long text0_stride;   // 0: one text matrix for all streams
  float* residual;          // [S,K,D] (read-only when adam == 0)
  float* adam_m;
  float* adam_v;
  const int* adam_t;        // [S] steps taken before this call
  int t_index;              // 1-based step inside this call
  double lr, beta1, beta2, adam_eps;
  const float* dX;          // [S,K,D]
  const float* rdot;        // [S,K,NB]
  int NB;
  float* X;                 // [S,K,D] out: normalised embedding
  float* nrm;               // [S,K]
  float* out_grad;          // [S,K,D] or null: gradient w.r.t. the residual
  int K, D;
  int mode;                 // 0: embed only, 1: gradient (+Adam if adam_m), then embed
};

__global__ void __launch_bounds__(128) resid_embed_kernel(const EmbedParams p) {
  __shared__ float s_tmp[4];
  __shared__ float s_bc[2];
  const int row = blockIdx.x, s = row / p.K, i = row - s * p.K;
  const int D = p.D, tid = threadIdx.x;
  const float* t0 = p.text0 + (size_t)s * p.text0_stride + (size_t)i * D;
  float* r = p.residual + (size_t)row * D;
  float* X = p.X + (size_t)row * D;
  if (p.mode == 1) {
    float rowdot = 0.f;
    for (int b = 0; b < p.NB; ++b) rowdot += p.rdot[(size_t)row * p.NB + b];
    const float inv_n = __fdiv_rn(1.0f, p.nrm[row]);
    const float* dX = p.dX + (size_t)row * D;
    if (p.adam_m) {
      if (tid == 0) {   // torch.optim.Adam: bias corrections in double, like the Python floats of the reference
        const double t = (double)(p.adam_t[s] + p.t_index);
        s_bc[0] = (float)(p.lr / (1.0 - pow(p.beta1, t)));          // step size
        s_bc[1] = (float)sqrt(1.0 - pow(p.beta2, t));               // sqrt(bias_correction2)
      }
      __syncthreads();
    }
    float* am = p.adam_m ? p.adam_m + (size_t)row * D : nullptr;
    float* av = p.adam_v ? p.adam_v + (size_t)row * D : nullptr;
    const float w1 = (float)(1.0 - p.beta1), w2 = (float)(1.0 - p.beta2), b2 = (float)p.beta2, aeps = (float)p.adam_eps;
    for (int d = tid; d < D; d += 128) {
      const float g = __fmul_rn(__fsub_rn(dX[d], __fmul_rn(X[d], rowdot)), inv_n);
      if (p.out_grad) p.out_grad[(size_t)row * D + d] = g;
      if (am) {
        const float m = fmaf(__fsub_rn(g, am[d]), w1, am[d]);                       // lerp_
        const float v = fmaf(__fmul_rn(g, g), w2, __fmul_rn(av[d], b2));              // mul_ + addcmul_
        am[d] = m, av[d] = v;
        const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), s_bc[1]), aeps);
        r[d] = __fsub_rn(r[d], __fmul_rn(s_bc[0], __fdiv_rn(m, denom)));                        // addcdiv_
      }
    }
    __syncthreads();
  }
  float ss = 0.f;
  for (int d = tid; d < D; d += 128) {
    const float e = __fadd_rn(t0[d], r[d]);
    ss = fmaf(e, e, ss);
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) s_tmp[tid >> 5] = ss;
  __syncthreads();
  const float n = sqrtf(s_tmp[0] + s_tmp[1] + s_tmp[2] + s_tmp[3]);
  for (int d = tid; d < D; d += 128) X[d] = __fdiv_rn(__fadd_rn(t0[d], r[d]), n);
  if (tid == 0) p.nrm[row] = n;
}

__global__ void resid_bump_t_kernel(int* adam_t, int S, int iters) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < S) adam_t[s] += iters;
}

// ---- forward: LM[i,k] and the mode softmax weights ---------------------------------------------------------------
struct FwdParams {
  const float* X;        // [S,K,D]
  const float* mu;       // [S,K,M,D]
  const float* var;      // [S,K,M,D]: 1 / max(var + eps, 1e-8), precomputed once per call (resid_prep)
  const float* consts;   // [S,K,M,2]
  float* LM;             // [S,K,K]
  float* Wt;             // [S,K,K,M]: softmax over modes of lj (row i, class k)
  int K, M, D, CB;       // CB classes per CTA
  float eps;
};

__global__ void __launch_bounds__(kThreads, 3) resid_forward_kernel(const FwdParams p) {
  extern __shared__ __align__(16) float s_f[];
  const int K = p.K, M = p.M, D = p.D;
  const int s = blockIdx.y, k0 = blockIdx.x * p.CB;
  const int ncls = min(p.CB, K - k0), ncols = ncls * M, maxcols = p.CB * M;
  float* t_mu = s_f;                                  // [maxcols][D]
  float* t_iv = t_mu + (size_t)maxcols * D;           // [maxcols][D]
  float* s_lj = t_iv + (size_t)maxcols * D;           // [K][maxcols]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  const float* gmu = p.mu + ((size_t)s * K + k0) * M * D;
  const float* gvar = p.var + ((size_t)s * K + k0) * M * D;
  for (int idx = tid * 4; idx < ncols * D; idx += kThreads * 4) {
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(gmu + idx));
    const float4 i4 = __ldg(reinterpret_cast<const float4*>(gvar + idx));     // 1 / v, precomputed by resid_prep
    *reinterpret_cast<float4*>(t_mu + idx) = m4;
    *reinterpret_cast<float4*>(t_iv + idx) = i4;
  }
  __syncthreads();

  const float* Xs = p.X + (size_t)s * K * D;
  const float* cst = p.consts + ((size_t)s * K + k0) * M * 2;
  const int ncg = (ncols + kCJ - 1) / kCJ, nrg = (K + kRI - 1) / kRI;
  // (column group, row group) register tiles are dealt round-robin to the warps: a CTA holds only a few classes, so
  // that several CTAs are resident per SM and hide each other's shared-memory / L2 latency
  for (int item = warp; item < ncg * nrg; item += kThreads / 32) {
    const int cg = item / nrg, rg = item - cg * nrg;
    int col[kCJ];
#pragma unroll
    for (int c = 0; c < kCJ; ++c) col[c] = min(cg * kCJ + c, ncols - 1);
    {
      int rowi[kRI];
#pragma unroll
      for (int r = 0; r < kRI; ++r) rowi[r] = min(rg * kRI + r, K - 1);
      float acc[kRI][kCJ];
#pragma unroll
      for (int r = 0; r < kRI; ++r)
#pragma unroll
        for (int c = 0; c < kCJ; ++c) acc[r][c] = 0.f;
      for (int d = lane * 4; d < D; d += 128) {
        float4 x4[kRI];
#pragma unroll
        for (int r = 0; r < kRI; ++r) x4[r] = __ldg(reinterpret_cast<const float4*>(Xs + (size_t)rowi[r] * D + d));
#pragma unroll
        for (int c = 0; c < kCJ; ++c) {
          const float4 m4 = *reinterpret_cast<const float4*>(t_mu + (size_t)col[c] * D + d);
          const float4 i4 = *reinterpret_cast<const float4*>(t_iv + (size_t)col[c] * D + d);
#pragma unroll
          for (int r = 0; r < kRI; ++r) {
            float df;
            df = x4[r].x - m4.x, acc[r][c] = fmaf(df * df, i4.x, acc[r][c]);
            df = x4[r].y - m4.y, acc[r][c] = fmaf(df * df, i4.y, acc[r][c]);
            df = x4[r].z - m4.z, acc[r][c] = fmaf(df * df, i4.z, acc[r][c]);
            df = x4[r].w - m4.w, acc[r][c] = fmaf(df * df, i4.w, acc[r][c]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kRI; ++r)
#pragma unroll
        for (int c = 0; c < kCJ; ++c) {
          const float maha = warp_sum(acc[r][c]);
          if (lane == 0 && rg * kRI + r < K && cg * kCJ + c < ncols) {
            const float ld = cst[2 * col[c] + 0], lp = cst[2 * col[c] + 1];
            s_lj[rowi[r] * maxcols + col[c]] = __fadd_rn(lp, __fmul_rn(-0.5f, __fadd_rn(ld, maha)));
          }
        }
    }
  }
  __syncthreads();
  // logsumexp over the modes and the softmax weights (kept for the backward pass)
  for (int e = tid; e < K * ncls; e += kThreads) {
    const int i = e / ncls, kk = e - i * ncls;
    const float* lj = s_lj + i * maxcols + kk * M;
    float mx = -INFINITY;
    for (int m = 0; m < M; ++m) mx = fmaxf(mx, lj[m]);
    float se = 0.f;
    for (int m = 0; m < M; ++m) se += expf(lj[m] - mx);
    const float lse = __fadd_rn(logf(se), mx);
    p.LM[((size_t)s * K + i) * K + k0 + kk] = lse;
    float* w = p.Wt + (((size_t)s * K + i) * K + k0 + kk) * M;
    for (int m = 0; m < M; ++m) w[m] = expf(lj[m] - lse);
  }
}

// ---- loss and its gradient w.r.t. the likelihood matrix; W <- dLM * softmax weights ----------------------------
// WHERE = 0: LM and P = exp(exp(LM / max)) both in shared memory (K <= ~160); 1: P in shared memory, LM read from global
// memory (L2) each time (K <= ~230); 2: P in a global scratch array as well (any K: one CTA per stream walks K*K elements,
// which is small next to the K*K*M*D contraction of the forward / backward kernels around it).
template <int WHERE>
__global__ void __launch_bounds__(kLossThreads) resid_loss_kernel(const float* __restrict__ LM, float* __restrict__ Wt,
                                                              float* __restrict__ Pg, int K, int M,
                                                              float* __restrict__ out_loss, int loss_stride, int loss_index) {
  extern __shared__ __align__(16) float s_l[];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int KK = K * K;
  const float* lm = LM + (size_t)s * KK;
  float* s_s1 = s_l;                 // [K] row sums
  float* s_s2 = s_s1 + K;            // [K] column sums
  float* s_dg = s_s2 + K;            // [K] diagonal of P
  float* s_p = WHERE == 2 ? Pg + (size_t)s * KK : s_dg + K;          // [K*K]
  float* s_lmw = WHERE == 0 ? s_dg + K + KK : nullptr;                // [K*K] (WHERE == 0 only)
  const float* s_lm = WHERE == 0 ? s_lmw : lm;
  __shared__ float s_tmp[32];
  __shared__ float s_best[32];
  __shared__ int s_besti[32];
  constexpr int kT = kLossThreads;   // one CTA per stream walks K*K elements in a chain of dependent phases: 1024 threads
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int e = tid; e < KK; e += kT) {
    const float v = lm[e];
    if (WHERE == 0) s_lmw[e] = v;
    if (v > best) best = v, besti = e;   // ascending e per thread: keeps the first maximum
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(kFullMask, best, o);
    const int oi = __shfl_xor_sync(kFullMask, besti, o);
    if (ob > best || (ob == best && oi < besti)) best = ob, besti = oi;
  }
  if (lane == 0) s_best[warp] = best, s_besti[warp] = besti;
  __syncthreads();
  best = s_best[0], besti = s_besti[0];
  for (int w = 1; w < kT / 32; ++w)
    if (s_best[w] > best || (s_best[w] == best && s_besti[w] < besti)) best = s_best[w], besti = s_besti[w];
  const float lmax = best;
  for (int e = tid; e < KK; e += kT) s_p[e] = expf(expf(__fdiv_rn(s_lm[e], lmax)));
  __syncthreads();
  for (int i = warp; i < K; i += kT / 32) {   // row sums
    float t = 0.f;
    for (int k = lane; k < K; k += 32) t += s_p[i * K + k];
    t = warp_sum(t);
    if (lane == 0) s_s1[i] = t;
  }
  for (int k = tid; k < K; k += kT) {         // column sums, diagonal
    float t = 0.f;
    for (int i = 0; i < K; ++i) t += s_p[i * K + k];
    s_s2[k] = t;
    s_dg[k] = s_p[k * K + k];
  }
  __syncthreads();
  const float invK = __fdiv_rn(1.0f, (float)K);
  if (out_loss) {
    float t = 0.f;
    for (int i = tid; i < K; i += kT) {
      t += __fdiv_rn(s_dg[i], s_s1[i]) + __fdiv_rn(s_dg[i], s_s2[i]);
    }
    t = block_sum_256(t, s_tmp);
    if (tid == 0) out_loss[(size_t)s * loss_stride + loss_index] = -t * invK;
  }
  // dP[i,k] = (diag_i / s1_i^2 + diag_k / s2_k^2) / K  -  [i == k] (1/s1_i + 1/s2_i) / K
  float tot = 0.f;
  for (int e = tid; e < KK; e += kT) {
    const int i = e / K, k = e - i * K;
    const float di = s_dg[i], dk = s_dg[k];
    float dP = (__fdiv_rn(di, s_s1[i] * s_s1[i]) + __fdiv_rn(dk, s_s2[k] * s_s2[k])) * invK;
    if (i == k) dP -= (__fdiv_rn(1.0f, s_s1[i]) + __fdiv_rn(1.0f, s_s2[i])) * invK;
    const float lme = s_lm[e];
    const float q = expf(__fdiv_rn(lme, lmax));
    const float dZ = dP * s_p[e] * q;
    tot = fmaf(dZ, lme, tot);
    s_p[e] = __fdiv_rn(dZ, lmax);   // dLM without the arg-max term
  }
  tot = block_sum_256(tot, s_tmp);
  __syncthreads();
  if (tid == 0) s_p[besti] -= __fdiv_rn(tot, lmax * lmax);   // gradient through max(): lands on the arg-max
  __syncthreads();
  float* w = Wt + (size_t)s * KK * M;
  for (int e = tid; e < KK * M; e += kT) w[e] *= s_p[e / M];   // (K*K*M < 2^31 is checked on the host: 32-bit division)
}

// ---- backward: dX and the row dots of the normalisation backward ---------------------------------------------------
struct BwdParams {
  const float* X;
  const float* mu;
  const float* var;
  const float* Wt;     // [S,K,K*M]: dLM * softmax weights
  float* dX;           // [S,K,D]
  float* rdot;         // [S,K,NB]
  int K, M, D, NB;
  float eps;
};

template <int DBL>   // floats per lane along D; the CTA owns a 32*DBL-wide slice of D
__global__ void __launch_bounds__(kThreads, DBL == 1 ? 2 : 1) resid_backward_kernel(const BwdParams p) {
  extern __shared__ __align__(16) float s_b[];
  constexpr int DB = 32 * DBL;
  const int K = p.K, M = p.M, D = p.D, ncols = K * M;
  const int s = blockIdx.y, dbase = blockIdx.x * DB;
  float* t_mu = s_b;                            // [ncols][DB]
  float* t_iv = t_mu + (size_t)ncols * DB;      // [ncols][DB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* gmu = p.mu + (size_t)s * ncols * D + dbase;
  const float* gvar = p.var + (size_t)s * ncols * D + dbase;
  for (int idx = tid; idx < ncols * DB; idx += kThreads) {
    const int col = idx / DB, d = idx - col * DB;
    t_mu[idx] = __ldg(gmu + (size_t)col * D + d);
    t_iv[idx] = __ldg(gvar + (size_t)col * D + d);          // 1 / v, precomputed by resid_prep
  }
  __syncthreads();
  const int nrg = (K + kBR - 1) / kBR;
  for (int rg = warp; rg < nrg; rg += kThreads / 32) {
    int rowi[kBR];
#pragma unroll
    for (int r = 0; r < kBR; ++r) rowi[r] = min(rg * kBR + r, K - 1);
    float x[kBR][DBL], acc[kBR][DBL];
#pragma unroll
    for (int r = 0; r < kBR; ++r)
#pragma unroll
      for (int e = 0; e < DBL; ++e) {
        x[r][e] = __ldg(p.X + ((size_t)s * K + rowi[r]) * D + dbase + lane * DBL + e);
        acc[r][e] = 0.f;
      }
    const float* wrow[kBR];
#pragma unroll
    for (int r = 0; r < kBR; ++r) wrow[r] = p.Wt + ((size_t)s * K + rowi[r]) * ncols;
    for (int c4 = 0; c4 < ncols; c4 += 4) {   // K*M is a multiple of 4 whenever M is (checked on the host)
      float4 w4[kBR];
#pragma unroll
      for (int r = 0; r < kBR; ++r) w4[r] = __ldg(reinterpret_cast<const float4*>(wrow[r] + c4));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float mq[DBL], iq[DBL];
#pragma unroll
        for (int e = 0; e < DBL; ++e) {
          mq[e] = t_mu[(size_t)(c4 + q) * DB + lane * DBL + e];
          iq[e] = t_iv[(size_t)(c4 + q) * DB + lane * DBL + e];
        }
#pragma unroll
        for (int r = 0; r < kBR; ++r) {
          const float w = q == 0 ? w4[r].x : (q == 1 ? w4[r].y : (q == 2 ? w4[r].z : w4[r].w));
#pragma unroll
          for (int e = 0; e < DBL; ++e) acc[r][e] = fmaf(w, (x[r][e] - mq[e]) * iq[e], acc[r][e]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kBR; ++r) {
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < DBL; ++e) {
        const float dx = -acc[r][e];
        dot = fmaf(x[r][e], dx, dot);
        if (rg * kBR + r < K) p.dX[((size_t)s * K + rowi[r]) * D + dbase + lane * DBL + e] = dx;
      }
      dot = warp_sum(dot);
      if (lane == 0 && rg * kBR + r < K) p.rdot[((size_t)s * K + rowi[r]) * p.NB + blockIdx.x] = dot;
    }
  }
}

// The same contraction when the K*M columns of a D slice do not fit in shared memory (K*M > 880: e.g. OmniObject3D's 216
// classes): grid (D slices, row blocks, S). A warp owns ONE group of kBR rows for the whole launch, so its accumulators stay
// in registers while the CTA walks the columns in chunks of CC, restaging mu and 1/v per chunk (from L2: the state of a
// stream is a few MB). Same arithmetic and summation order over the columns as the kernel above.
template <int DBL>
__global__ void __launch_bounds__(kThreads, 2) resid_backward_chunked_kernel(const BwdParams p, int CC) {
  extern __shared__ __align__(16) float s_b[];
  constexpr int DB = 32 * DBL;
  const int K = p.K, M = p.M, D = p.D, ncols = K * M;
  const int s = blockIdx.z, dbase = blockIdx.x * DB;
  float* t_mu = s_b;                         // [CC][DB]
  float* t_iv = t_mu + (size_t)CC * DB;      // [CC][DB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* gmu = p.mu + (size_t)s * ncols * D + dbase;
  const float* gvar = p.var + (size_t)s * ncols * D + dbase;
  const int rg = blockIdx.y * (kThreads / 32) + warp;
  const bool have = rg * kBR < K;
  int rowi[kBR];
#pragma unroll
  for (int r = 0; r < kBR; ++r) rowi[r] = min(rg * kBR + r, K - 1);
  float x[kBR][DBL], acc[kBR][DBL];
  const float* wrow[kBR];
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    wrow[r] = p.Wt + ((size_t)s * K + rowi[r]) * ncols;
#pragma unroll
    for (int e = 0; e < DBL; ++e) {
      x[r][e] = __ldg(p.X + ((size_t)s * K + rowi[r]) * D + dbase + lane * DBL + e);
      acc[r][e] = 0.f;
    }
  }
  for (int c0 = 0; c0 < ncols; c0 += CC) {
    const int nc = min(CC, ncols - c0);       // a multiple of 4 (CC and K*M are)
    __syncthreads();                          // the previous chunk has been consumed by every warp
    for (int idx = tid; idx < nc * DB; idx += kThreads) {
      const int col = idx / DB, d = idx - col * DB;
      t_mu[idx] = __ldg(gmu + (size_t)(c0 + col) * D + d);
      t_iv[idx] = __ldg(gvar + (size_t)(c0 + col) * D + d);
    }
    __syncthreads();
    if (have) {
      for (int c4 = 0; c4 < nc; c4 += 4) {
        float4 w4[kBR];
#pragma unroll
        for (int r = 0; r < kBR; ++r) w4[r] = __ldg(reinterpret_cast<const float4*>(wrow[r] + c0 + c4));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float mq[DBL], iq[DBL];
#pragma unroll
          for (int e = 0; e < DBL; ++e) {
            mq[e] = t_mu[(size_t)(c4 + q) * DB + lane * DBL + e];
            iq[e] = t_iv[(size_t)(c4 + q) * DB + lane * DBL + e];
          }
#pragma unroll
          for (int r = 0; r < kBR; ++r) {
            const float w = q == 0 ? w4[r].x : (q == 1 ? w4[r].y : (q == 2 ? w4[r].z : w4[r].w));
#pragma unroll
            for (int e = 0; e < DBL; ++e) acc[r][e] = fmaf(w, (x[r][e] - mq[e]) * iq[e], acc[r][e]);
          }
        }
      }
    }
  }
  if (!have) return;
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    float dot = 0.f;
#pragma unroll
    for (int e = 0; e < DBL; ++e) {
      const float dx = -acc[r][e];
      dot = fmaf(x[r][e], dx, dot);
      if (rg * kBR + r < K) p.dX[((size_t)s * K + rowi[r]) * D + dbase + lane * DBL + e] = dx;
    }
    dot = warp_sum(dot);
    if (lane == 0 && rg * kBR + r < K) p.rdot[((size_t)s * K + rowi[r]) * p.NB + blockIdx.x] = dot;
  }
}

struct Plan {
  int CB, nblk, DBL, NB;
  int loss_where;      // resid_loss_kernel<WHERE>
  int bwd_cc, bwd_rb;  // chunked backward: columns per chunk (0 = all columns resident), row blocks
  size_t smem_fwd, smem_loss, smem_bwd;
  // scratch offsets (floats)
  size_t o_nrm, o_consts, o_lm, o_wt, o_dx, o_rdot, o_inv, o_p, total;
};

int make_plan(int S, int K, int M, int D, Plan& pl) {
  UA_UNSUPPORTED(D % 128 != 0, "residual learning: D=%d must be a multiple of 128", D);
  UA_UNSUPPORTED(M % 4 != 0 || M > 16, "residual learning: M=%d must be 4, 8, 12 or 16", M);
  const size_t budget = 220 * 1024;
  // forward: CB classes per CTA, tiles 2*CB*M*D floats + lj K*CB*M floats. One class per CTA on purpose: several small
  // CTAs per SM hide latency far better than one fat CTA with eight warps (measured at S=15, K=40: 10 Adam steps take
  // 872 us with CB=1, 940 with CB=2, 1015 with CB=5).
  const size_t per_class = ((size_t)2 * M * D + (size_t)K * M) * sizeof(float);
  int cbmax = (int)(budget / per_class);
  UA_UNSUPPORTED(cbmax < 1, "residual learning: M*D=%d does not fit in shared memory", M * D);
  int cb = 1;
  if (g_resid_cb > 0) cb = g_resid_cb;
  if (cb > cbmax) cb = cbmax;
  if (cb > K) cb = K;
  pl.CB = cb;
  pl.nblk = (K + pl.CB - 1) / pl.CB;
  pl.smem_fwd = ((size_t)2 * pl.CB * M * D + (size_t)K * pl.CB * M) * sizeof(float);
  UA_UNSUPPORTED(pl.smem_fwd > budget, "residual learning: K*M=%d rows of log-likelihoods do not fit next to the class tile", K * M);
  // loss: LM and P in shared memory while they fit, then P only, then P in global scratch (one CTA per stream either way)
  const size_t kk = (size_t)K * K * sizeof(float), k3 = (size_t)3 * K * sizeof(float);
  UA_UNSUPPORTED(k3 > budget || (long long)K * K * M >= (1LL << 31), "residual learning: K=%d too large", K);
  pl.loss_where = 2 * kk + k3 <= budget ? 0 : (kk + k3 <= budget ? 1 : 2);
  pl.smem_loss = k3 + (pl.loss_where == 0 ? 2 * kk : (pl.loss_where == 1 ? kk : 0));
  // backward: the D slice (32 or 64 wide) whose K*M columns fit; the narrow slice when that makes two CTAs resident.
  // When not even the narrow slice fits (K*M > 880), the chunked kernel: 400 columns per chunk, 40 rows per CTA.
  pl.DBL = 0, pl.bwd_cc = 0, pl.bwd_rb = 1;
  for (int dbl = 2; dbl >= 1 && !pl.DBL; --dbl)
    if ((size_t)2 * K * M * 32 * dbl * sizeof(float) <= budget && D % (32 * dbl) == 0) pl.DBL = dbl;
  if (!pl.DBL) {
    pl.DBL = 1;
    pl.bwd_cc = 400;
    pl.bwd_rb = ((K + kBR - 1) / kBR + kThreads / 32 - 1) / (kThreads / 32);
    UA_UNSUPPORTED(pl.bwd_rb > 65535, "residual learning: K=%d too large", K);
  } else {
    if (pl.DBL == 2 && (size_t)2 * K * M * 32 * sizeof(float) <= budget / 2) pl.DBL = 1;   // two CTAs resident per SM
    if (g_resid_dbl == 1 || g_resid_dbl == 2) {
      if ((size_t)2 * K * M * 32 * g_resid_dbl * sizeof(float) <= budget && D % (32 * g_resid_dbl) == 0) pl.DBL = g_resid_dbl;
    }
  }
  pl.NB = D / (32 * pl.DBL);
  pl.smem_bwd = (size_t)2 * (pl.bwd_cc ? pl.bwd_cc : K * M) * 32 * pl.DBL * sizeof(float);
  return UA_OK;
}

size_t align4(size_t v) { return (v + 3) / 4 * 4; }

int plan_aligned(int S, int K, int M, int D, Plan& pl) {
  int rc = make_plan(S, K, M, D, pl);
  if (rc != UA_OK) return rc;
  size_t o = 0;
  pl.o_nrm = o, o = align4(o + (size_t)S * K);
  pl.o_consts = o, o = align4(o + (size_t)S * K * M * 2);
  pl.o_lm = o, o = align4(o + (size_t)S * K * K);
  pl.o_wt = o, o = align4(o + (size_t)S * K * K * M);
  pl.o_dx = o, o = align4(o + (size_t)S * K * D);
  pl.o_rdot = o, o = align4(o + (size_t)S * K * pl.NB);
  pl.o_inv = o, o = align4(o + (size_t)S * K * M * D);
  pl.o_p = o, o = align4(o + (pl.loss_where == 2 ? (size_t)S * K * K : 0));
  pl.total = o;
  return UA_OK;
}

template <typename Kern>
cudaError_t opt_in(Kern kern, size_t smem) {
  return smem > 48 * 1024 ? cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : cudaSuccess;
}

struct Run {
  Plan pl;
  const float *text0, *mu, *var, *pi;
  long long text0_stride;
  int S, K, M, D;
  float eps;
  float *X, *scratch;
  cudaStream_t st;
};

int launch_prep(const Run& r) {
  const int rows = r.S * r.K * r.M;
  resid_prep_kernel<<<(rows + 7) / 8, kThreads, 0, r.st>>>(r.var, r.pi, rows, r.D, r.eps, r.scratch + r.pl.o_consts,
                                                               r.scratch + r.pl.o_inv);
  return check_launch("resid_prep");
}

int launch_embed(const Run& r, EmbedParams ep) {
  ep.text0 = r.text0, ep.text0_stride = r.text0_stride;
  ep.dX = r.scratch + r.pl.o_dx, ep.rdot = r.scratch + r.pl.o_rdot, ep.NB = r.pl.NB;
  ep.X = r.X, ep.nrm = r.scratch + r.pl.o_nrm, ep.K = r.K, ep.D = r.D;
  resid_embed_kernel<<<r.S * r.K, 128, 0, r.st>>>(ep);
  return check_launch("resid_embed");
}

int launch_fwd_loss_bwd(const Run& r, float* out_loss, int loss_stride, int loss_index, bool backward) {
  FwdParams fp;
  fp.X = r.X, fp.mu = r.mu, fp.var = r.scratch + r.pl.o_inv, fp.consts = r.scratch + r.pl.o_consts;
  fp.LM = r.scratch + r.pl.o_lm, fp.Wt = r.scratch + r.pl.o_wt;
  fp.K = r.K, fp.M = r.M, fp.D = r.D, fp.CB = r.pl.CB, fp.eps = r.eps;
  if (opt_in(resid_forward_kernel, r.pl.smem_fwd) != cudaSuccess) {
    set_error("residual learning: cannot opt in to %zu B of shared memory (forward)", r.pl.smem_fwd);
    return UA_ERR_CUDA;
  }
  resid_forward_kernel<<<dim3(r.pl.nblk, r.S), kThreads, r.pl.smem_fwd, r.st>>>(fp);
  int rc = check_launch("resid_forward");
  if (rc != UA_OK) return rc;
  {
    float* pg = r.scratch + r.pl.o_p;
    cudaError_t e;
    if (r.pl.loss_where == 0) {
      e = opt_in(resid_loss_kernel<0>, r.pl.smem_loss);
      if (e == cudaSuccess) resid_loss_kernel<0><<<r.S, kLossThreads, r.pl.smem_loss, r.st>>>(fp.LM, fp.Wt, pg, r.K, r.M, out_loss, loss_stride, loss_index);
    } else if (r.pl.loss_where == 1) {
      e = opt_in(resid_loss_kernel<1>, r.pl.smem_loss);
      if (e == cudaSuccess) resid_loss_kernel<1><<<r.S, kLossThreads, r.pl.smem_loss, r.st>>>(fp.LM, fp.Wt, pg, r.K, r.M, out_loss, loss_stride, loss_index);
    } else {
      e = opt_in(resid_loss_kernel<2>, r.pl.smem_loss);
      if (e == cudaSuccess) resid_loss_kernel<2><<<r.S, kLossThreads, r.pl.smem_loss, r.st>>>(fp.LM, fp.Wt, pg, r.K, r.M, out_loss, loss_stride, loss_index);
    }
    if (e != cudaSuccess) {
      set_error("residual learning: cannot opt in to %zu B of shared memory (loss)", r.pl.smem_loss);
      return UA_ERR_CUDA;
    }
  }
  rc = check_launch("resid_loss");
  if (rc != UA_OK || !backward) return rc;
  BwdParams bp;
  bp.X = r.X, bp.mu = r.mu, bp.var = r.scratch + r.pl.o_inv, bp.Wt = fp.Wt, bp.dX = r.scratch + r.pl.o_dx;
  bp.rdot = r.scratch + r.pl.o_rdot, bp.K = r.K, bp.M = r.M, bp.D = r.D, bp.NB = r.pl.NB, bp.eps = r.eps;
  cudaError_t e;
  if (r.pl.bwd_cc) {
    e = opt_in(resid_backward_chunked_kernel<1>, r.pl.smem_bwd);
    if (e == cudaSuccess)
      resid_backward_chunked_kernel<1><<<dim3(r.pl.NB, r.pl.bwd_rb, r.S), kThreads, r.pl.smem_bwd, r.st>>>(bp, r.pl.bwd_cc);
  } else if (r.pl.DBL == 2) {
    e = opt_in(resid_backward_kernel<2>, r.pl.smem_bwd);
    if (e == cudaSuccess) resid_backward_kernel<2><<<dim3(r.pl.NB, r.S), kThreads, r.pl.smem_bwd, r.st>>>(bp);
  } else {
    e = opt_in(resid_backward_kernel<1>, r.pl.smem_bwd);
    if (e == cudaSuccess) resid_backward_kernel<1><<<dim3(r.pl.NB, r.S), kThreads, r.pl.smem_bwd, r.st>>>(bp);
  }
  if (e != cudaSuccess) {
    set_error("residual learning: cannot opt in to %zu B of shared memory (backward)", r.pl.smem_bwd);
    return UA_ERR_CUDA;
  }
  return check_launch("resid_backward");
}

int check_common(const float* text0, const float* mu, const float* var, const float* pi, int S, int K, int M, int D,
                 float* scratch, long long scratch_floats, Plan& pl) {
  UA_REQUIRE(text0 && mu && var && pi, "residual learning: NULL input");
  UA_REQUIRE(S >= 1 && K >= 1 && M >= 1 && D >= 1, "residual learning: bad sizes S=%d K=%d M=%d D=%d", S, K, M, D);
  int rc = plan_aligned(S, K, M, D, pl);
  if (rc != UA_OK) return rc;
  UA_REQUIRE(scratch && scratch_floats >= (long long)pl.total, "residual learning: scratch of %lld floats, need %zu",
             scratch_floats, pl.total);
  UA_REQUIRE(((uintptr_t)scratch % 16 == 0) && ((uintptr_t)mu % 16 == 0) && ((uintptr_t)var % 16 == 0),
             "residual learning: mu, var and scratch must be 16-byte aligned");
  return UA_OK;
}

}  // namespace
}  // namespace ua

extern "C" long long ua_residual_scratch_floats(int S, int K, int M, int D) {
  ua::Plan pl;
  if (S < 1 || K < 1 || M < 1 || D < 1 || ua::plan_aligned(S, K, M, D, pl) != UA_OK) return -1;
  return (long long)pl.total;
}

extern "C" int ua_residual_learn_f32(const float* text0, long long text0_stream_stride, float* residual, float* adam_m,
                                     float* adam_v, int32_t* adam_t, const float* mu, const float* var,
                                     const float* pi, int S, int K, int M, int D, float eps, double lr, double beta1,
                                     double beta2, double adam_eps, int iters, float* out_text, float* out_loss,
                                     float* scratch, long long scratch_floats, void* stream) {
  using namespace ua;
  Run r;
  int rc = check_common(text0, mu, var, pi, S, K, M, D, scratch, scratch_floats, r.pl);
  if (rc != UA_OK) return rc;
  UA_REQUIRE(residual && out_text, "ua_residual_learn_f32: residual / out_text is NULL");
  UA_REQUIRE(iters >= 0 && (iters == 0 || (adam_m && adam_v && adam_t)), "ua_residual_learn_f32: Adam state is NULL");
  UA_REQUIRE((uintptr_t)out_text % 16 == 0, "ua_residual_learn_f32: out_text must be 16-byte aligned");
  r.text0 = text0, r.text0_stride = text0_stream_stride, r.mu = mu, r.var = var, r.pi = pi;
  r.S = S, r.K = K, r.M = M, r.D = D, r.eps = eps, r.X = out_text, r.scratch = scratch, r.st = (cudaStream_t)stream;
  EmbedParams ep = {};
  ep.residual = residual, ep.mode = 0;
  if (iters > 0 && (rc = launch_prep(r)) != UA_OK) return rc;
  if ((rc = launch_embed(r, ep)) != UA_OK) return rc;
  for (int it = 1; it <= iters; ++it) {
    if ((rc = launch_fwd_loss_bwd(r, out_loss, iters, it - 1, true)) != UA_OK) return rc;
    ep.mode = 1, ep.adam_m = adam_m, ep.adam_v = adam_v, ep.adam_t = adam_t, ep.t_index = it;
    ep.lr = lr, ep.beta1 = beta1, ep.beta2 = beta2, ep.adam_eps = adam_eps, ep.out_grad = nullptr;
    if ((rc = launch_embed(r, ep)) != UA_OK) return rc;
  }
  if (iters > 0) {
    resid_bump_t_kernel<<<(S + 127) / 128, 128, 0, r.st>>>(adam_t, S, iters);
    rc = check_launch("resid_bump_t");
  }
  return rc;
}

extern "C" int ua_align_loss_grad_f32(const float* text0, long long text0_stream_stride, const float* residual,
                                      const float* mu, const float* var, const float* pi, int S, int K, int M, int D,
                                      float eps, float* out_emb, float* out_loss, float* out_lm, float* out_grad,
                                      float* scratch, long long scratch_floats, void* stream) {
  using namespace ua;
  Run r;
  int rc = check_common(text0, mu, var, pi, S, K, M, D, scratch, scratch_floats, r.pl);
  if (rc != UA_OK) return rc;
  UA_REQUIRE(residual && out_emb, "ua_align_loss_grad_f32: residual / out_emb is NULL");
  UA_REQUIRE((uintptr_t)out_emb % 16 == 0, "ua_align_loss_grad_f32: out_emb must be 16-byte aligned");
  r.text0 = text0, r.text0_stride = text0_stream_stride, r.mu = mu, r.var = var, r.pi = pi;
  r.S = S, r.K = K, r.M = M, r.D = D, r.eps = eps, r.X = out_emb, r.scratch = scratch, r.st = (cudaStream_t)stream;
  EmbedParams ep = {};
  ep.residual = const_cast<float*>(residual), ep.mode = 0;
  if ((rc = launch_prep(r)) != UA_OK) return rc;
  if ((rc = launch_embed(r, ep)) != UA_OK) return rc;
  if ((rc = launch_fwd_loss_bwd(r, out_loss, 1, 0, out_grad != nullptr)) != UA_OK) return rc;
  if (out_lm) {
    cudaError_t e = cudaMemcpyAsync(out_lm, scratch + r.pl.o_lm, (size_t)S * K * K * sizeof(float),
                                    cudaMemcpyDeviceToDevice, r.st);
    if (e != cudaSuccess) {
      set_error("ua_align_loss_grad_f32: copy of the likelihood matrix failed: %s", cudaGetErrorString(e));
      return UA_ERR_CUDA;
    }
  }
  if (out_grad) {
    ep.mode = 1, ep.out_grad = out_grad;   // gradient only: no Adam state, the residual is left untouched
    rc = launch_embed(r, ep);
  }
  return rc;
}
