// DOTA (full-covariance) cache kernels for sm_100a. Replaces dota.py:41-63 (fit) and :72-87 (predict);
// dota.py:66-69 (update) is csrc/spdinv.cu; ua_dota_regularize_f32 feeds the library inverse kept for widths that kernel does not take.
//
// fit is a pure HBM stream over Sigma [K,D,D] (read + write once, 8*K*D^2 bytes): each thread owns one float4 of a
// (row, 4 columns) position and walks the K classes, applying the rank-B update and accumulating the class mean
// (overall_Sigma) on the fly, so neither delta (K,D,D) nor a second pass for the mean ever touch memory.
#include <cooperative_groups.h>
#include "common.cuh"
#include "modedota_params.cuh"

namespace cg = cooperative_groups;

namespace ua {

int g_dota_ksplit = 0;   // tuning: class splits (cluster size) of the fit kernel, 0 = heuristic
int g_dota_staged = 1;   // tuning: 0 disables the staged batch-1 kernel (dota_sigma_b1_kernel)
int g_dota_ka = 0;       // tuning: Sigma loads in flight per thread of the staged kernel (8, 10, 20; 0 = heuristic)
int g_dota_pdl = 1;      // tuning: 0 launches the mean kernel without programmatic dependent launch

namespace {

constexpr int kTileRows = 8;     // rows of Sigma per CTA
constexpr int kTileCols = 128;   // columns per CTA (32 threads x float4)

// grid: (ceil(D/128), ceil(D/8), KS); block: (32, 8); cluster (1, 1, KS).
// The K classes are split over the KS CTAs of a thread-block cluster (at D = 512 a grid of tiles alone is 256 CTAs:
// a fifth of the machine's thread slots, too few loads in flight for HBM); each CTA streams its classes and keeps the
// partial class sum of its tile in registers, the partials meet in rank 0 through distributed shared memory in rank
// order (deterministic), which writes overall_Sigma. KS = 1 runs without a cluster.
__global__ void __launch_bounds__(256)
    dota_sigma_kernel(const float* __restrict__ x, const float* __restrict__ y, int B, const float* __restrict__ mu,
                      const float* __restrict__ c, float* __restrict__ Sigma, float* __restrict__ overall, int K,
                      int D) {
  __shared__ __align__(16) float4 s_partial[256];
  const int KS = (int)gridDim.z, ks = (int)blockIdx.z;
  const int j = blockIdx.x * kTileCols + threadIdx.x * 4;
  const int i = blockIdx.y * kTileRows + threadIdx.y;
  const bool inside = i < D && j < D;
  float4 mean = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t DD = (size_t)D * D;
  float* sp = Sigma + (size_t)(inside ? i : 0) * D + (inside ? j : 0);
  const int kper = (K + KS - 1) / KS, k_begin = ks * kper, k_end = min(K, k_begin + kper);
  // one class: rank-B update of this thread's float4 of Sigma_k, class mean accumulated on the fly
  auto update_class = [&](int k, const float4 sg) {
    const float mi = __ldg(mu + (size_t)k * D + i);
    const float4 mj = __ldg(reinterpret_cast<const float4*>(mu + (size_t)k * D + j));
    const float ck = __ldg(c + k);
    float4 delta;
    float sw;
    {
      const float y0 = __ldg(y + k);
      const float xi = __fsub_rn(__ldg(x + i), mi);
      const float4 xj = __ldg(reinterpret_cast<const float4*>(x + j));
      const float wi = __fmul_rn(y0, xi);
      delta.x = __fmul_rn(wi, __fsub_rn(xj.x, mj.x));
      delta.y = __fmul_rn(wi, __fsub_rn(xj.y, mj.y));
      delta.z = __fmul_rn(wi, __fsub_rn(xj.z, mj.z));
      delta.w = __fmul_rn(wi, __fsub_rn(xj.w, mj.w));
      sw = y0;
    }
    for (int b = 1; b < B; ++b) {
      const float yb = __ldg(y + (size_t)b * K + k);
      const float xi = __fsub_rn(__ldg(x + (size_t)b * D + i), mi);
      const float4 xj = __ldg(reinterpret_cast<const float4*>(x + (size_t)b * D + j));
      const float wi = __fmul_rn(yb, xi);
      delta.x = __fmaf_rn(wi, __fsub_rn(xj.x, mj.x), delta.x);
      delta.y = __fmaf_rn(wi, __fsub_rn(xj.y, mj.y), delta.y);
      delta.z = __fmaf_rn(wi, __fsub_rn(xj.z, mj.z), delta.z);
      delta.w = __fmaf_rn(wi, __fsub_rn(xj.w, mj.w), delta.w);
      sw = __fadd_rn(sw, yb);
    }
    const float denom = __fadd_rn(ck, sw);
    float4 out;
    out.x = __fdiv_rn(__fadd_rn(__fmul_rn(ck, sg.x), delta.x), denom);
    out.y = __fdiv_rn(__fadd_rn(__fmul_rn(ck, sg.y), delta.y), denom);
    out.z = __fdiv_rn(__fadd_rn(__fmul_rn(ck, sg.z), delta.z), denom);
    out.w = __fdiv_rn(__fadd_rn(__fmul_rn(ck, sg.w), delta.w), denom);
    *reinterpret_cast<float4*>(sp + (size_t)k * DD) = out;
    mean.x += out.x, mean.y += out.y, mean.z += out.z, mean.w += out.w;
  };
  // The Sigma stream is what HBM sees: four independent 16-byte loads per thread are issued before any of them is
  // consumed (the straight #pragma unroll left one load in flight per thread: 2.2 TB/s at D = 512).
  constexpr int kAhead = 4;
  int k = inside ? k_begin : k_end;
  for (; k + kAhead <= k_end; k += kAhead) {
    float4 sg[kAhead];
#pragma unroll
    for (int u = 0; u < kAhead; ++u) sg[u] = __ldcs(reinterpret_cast<const float4*>(sp + (size_t)(k + u) * DD));
#pragma unroll
    for (int u = 0; u < kAhead; ++u) update_class(k + u, sg[u]);
  }
  for (; k < k_end; ++k) update_class(k, __ldcs(reinterpret_cast<const float4*>(sp + (size_t)k * DD)));
  if (KS > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    const int t = threadIdx.y * 32 + threadIdx.x;
    s_partial[t] = mean;
    cluster.sync();
    if (ks == 0) {
      for (int r = 1; r < KS; ++r) {
        const float4 p = cluster.map_shared_rank(s_partial, r)[t];
        mean.x += p.x, mean.y += p.y, mean.z += p.z, mean.w += p.w;
      }
    }
    cluster.sync();      // peers keep their partials alive until rank 0 has read them
    if (ks != 0) return;
  }
  if (!inside) return;
  const float kf = (float)K;
  float4 m4 = make_float4(__fdiv_rn(mean.x, kf), __fdiv_rn(mean.y, kf), __fdiv_rn(mean.z, kf), __fdiv_rn(mean.w, kf));
  *reinterpret_cast<float4*>(overall + (size_t)i * D + j) = m4;
}

// Batch-1 fit (the per-sample step of the reference loop, Uni_Adapter.py:411) with everything that is NOT the Sigma stream
// staged in shared memory first: per class k the tile needs mu_k on its 8 rows and 128 columns, c_k and y_k. The general
// kernel above fetches them class by class from global memory inside the streaming loop (four dependent small loads and
// four IEEE divisions per class and thread: 30 us at K = 40, D = 512 for 85 MB, issue-active 30 %). Here the CTA first
// forms, for all K classes,
//     wi[k][r]  = y_k * (x_i - mu_k[i])          (8 rows)
//     dj[k][c]  = x_j - mu_k[j]                  (128 columns)
//     sc[k]     = { c_k, 1 / (c_k + y_k) }       (one correctly-rounded reciprocal instead of four divisions)
// (K * 138 floats: 22 KB at K = 40), and the loop over the classes is then a pure stream: KA 16-byte loads of Sigma in
// flight per thread, two shared-memory reads, 16 multiply / add operations with the reference's rounding points
// (delta = wi * dj, out = (c_k * Sigma + delta) * rcp), one 16-byte store, the class mean accumulated on the fly.
template <int KA>
__global__ void __launch_bounds__(256)
    dota_sigma_b1_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ mu,
                         const float* __restrict__ c, float* __restrict__ Sigma, float* __restrict__ overall, int K, int D) {
  extern __shared__ __align__(16) float s_dyn[];
  float* s_dj = s_dyn;                                  // [K][128]
  float* s_wi = s_dj + (size_t)K * kTileCols;           // [K][8]
  float2* s_sc = reinterpret_cast<float2*>(s_wi + (size_t)K * kTileRows);   // [K]
  const int t = threadIdx.y * 32 + threadIdx.x;
  const int j0 = blockIdx.x * kTileCols, i0 = blockIdx.y * kTileRows;
  // the mean kernel may be scheduled behind this grid right away (programmatic dependent launch): it blocks in
  // griddepcontrol.wait until this grid has completed, so only its launch latency overlaps
  asm volatile("griddepcontrol.launch_dependents;");
  const int j = j0 + threadIdx.x * 4, i = i0 + threadIdx.y;
  const bool inside = i < D && j < D;
  const size_t DD = (size_t)D * D;
  float* sp = Sigma + (size_t)(inside ? i : 0) * D + (inside ? j : 0);
  // the first KA tiles of the Sigma stream do not depend on the staging below: they fly while it runs
  float4 sg0[KA];
  const bool pre = inside && K >= KA;
  if (pre) {
#pragma unroll
    for (int u = 0; u < KA; ++u) sg0[u] = __ldcs(reinterpret_cast<const float4*>(sp + (size_t)u * DD));
  }
  for (int idx = t; idx < K * kTileCols; idx += 256) {
    const int k = idx / kTileCols, col = idx - k * kTileCols, jj = j0 + col;
    s_dj[idx] = jj < D ? __fsub_rn(__ldg(x + jj), __ldg(mu + (size_t)k * D + jj)) : 0.f;
  }
  for (int idx = t; idx < K * kTileRows; idx += 256) {
    const int k = idx / kTileRows, r = idx - k * kTileRows, ii = i0 + r;
    s_wi[idx] = ii < D ? __fmul_rn(__ldg(y + k), __fsub_rn(__ldg(x + ii), __ldg(mu + (size_t)k * D + ii))) : 0.f;
  }
  for (int k = t; k < K; k += 256) {
    const float ck = __ldg(c + k);
    s_sc[k] = make_float2(ck, rcp_rn_normal(__fadd_rn(ck, __ldg(y + k))));
  }
  __syncthreads();
  if (!inside) return;
  float4 mean = make_float4(0.f, 0.f, 0.f, 0.f);
  auto update_class = [&](int k, const float4 sg) {
    const float4 dj = *reinterpret_cast<const float4*>(s_dj + (size_t)k * kTileCols + threadIdx.x * 4);
    const float wi = s_wi[k * kTileRows + threadIdx.y];
    const float2 sc = s_sc[k];
    float4 out;
    out.x = __fmul_rn(__fadd_rn(__fmul_rn(sc.x, sg.x), __fmul_rn(wi, dj.x)), sc.y);
    out.y = __fmul_rn(__fadd_rn(__fmul_rn(sc.x, sg.y), __fmul_rn(wi, dj.y)), sc.y);
    out.z = __fmul_rn(__fadd_rn(__fmul_rn(sc.x, sg.z), __fmul_rn(wi, dj.z)), sc.y);
    out.w = __fmul_rn(__fadd_rn(__fmul_rn(sc.x, sg.w), __fmul_rn(wi, dj.w)), sc.y);
    __stcs(reinterpret_cast<float4*>(sp + (size_t)k * DD), out);
    mean.x += out.x, mean.y += out.y, mean.z += out.z, mean.w += out.w;
  };
  // full batches of KA classes, every load of a batch in flight before the first is consumed; what is left runs as
  // guarded batches of 8 (K = 15 ran as 8 + 7 single dependent loads before)
  int k = 0;
  if (pre) {
#pragma unroll
    for (int u = 0; u < KA; ++u) update_class(u, sg0[u]);
    k = KA;
  }
  for (; k + KA <= K; k += KA) {
    float4 sg[KA];
#pragma unroll
    for (int u = 0; u < KA; ++u) sg[u] = __ldcs(reinterpret_cast<const float4*>(sp + (size_t)(k + u) * DD));
#pragma unroll
    for (int u = 0; u < KA; ++u) update_class(k + u, sg[u]);
  }
  if constexpr (KA >= 16) {   // (the wide variant has no registers to spare for a batched tail)
    for (; k < K; ++k) update_class(k, __ldcs(reinterpret_cast<const float4*>(sp + (size_t)k * DD)));
  } else {
    for (; k < K; k += 8) {
      float4 sg[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (k + u < K) sg[u] = __ldcs(reinterpret_cast<const float4*>(sp + (size_t)(k + u) * DD));
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (k + u < K) update_class(k + u, sg[u]);
    }
  }
  const float kf = (float)K;
  *reinterpret_cast<float4*>(overall + (size_t)i * D + j) =
      make_float4(__fdiv_rn(mean.x, kf), __fdiv_rn(mean.y, kf), __fdiv_rn(mean.z, kf), __fdiv_rn(mean.w, kf));
}

// mu' = (y^T x + c*mu) / (s + c), c' = c + s   (run after the Sigma kernel, which needs the old mu and c)
__global__ void __launch_bounds__(256)
    dota_mean_kernel(const float* __restrict__ x, const float* __restrict__ y, int B, float* __restrict__ mu,
                     float* __restrict__ c, int K, int D) {
  const int k = blockIdx.x;
  float sw = __ldg(y + k);
  for (int b = 1; b < B; ++b) sw = __fadd_rn(sw, __ldg(y + (size_t)b * K + k));
  // launched as a programmatic dependent of the Sigma kernel (which reads the OLD mu and c): everything above overlaps
  // with it, mu and c are touched only once that grid has completed (a no-op for a plain launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const float ck = c[k];
  const float denom = __fadd_rn(sw, ck);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float wx = __fmul_rn(__ldg(y + k), __ldg(x + d));
    for (int b = 1; b < B; ++b) wx = __fmaf_rn(__ldg(y + (size_t)b * K + k), __ldg(x + (size_t)b * D + d), wx);
    const float m = mu[(size_t)k * D + d];
    mu[(size_t)k * D + d] = __fdiv_rn(__fadd_rn(wx, __fmul_rn(ck, m)), denom);
  }
  __syncthreads();
  if (threadIdx.x == 0) c[k] = __fadd_rn(ck, sw);
}

// One CTA per class k. W[:,k] = Lambda @ M[:,k] (fp32 accumulate, rounded to fp16), then the discriminant in
// fp16 arithmetic with the reference's rounding points.
__global__ void __launch_bounds__(512)
    dota_predict_kernel(const __half* __restrict__ xh, int R, const __half* __restrict__ lam,
                        const float* __restrict__ mu, int K, int D, __half* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  float* s_m = reinterpret_cast<float*>(s_raw);  // [D] fp16-rounded class mean (as float)
  float* s_w = s_m + D;                          // [D] fp16-rounded W column (as float)
  __shared__ float s_part[16];
  __shared__ float s_c;
  const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
  for (int d = tid; d < D; d += blockDim.x) s_m[d] = __half2float(__float2half_rn(__ldg(mu + (size_t)k * D + d)));
  __syncthreads();
  if ((D & 7) == 0) {
    // four rows of Lambda per warp and pass: their 16-byte loads are in flight together and share the eight staged mean
    // values (one row at a time left one dependent L2 round trip per row on the critical path: 32 of them per warp at
    // D = 512); the arithmetic of a row is unchanged
    for (int i0 = warp * 4; i0 < D; i0 += W * 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int jj = lane * 8; jj < D; jj += 256) {
        uint4 raw[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) raw[r] = __ldg(reinterpret_cast<const uint4*>(lam + (size_t)(i0 + r) * D + jj));
        float m8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) m8[q] = s_m[jj + q];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const __half2* h2 = reinterpret_cast<const __half2*>(&raw[r]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __half22float2(h2[q]);
            acc[r] = fmaf(f.x, m8[2 * q], acc[r]);
            acc[r] = fmaf(f.y, m8[2 * q + 1], acc[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float a = warp_sum(acc[r]);
        if (lane == 0) s_w[i0 + r] = __half2float(__float2half_rn(a));
      }
    }
  } else {
    for (int i = warp; i < D; i += W) {
      const __half* lrow = lam + (size_t)i * D;
      float acc = 0.f;
      for (int jj = lane; jj < D; jj += 32) acc = fmaf(__half2float(lrow[jj]), s_m[jj], acc);
      acc = warp_sum(acc);
      if (lane == 0) s_w[i] = __half2float(__float2half_rn(acc));
    }
  }
  __syncthreads();
  // c = 0.5 * sum_i half(M_i * W_i)   (sum accumulated in fp32, rounded to half, then halved in half)
  float part = 0.f;
  for (int i = tid; i < D; i += blockDim.x) part += __half2float(__float2half_rn(s_m[i] * s_w[i]));
  part = warp_sum(part);
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < W; ++w) t += s_part[w];
    const float sum_h = __half2float(__float2half_rn(t));
    s_c = __half2float(__float2half_rn(0.5f * sum_h));
  }
  __syncthreads();
  for (int r = 0; r < R; ++r) {
    float dot = 0.f;
    for (int i = tid; i < D; i += blockDim.x) dot = fmaf(__half2float(xh[(size_t)r * D + i]), s_w[i], dot);
    dot = warp_sum(dot);
    __syncthreads();
    if (lane == 0) s_part[warp] = dot;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < W; ++w) t += s_part[w];
      const float s_h = __half2float(__float2half_rn(t));
      out[(size_t)r * K + k] = __float2half_rn(s_h - s_c);
    }
  }
}

__global__ void dota_regularize_kernel(const float* __restrict__ overall, int D, float one_minus_eps, float eps,
                                       float* __restrict__ out) {
  const size_t n = (size_t)D * D;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / D), j = (int)(t - (size_t)i * D);
    out[t] = __fadd_rn(__fmul_rn(one_minus_eps, overall[t]), i == j ? eps : 0.f);
  }
}

}  // namespace
}  // namespace ua

extern "C" int ua_dota_fit_f32(const float* x, const float* y, int B, float* mu, float* c, float* Sigma,
                               float* overall, int K, int D, void* stream) {
  using namespace ua;
  UA_REQUIRE(x && y && mu && c && Sigma && overall, "ua_dota_fit_f32: NULL pointer");
  UA_REQUIRE(B >= 1 && K >= 1 && D >= 1, "ua_dota_fit_f32: bad sizes B=%d K=%d D=%d", B, K, D);
  UA_UNSUPPORTED((D & 3) != 0, "ua_dota_fit_f32: D=%d must be a multiple of 4", D);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = ((D + kTileCols - 1) / kTileCols) * ((D + kTileRows - 1) / kTileRows);
  // class splits (cluster size): off by default. The sweep in tools/probe_dota_fit.py shows no reliable gain on B200
  // (D = 512: 38-40 us unsplit, 30-42 us with two splits depending on the box; D >= 1024: always slower): the
  // kernel is limited by its per-thread chain of dependent 16-byte loads, not by the number of CTAs.
  (void)tiles;
  int KS = g_dota_ksplit > 0 ? g_dota_ksplit : 1;
  if (KS > K) KS = 1;
  dim3 grid((D + kTileCols - 1) / kTileCols, (D + kTileRows - 1) / kTileRows, KS), block(32, kTileRows);
  const size_t staged_smem = (size_t)K * (kTileCols + kTileRows + 2) * sizeof(float);
  bool pdl = false;
  if (B == 1 && g_dota_staged && KS == 1 && staged_smem <= 96 * 1024) {
    // loads in flight per thread: 20 when there are that many classes (K = 40, D = 512: 27.7 -> 23.6 us against 8; two
    // CTAs per SM still fit), else 8 (+ one guarded batch for the rest)
    int ka = g_dota_ka > 0 ? g_dota_ka : (K >= 20 ? 20 : 8);
    auto kern = ka >= 20 ? dota_sigma_b1_kernel<20> : dota_sigma_b1_kernel<8>;
    if (staged_smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem);
    kern<<<dim3(grid.x, grid.y), block, staged_smem, st>>>(x, y, mu, c, Sigma, overall, K, D);
    pdl = g_dota_pdl != 0;
  } else if (KS == 1) {
    dota_sigma_kernel<<<grid, block, 0, st>>>(x, y, B, mu, c, Sigma, overall, K, D);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = 0, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = (unsigned)KS;
    cfg.attrs = attr, cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, dota_sigma_kernel, x, y, B, (const float*)mu, (const float*)c, Sigma, overall, K, D);
    if (e != cudaSuccess) {
      set_error("ua_dota_fit_f32: cluster launch failed: %s", cudaGetErrorString(e));
      return UA_ERR_CUDA;
    }
  }
  int rc = check_launch("ua_dota_fit_f32(sigma)");
  if (rc != UA_OK) return rc;
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(K), cfg.blockDim = dim3(256), cfg.dynamicSmemBytes = 0, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, dota_mean_kernel, x, y, B, mu, c, K, D) == cudaSuccess) return check_launch("ua_dota_fit_f32(mean)");
    (void)cudaGetLastError();   // fall through to the plain launch
  }
  dota_mean_kernel<<<K, 256, 0, st>>>(x, y, B, mu, c, K, D);
  return check_launch("ua_dota_fit_f32(mean)");
}

extern "C" int ua_dota_predict_f16(const void* x_h, int R, const void* Lambda_h, const float* mu, int K, int D,
                                   void* out_scores_h, void* stream) {
  using namespace ua;
  UA_REQUIRE(x_h && Lambda_h && mu && out_scores_h, "ua_dota_predict_f16: NULL pointer");
  UA_REQUIRE(R >= 1 && K >= 1 && D >= 1, "ua_dota_predict_f16: bad sizes R=%d K=%d D=%d", R, K, D);
  const size_t smem = (size_t)2 * D * sizeof(float);
  UA_UNSUPPORTED(smem > 200 * 1024, "ua_dota_predict_f16: D=%d too large", D);
  if (smem > 48 * 1024) cudaFuncSetAttribute(dota_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dota_predict_kernel<<<K, 512, smem, (cudaStream_t)stream>>>((const __half*)x_h, R, (const __half*)Lambda_h, mu, K, D,
                                                            (__half*)out_scores_h);
  return check_launch("ua_dota_predict_f16");
}

extern "C" int ua_dota_regularize_f32(const float* overall, int D, float eps, float* out, void* stream) {
  using namespace ua;
  UA_REQUIRE(overall && out && D >= 1, "ua_dota_regularize_f32: bad arguments");
  const float one_minus = (float)(1.0 - (double)eps);
  const size_t n = (size_t)D * D;
  int blocks = (int)((n + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  dota_regularize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(overall, D, one_minus, eps, out);
  return check_launch("ua_dota_regularize_f32");
}
