// MODE-DOTA cache step for BATCHED inputs (BASELINE cfg 5: B = 64 rows per step), sm_100a.
// Same arithmetic as modedota.cu (dota_mixture.py:117-156, :162-234, :236-267); different decomposition.
//
// With many rows per class the step is no longer a stream over the cache: the work is B*K*M*D (row, mode, d)
// terms against 16*K*M*D state bytes, and one CTA per class (the general kernel) leaves most SMs idle at K = 55
// and walks the rows two at a time. Here a thread-block CLUSTER owns one class and splits the feature axis:
//   * CTA `rank` of the cluster holds the D/nsplit slice of mu, var (-> 1/v) and of all Bp + B input rows in shared
//     memory (Ds = 128..256 columns: 65 rows x 128 columns = 33 KB);
//   * phase 1: a warp takes 4 rows x M modes as a register tile, every lane 4 columns per 128 (float4 shared-memory
//     loads, conflict-free), 3 instructions per (row, mode, d) term; a transposing warp reduction (31 shuffles for 32
//     sums) leaves one finished partial per lane;
//   * the slice partials are exchanged through DISTRIBUTED SHARED MEMORY (cluster.map_shared_rank), summed in rank
//     order by every CTA (deterministic, identical in all CTAs), so no scratch buffer and no second launch;
//   * phase 2 (log joint, logsumexp, responsibilities) is replicated per CTA, the M-step runs on the CTA's own slice
//     with the general kernel's operation order, and the new state goes straight from registers to HBM.
// Grid = (nsplit, S*K): 440 CTAs at K = 55, 2-3 resident per SM.
#include <cooperative_groups.h>
#include "modedota_params.cuh"

namespace cg = cooperative_groups;

namespace ua {

extern int g_modedota_batch;

namespace {

constexpr int kBT = 256;  // threads per CTA (8 warps)

// Sum NV (= 32 or 16) per-lane values over the warp so that lane l ends up with the total of value l (NV = 32) or of
// value l & 15 (NV = 16): each step exchanges half of the remaining values with the xor partner.
template <int NV>
__device__ __forceinline__ float warp_transpose_sum(float (&v)[NV], int lane) {
#pragma unroll
  for (int o = NV / 2; o >= 1; o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = hi ? v[i] : v[i + o];
      const float keep = hi ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(kFullMask, send, o);
    }
  }
  float r = v[0];
  if (NV == 16) r += __shfl_xor_sync(kFullMask, r, 16);
  return r;
}


// M-step of one feature column for MPT consecutive modes starting at m0: wx = sum_b gamma*x, wxsq = sum_b gamma*x^2 in
// the general kernel's order (first row by multiplication, the rest by FMA), then the reference's expanded-form update.
template <int MM, int MPT>
__device__ __forceinline__ void mstep_column(const float* __restrict__ xcol, int Ds, int B,
                                             const float* __restrict__ s_gamma, int m0, const float* s_cold,
                                             const float* s_sumg, const float* s_rden, const float* mu_col,
                                             const float* var_col, float* __restrict__ g_mu, float* __restrict__ g_var,
                                             int D) {
  static_assert(MPT % 2 == 0, "modes per thread must be even (vector loads of the responsibilities)");
  float wx[MPT], wxsq[MPT], gm[MPT];
  auto load_gamma = [&](int b) {
    if (MPT % 4 == 0) {
#pragma unroll
      for (int j = 0; j < MPT; j += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(s_gamma + b * MM + m0 + j);
        gm[j] = g4.x, gm[j + 1] = g4.y, gm[j + 2] = g4.z, gm[j + 3] = g4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < MPT; j += 2) {
        const float2 g2 = *reinterpret_cast<const float2*>(s_gamma + b * MM + m0 + j);
        gm[j] = g2.x, gm[j + 1] = g2.y;
      }
    }
  };
  {
    const float x0 = xcol[0], x0sq = __fmul_rn(x0, x0);
    load_gamma(0);
#pragma unroll
    for (int j = 0; j < MPT; ++j) wx[j] = __fmul_rn(gm[j], x0), wxsq[j] = __fmul_rn(gm[j], x0sq);
  }
#pragma unroll 4
  for (int b = 1; b < B; ++b) {
    const float xv = xcol[(size_t)b * Ds], xsq = __fmul_rn(xv, xv);
    load_gamma(b);
#pragma unroll
    for (int j = 0; j < MPT; ++j) wx[j] = __fmaf_rn(gm[j], xv, wx[j]), wxsq[j] = __fmaf_rn(gm[j], xsq, wxsq[j]);
  }
#pragma unroll
  for (int j = 0; j < MPT; ++j) {
    const int m = m0 + j;
    const float cold = s_cold[m], sg = s_sumg[m], rden = s_rden[m];
    const float mu_ = mu_col[m * Ds], var_ = var_col[m * Ds];
    const float mu_new = __fmul_rn(__fadd_rn(__fmul_rn(cold, mu_), wx[j]), rden);
    const float term2 = __fmul_rn(__fmul_rn(-2.0f, mu_), wx[j]);
    const float term3 = __fmul_rn(sg, __fmul_rn(mu_, mu_));
    const float wsd = __fadd_rn(__fadd_rn(wxsq[j], term2), term3);
    g_mu[(size_t)m * D] = mu_new;
    g_var[(size_t)m * D] = fmaxf(__fmul_rn(__fadd_rn(__fmul_rn(cold, var_), wsd), rden), 1e-8f);
  }
}

template <int MM>
__global__ void __launch_bounds__(kBT, 3) modedota_batch_kernel(const StepParams p, int Ds) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int nsplit = (int)gridDim.x, rank = (int)blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int item = blockIdx.y, s = item / p.K, k = item - s * p.K;
  const int D = p.D, Bp = p.Bp, B = p.B, R = Bp + B, d0 = rank * Ds;
  const int Ds4 = Ds >> 2;

  float* sx = reinterpret_cast<float*>(s_raw);       // [R][Ds]    input rows (predict rows first)
  float* smu = sx + (size_t)R * Ds;                  // [MM][Ds]
  float* svar = smu + MM * Ds;                       // [MM][Ds]   raw var (the M-step updates the raw value)
  float* sinv = svar + MM * Ds;                      // [MM][Ds]   1 / clamp(var + eps)
  float* s_gamma = sinv + MM * Ds;                   // [B][MM]    (16-byte aligned: read as float4 by the M-step)
  float* s_part = s_gamma + (size_t)max(B, 1) * MM;  // [R*MM + MM] this slice's Mahalanobis sums, then log-dets
  float* s_tot = s_part + (R * MM + MM);             // [R*MM + MM] summed over the cluster
  float* s_gc = s_tot + (R * MM + MM);               // [B]
  float* s_small = s_gc + max(B, 1);                 // logpi[MM], cold[MM], sumg[MM], rden[MM], cnew[MM]
  float* s_logpi = s_small;
  float* s_cold = s_small + MM;
  float* s_sumg = s_small + 2 * MM;
  float* s_rden = s_small + 3 * MM;
  float* s_cnew = s_small + 4 * MM;

  // ---- stage the slice ---------------------------------------------------------------------------------------
  for (int row = warp; row < R; row += kBT / 32) {      // a warp per row: no index division, 512-byte coalesced loads
    const float* src = row < Bp ? p.x_pred + ((size_t)s * Bp + row) * D : p.x_fit + ((size_t)s * B + (row - Bp)) * D;
    const float4* src4 = reinterpret_cast<const float4*>(src + d0);
    float4* dst4 = reinterpret_cast<float4*>(sx + (size_t)row * Ds);
    for (int q = lane; q < Ds4; q += 32) dst4[q] = __ldg(src4 + q);
  }
  for (int m = warp; m < MM; m += kBT / 32) {
    const size_t g = ((size_t)item * MM + m) * D + d0;
    for (int q = lane; q < Ds4; q += 32) {
      const float4 m4 = __ldg(reinterpret_cast<const float4*>(p.mu + g) + q);
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(p.var + g) + q);
      float4 i4;
      i4.x = rcp_rn_normal(fmaxf(__fadd_rn(v4.x, p.eps), 1e-8f));
      i4.y = rcp_rn_normal(fmaxf(__fadd_rn(v4.y, p.eps), 1e-8f));
      i4.z = rcp_rn_normal(fmaxf(__fadd_rn(v4.z, p.eps), 1e-8f));
      i4.w = rcp_rn_normal(fmaxf(__fadd_rn(v4.w, p.eps), 1e-8f));
      reinterpret_cast<float4*>(smu + m * Ds)[q] = m4;
      reinterpret_cast<float4*>(svar + m * Ds)[q] = v4;
      reinterpret_cast<float4*>(sinv + m * Ds)[q] = i4;
    }
  }
  if (tid < MM) {
    s_logpi[tid] = logf(__fadd_rn(p.pi[(size_t)item * MM + tid], 1e-10f));
    s_cold[tid] = p.c[(size_t)item * MM + tid];
  }
  if (tid < B) s_gc[tid] = __ldg(p.gamma + ((size_t)s * B + tid) * p.ldg + p.kg_off + k);
  __syncthreads();

  // ---- phase 1: Mahalanobis partial sums of this slice, 4 rows x MM modes per warp pass ------------------------
  const int RG = (R + 3) >> 2;
  for (int rg = warp; rg < RG; rg += kBT / 32) {
    float acc[4 * MM];
#pragma unroll
    for (int i = 0; i < 4 * MM; ++i) acc[i] = 0.f;
    int rows[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) rows[q] = min(rg * 4 + q, R - 1);
    const int nvalid = min(4, R - rg * 4);
    for (int col = lane * 4; col < Ds; col += 128) {
      float4 xr[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) xr[q] = *reinterpret_cast<const float4*>(sx + (size_t)rows[q] * Ds + col);
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        const float4 m4 = *reinterpret_cast<const float4*>(smu + m * Ds + col);
        const float4 i4 = *reinterpret_cast<const float4*>(sinv + m * Ds + col);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q > 0 && q >= nvalid) break;      // ragged last row group (R = 65: one predict row + 64 fit rows)
          const float dx = __fsub_rn(xr[q].x, m4.x), dy = __fsub_rn(xr[q].y, m4.y);
          const float dz = __fsub_rn(xr[q].z, m4.z), dw = __fsub_rn(xr[q].w, m4.w);
          float a = acc[q * MM + m];
          a = fmaf(dx * dx, i4.x, a);
          a = fmaf(dy * dy, i4.y, a);
          a = fmaf(dz * dz, i4.z, a);
          a = fmaf(dw * dw, i4.w, a);
          acc[q * MM + m] = a;
        }
      }
    }
    const float tot = warp_transpose_sum<4 * MM>(acc, lane);
    if (lane < 4 * MM) {
      const int row = rg * 4 + lane / MM;
      if (row < R) s_part[row * MM + (lane % MM)] = tot;
    }
  }
  // log-determinant partials: one logf per element like the reference, warp m sums mode m's slice
  if (warp < MM) {
    float ld = 0.f;
    for (int d = lane; d < Ds; d += 32) ld += logf(fmaxf(__fadd_rn(svar[warp * Ds + d], p.eps), 1e-8f));
    ld = warp_sum(ld);
    if (lane == 0) s_part[R * MM + warp] = ld;
  }

  // ---- exchange: every CTA sums the slices in rank order (distributed shared memory) ---------------------------
  cluster.sync();
  for (int e = tid; e < R * MM + MM; e += kBT) {
    float t = 0.f;
    for (int r = 0; r < nsplit; ++r) t += cluster.map_shared_rank(s_part, r)[e];
    s_tot[e] = t;
  }
  __syncthreads();

  // ---- phase 2: log joint, predict logits, responsibilities (replicated in every CTA of the cluster) -----------
  for (int r = tid; r < R; r += kBT) {
    float lj[MM];
    float mx = -INFINITY;
#pragma unroll
    for (int m = 0; m < MM; ++m) {
      const float ll = __fmul_rn(-0.5f, __fadd_rn(s_tot[R * MM + m], s_tot[r * MM + m]));
      lj[m] = __fadd_rn(s_logpi[m], ll);
      mx = fmaxf(mx, lj[m]);
    }
    float se = 0.f;
#pragma unroll
    for (int m = 0; m < MM; ++m) se += expf(lj[m] - mx);
    const float lse = __fadd_rn(logf(se), mx);
    if (r < Bp) {
      if (rank == 0) p.out_logits[((size_t)s * Bp + r) * p.ldo + p.ko_off + k] = lse;
    } else {
      // exp(log_joint - logsumexp) exactly as dota_mixture.py:182-183 (not the softmax quotient)
      const int b = r - Bp;
      const float gc = s_gc[b];
#pragma unroll
      for (int m = 0; m < MM; ++m) s_gamma[b * MM + m] = __fmul_rn(gc, expf(__fsub_rn(lj[m], lse)));
    }
  }
  __syncthreads();

  if (B > 0) {
    // ---- phase 3: soft counts (warp m sums mode m over the rows) ----------------------------------------------
    if (warp < MM) {
      float sg = 0.f;
      for (int b = lane; b < B; b += 32) sg += s_gamma[b * MM + warp];
      sg = warp_sum(sg);
      if (lane == 0) {
        const float cnew = __fadd_rn(s_cold[warp], sg);
        s_sumg[warp] = sg;
        s_cnew[warp] = cnew;
        s_rden[warp] = rcp_rn_normal(__fadd_rn(cnew, 1e-10f));
      }
    }
    __syncthreads();
    if (rank == 0) {
      if (tid < MM) {
        float ck = 0.f;
#pragma unroll
        for (int m = 0; m < MM; ++m) ck += s_cnew[m];
        p.c[(size_t)item * MM + tid] = s_cnew[tid];
        p.pi[(size_t)item * MM + tid] = __fdiv_rn(s_cnew[tid], __fadd_rn(ck, 1e-10f));
      }
      if (warp == 1) {
        float gs = 0.f;
        for (int b = lane; b < B; b += 32) gs += s_gc[b];
        gs = warp_sum(gs);
        if (lane == 0) p.class_counts[item] = p.class_counts[item] + gs;
      }
    }

    // ---- phase 4: M-step on this slice: thread = one column d, MM / MG modes ----------------------------------
    const int MG = (kBT / Ds >= 2 && MM % 2 == 0) ? 2 : 1;     // Ds = 128: two mode groups, else one
    if (MG == 2) {
      if (tid < 2 * Ds) {
        const int d = tid % Ds, g = tid / Ds;
        const size_t gbase = (size_t)item * MM * D + d0 + d;
        mstep_column<MM, MM / 2>(sx + (size_t)Bp * Ds + d, Ds, B, s_gamma, g * (MM / 2), s_cold, s_sumg, s_rden, smu + d,
                                 svar + d, p.mu + gbase, p.var + gbase, D);
      }
    } else {
      // slices wider than the CTA (Ds = 384, 512: D = 1152, 3072, 4096) take several columns per thread
      for (int d = tid; d < Ds; d += kBT) {
        const size_t gbase = (size_t)item * MM * D + d0 + d;
        mstep_column<MM, MM>(sx + (size_t)Bp * Ds + d, Ds, B, s_gamma, 0, s_cold, s_sumg, s_rden, smu + d, svar + d,
                             p.mu + gbase, p.var + gbase, D);
      }
    }
  }
  cluster.sync();   // nobody retires while a peer may still read its partials
}

size_t batch_smem_bytes(int R, int B, int M, int Ds) {
  return ((size_t)R * Ds + (size_t)3 * M * Ds + (size_t)2 * (R * M + M) + (size_t)(B > 0 ? B : 1) * (M + 1) + 5 * M) *
             sizeof(float) + 16;
}

template <int MM>
cudaError_t launch_batch(const StepParams& p, int nsplit, int Ds, size_t smem, cudaStream_t st) {
  auto kern = modedota_batch_kernel<MM>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)nsplit, (unsigned)(p.S * p.K), 1);
  cfg.blockDim = dim3(kBT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)nsplit;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p, Ds);
}

}  // namespace

int modedota_batch_launch(const StepParams& p, cudaStream_t st) {
  const int R = p.Bp + p.B;
  if (g_modedota_batch < 0 || R < 8 || !(p.M == 4 || p.M == 8) || !p.vec_ok) return 1;
  if ((long long)p.S * p.K > 65535) return 1;
  if (((uintptr_t)p.mu & 15) || ((uintptr_t)p.var & 15)) return 1;
  // D-splits = cluster size: the most CTAs (<= 8, the portable cluster limit) whose slice is a multiple of 128 columns
  int nsplit = 0;
  for (int cand = 8; cand >= 1 && !nsplit; --cand) {
    if (g_modedota_batch > 0 && cand != g_modedota_batch) continue;
    if (p.D % cand) continue;
    const int Ds = p.D / cand;
    if (Ds % 128 || Ds > 256 * 2) continue;
    if (batch_smem_bytes(R, p.B, p.M, Ds) > 200 * 1024) continue;
    nsplit = cand;
  }
  if (!nsplit) return 1;
  const int Ds = p.D / nsplit;
  const size_t smem = batch_smem_bytes(R, p.B, p.M, Ds);
  cudaError_t e = p.M == 4 ? launch_batch<4>(p, nsplit, Ds, smem, st) : launch_batch<8>(p, nsplit, Ds, smem, st);
  if (e != cudaSuccess) {
    set_error("ua_modedota_step_f32(batch): launch failed: %s", cudaGetErrorString(e));
    return UA_ERR_CUDA;
  }
  return check_launch("ua_modedota_step_f32(batch)");
}

}  // namespace ua
