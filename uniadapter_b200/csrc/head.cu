// Zero-shot cosine-logit head (SIMT path): L2-normalise, scale, dot with the text rows, softmax / entropy / argmax.
// Replaces Uni_Adapter.py:21-26 (softmax_entropy) and :53-75 (get_logits_wrapper after the encoder call).
// At batch 1 the op is a GEMV that streams the (K,D) text matrix once: one warp per class, float4 loads,
// grid sized to cover the SMs. The batched variant (B >= 64) normalises and splits the rows here (ua_head_prepare_f32),
// contracts them with the text rows on the tcgen05 GEMM (gemm_tf32x3.cu) and finishes with ua_row_stats_f32.
#include "common.cuh"

namespace ua {
namespace {

constexpr int kRowsPerPass = 8;  // batch rows staged in shared memory per pass over a text row

// out_xnorm[b,:] = x[b,:] / ||x[b,:]||
__global__ void __launch_bounds__(256) l2norm_kernel(const float* __restrict__ x, int D, float* __restrict__ out) {
  __shared__ float s_part[8];
  const int b = blockIdx.x;
  const float* row = x + (size_t)b * D;
  float ss = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float v = __ldg(row + d);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_part[w];
  const float nrm = sqrtf(tot);
  for (int d = threadIdx.x; d < D; d += blockDim.x) out[(size_t)b * D + d] = __fdiv_rn(__ldg(row + d), nrm);
}

// logits[b,k] = sum_d (scale * xnorm[b,d]) * text[g(b)][k,d]; one warp per class k, blockIdx.y = text group
// (rows [g*rows_per_group, (g+1)*rows_per_group) use text matrix g; a single shared matrix is group 0 of 1).
__global__ void __launch_bounds__(256)
    logits_kernel(const float* __restrict__ xnorm, int rows_per_group, int D, const float* __restrict__ text_all, int K,
                  float scale, float* __restrict__ logits_all) {
  extern __shared__ __align__(16) float s_x[];  // [kRowsPerPass][D] scaled rows
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int k = blockIdx.x * W + warp;
  const int grp = blockIdx.y;
  const int B = rows_per_group;
  const float* text = text_all + (size_t)grp * K * D;
  xnorm += (size_t)grp * B * D;
  float* logits = logits_all + (size_t)grp * B * K;
  const float* trow = text + (size_t)(k < K ? k : 0) * D;
  const bool vec = (D & 3) == 0;
  // blockIdx.z walks the row blocks (a batch of rows meets few classes: 64 rows x 55 classes is 7 x 8 CTAs instead of 7)
  for (int b0 = blockIdx.z * kRowsPerPass; b0 < B; b0 += gridDim.z * kRowsPerPass) {
    const int nb = min(kRowsPerPass, B - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * D; i += blockDim.x)
      s_x[i] = __fmul_rn(scale, __ldg(xnorm + (size_t)b0 * D + i));
    __syncthreads();
    if (k >= K) continue;
    float acc[kRowsPerPass];
#pragma unroll
    for (int r = 0; r < kRowsPerPass; ++r) acc[r] = 0.f;
    if (vec) {
      for (int d = lane * 4; d < D; d += 128) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(trow + d));
#pragma unroll
        for (int r = 0; r < kRowsPerPass; ++r) {
          if (r < nb) {
            const float4 xv = *reinterpret_cast<const float4*>(s_x + r * D + d);
            acc[r] = fmaf(xv.x, t.x, acc[r]);
            acc[r] = fmaf(xv.y, t.y, acc[r]);
            acc[r] = fmaf(xv.z, t.z, acc[r]);
            acc[r] = fmaf(xv.w, t.w, acc[r]);
          }
        }
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        const float t = __ldg(trow + d);
#pragma unroll
        for (int r = 0; r < kRowsPerPass; ++r)
          if (r < nb) acc[r] = fmaf(s_x[r * D + d], t, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerPass; ++r) {
      if (r < nb) {
        const float v = warp_sum(acc[r]);
        if (lane == 0) logits[(size_t)(b0 + r) * K + k] = v;
      }
    }
  }
}

// Per-row softmax, entropy -sum p*log(p+1e-10), first-index argmax.
__global__ void __launch_bounds__(256)
    row_stats_kernel(const float* __restrict__ logits, int K, long long ld, float* __restrict__ prob,
                     float* __restrict__ entropy, int* __restrict__ argmax) {
  __shared__ float s_f[8];
  __shared__ unsigned s_u[8];
  __shared__ float s_bcast[2];
  __shared__ unsigned s_arg;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const float* row = logits + (size_t)b * ld;
  // max + first argmax
  float mx = -INFINITY;
  unsigned mi = 0xffffffffu;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float v = row[k];
    if (v > mx) mx = v, mi = k;
  }
  const float wmx = warp_max(mx);
  const unsigned wmi = __reduce_min_sync(kFullMask, mx == wmx ? mi : 0xffffffffu);
  if (lane == 0) s_f[warp] = wmx, s_u[warp] = wmi;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = s_f[0];
    unsigned a = s_u[0];
    for (int w = 1; w < W; ++w)
      if (s_f[w] > m || (s_f[w] == m && s_u[w] < a)) m = s_f[w], a = s_u[w];
    s_bcast[0] = m;
    s_arg = a;
  }
  __syncthreads();
  mx = s_bcast[0];
  float se = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) se += expf(row[k] - mx);
  se = warp_sum(se);
  __syncthreads();
  if (lane == 0) s_f[warp] = se;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < W; ++w) t += s_f[w];
    s_bcast[1] = t;
  }
  __syncthreads();
  const float denom = s_bcast[1];
  float ent = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float p = __fdiv_rn(expf(row[k] - mx), denom);
    if (prob) prob[(size_t)b * K + k] = p;
    ent += p * logf(p + 1e-10f);
  }
  ent = warp_sum(ent);
  __syncthreads();
  if (lane == 0) s_f[warp] = ent;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < W; ++w) t += s_f[w];
    if (entropy) entropy[b] = -t;
    if (argmax) argmax[b] = (int)s_arg;
  }
}

}  // namespace

// xnorm = x / |x| and the (hi, lo) tf32 pair of scale * xnorm (the A operand of the tensor-core head)
__global__ void __launch_bounds__(256) l2norm_scale_split_kernel(const float* __restrict__ x, int D, float scale,
                                                                 float* __restrict__ xnorm, float* __restrict__ hi,
                                                                 float* __restrict__ lo) {
  __shared__ float s_part[8];
  const int b = blockIdx.x;
  const float* row = x + (size_t)b * D;
  float ss = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float v = __ldg(row + d);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_part[w];
  const float nrm = sqrtf(tot);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float xn = __fdiv_rn(__ldg(row + d), nrm);
    xnorm[(size_t)b * D + d] = xn;
    const float sv = __fmul_rn(scale, xn);            // the reference scales x before the contraction
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(sv));
    const float h = __uint_as_float(r);
    hi[(size_t)b * D + d] = h;
    lo[(size_t)b * D + d] = sv - h;
  }
}

int launch_row_stats(const float* logits, int B, int K, long long ld, float* prob, float* entropy, int* argmax,
                     cudaStream_t st) {
  row_stats_kernel<<<B, 256, 0, st>>>(logits, K, ld, prob, entropy, argmax);
  return check_launch("ua_head(row_stats)");
}

int launch_l2norm(const float* x, int B, int D, float* out, cudaStream_t st) {
  l2norm_kernel<<<B, 256, 0, st>>>(x, D, out);
  return check_launch("ua_head(l2norm)");
}

}  // namespace ua

extern "C" int ua_head_f32(const float* x, int B, int D, const float* text, int num_text, int K, float scale,
                           float* out_xnorm, float* out_logits, float* out_prob, float* out_entropy,
                           int32_t* out_argmax, void* stream) {
  using namespace ua;
  UA_REQUIRE(x && text && out_logits && out_xnorm, "ua_head_f32: x/text/out_logits/out_xnorm must be non-NULL");
  UA_REQUIRE(B >= 0 && D >= 1 && K >= 1, "ua_head_f32: bad sizes B=%d D=%d K=%d", B, D, K);
  UA_REQUIRE(num_text >= 1 && B % num_text == 0, "ua_head_f32: B=%d must be a multiple of num_text=%d", B, num_text);
  UA_REQUIRE(num_text <= 65535, "ua_head_f32: num_text=%d > 65535", num_text);
  if (B == 0) return UA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_l2norm(x, B, D, out_xnorm, st);
  if (rc != UA_OK) return rc;
  const int W = 8;
  const size_t smem = (size_t)kRowsPerPass * D * sizeof(float);
  UA_UNSUPPORTED(smem > 200 * 1024, "ua_head_f32: D=%d too large", D);
  if (smem > 48 * 1024) cudaFuncSetAttribute(logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int row_blocks = (B / num_text + kRowsPerPass - 1) / kRowsPerPass;
  dim3 grid((K + W - 1) / W, num_text, row_blocks < 1 ? 1 : (row_blocks > 256 ? 256 : row_blocks));
  logits_kernel<<<grid, W * 32, smem, st>>>(out_xnorm, B / num_text, D, text, K, scale, out_logits);
  rc = check_launch("ua_head_f32(logits)");
  if (rc != UA_OK) return rc;
  if (out_prob || out_entropy || out_argmax)
    return launch_row_stats(out_logits, B, K, K, out_prob, out_entropy, out_argmax, st);
  return UA_OK;
}

extern "C" int ua_head_prepare_f32(const float* x, int B, int D, float scale, float* out_xnorm, float* out_hi, float* out_lo,
                                   void* stream) {
  using namespace ua;
  UA_REQUIRE(x && out_xnorm && out_hi && out_lo, "ua_head_prepare_f32: NULL pointer");
  UA_REQUIRE(B >= 0 && D >= 1, "ua_head_prepare_f32: bad sizes B=%d D=%d", B, D);
  if (B == 0) return UA_OK;
  l2norm_scale_split_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, D, scale, out_xnorm, out_hi, out_lo);
  return check_launch("ua_head_prepare_f32");
}

extern "C" int ua_row_stats_f32(const float* logits, int B, int K, long long ld, float* out_prob, float* out_entropy,
                                int32_t* out_argmax, void* stream) {
  using namespace ua;
  UA_REQUIRE(logits && B >= 0 && K >= 1 && ld >= K, "ua_row_stats_f32: bad arguments (B=%d K=%d ld=%lld)", B, K, ld);
  if (B == 0) return UA_OK;
  return launch_row_stats(logits, B, K, ld, out_prob, out_entropy, out_argmax, (cudaStream_t)stream);
}
