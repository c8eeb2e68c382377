// Fusion of zero-shot logits with cache logits, one CTA per row.
//   mode 1 (MODE-DOTA, Uni_Adapter.py:491-521): d = w*dota; entropy-weighted blend of clip and d, where the second
//           weight is normalised with the ALREADY-normalised first weight (the reference's own arithmetic).
//   mode 0 (DOTA, dota_mixture.py:289-293): final = clip + w*dota, the product taken in fp16 when dota is fp16
//           (torch type promotion of a 0-dim fp32 tensor times a half tensor).
//   w = min(rho * mean(c) / batch, eta)
#include "common.cuh"
#include "fuse_dev.cuh"

namespace ua {
namespace {

__global__ void __launch_bounds__(256)
    fuse_kernel(const float* __restrict__ clip, const void* __restrict__ dota, int dota_is_f16, int K,
                const float* __restrict__ c, int count_c, int c_row_stride, float c_sum_override, float c_count_total,
                float rho,
                float eta, float batch, int mode, float* __restrict__ out_final, int* __restrict__ out_argmax,
                float* __restrict__ out_scaled) {
  __shared__ float s_tmp[32];
  __shared__ float s_best[8];
  __shared__ unsigned s_besti[8];
  const int r = blockIdx.x;
  const float* crow = clip + (size_t)r * K;
  const float* drow32 = dota_is_f16 ? nullptr : reinterpret_cast<const float*>(dota) + (size_t)r * K;
  const __half* drow16 = dota_is_f16 ? reinterpret_cast<const __half*>(dota) + (size_t)r * K : nullptr;

  float csum = c_sum_override;
  if (!(c_sum_override >= 0.f)) {
    float part = 0.f;
    const float* crow_c = c + (size_t)r * c_row_stride;
    for (int i = threadIdx.x; i < count_c; i += blockDim.x) part += __ldg(crow_c + i);
    csum = block_sum(part, s_tmp);
  }
  const float w = cache_weight(csum, c_count_total, rho, batch, eta);

  auto scaled = [&](int k) -> float {
    if (dota_is_f16) {
      const float wh = __half2float(__float2half_rn(w));
      return __half2float(__float2half_rn(wh * __half2float(drow16[k])));
    }
    return __fmul_rn(w, drow32[k]);
  };

  float wc = 1.f, wd = 1.f;
  if (mode == 1) {
    const float hc = softmax_entropy([&](int k) { return crow[k]; }, K, s_tmp);
    const float hd = softmax_entropy(scaled, K, s_tmp);
    entropy_weights(hc, hd, wc, wd);      // sic: the second weight is normalised with the updated clip weight
  }
  float best = -INFINITY;
  unsigned besti = 0xffffffffu;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float d = scaled(k);
    const float f = mode == 1 ? __fadd_rn(__fmul_rn(wc, crow[k]), __fmul_rn(wd, d)) : __fadd_rn(crow[k], d);
    out_final[(size_t)r * K + k] = f;
    if (out_scaled) out_scaled[(size_t)r * K + k] = d;
    if (f > best) best = f, besti = k;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const float wmx = warp_max(best);
  const unsigned wmi = __reduce_min_sync(kFullMask, best == wmx ? besti : 0xffffffffu);
  __syncthreads();
  if (lane == 0) s_best[warp] = wmx, s_besti[warp] = wmi;
  __syncthreads();
  if (threadIdx.x == 0 && out_argmax) {
    float m = s_best[0];
    unsigned a = s_besti[0];
    for (int q = 1; q < W; ++q)
      if (s_best[q] > m || (s_best[q] == m && s_besti[q] < a)) m = s_best[q], a = s_besti[q];
    out_argmax[r] = (int)a;
  }
}

}  // namespace
}  // namespace ua

extern "C" int ua_fuse_logits_f32(const float* clip_logits, const void* dota_logits, int dota_is_f16, int R, int K,
                                  const float* c, int count_c, int c_row_stride, float c_sum_override,
                                  float c_count_total, float rho, float eta, float batch, int mode, float* out_final,
                                  int32_t* out_argmax, float* out_scaled_dota, void* stream) {
  using namespace ua;
  UA_REQUIRE(clip_logits && dota_logits && out_final, "ua_fuse_logits_f32: NULL pointer");
  UA_REQUIRE(R >= 1 && K >= 1, "ua_fuse_logits_f32: bad sizes R=%d K=%d", R, K);
  UA_REQUIRE(c_sum_override >= 0.f || (c && count_c >= 1), "ua_fuse_logits_f32: need c[] or c_sum_override");
  UA_REQUIRE(c_count_total > 0.f && batch > 0.f, "ua_fuse_logits_f32: c_count_total and batch must be > 0");
  UA_REQUIRE(mode == 0 || mode == 1, "ua_fuse_logits_f32: mode must be 0 or 1");
  fuse_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(clip_logits, dota_logits, dota_is_f16, K, c, count_c, c_row_stride,
                                                   c_sum_override, c_count_total, rho, eta, batch, mode, out_final,
                                                   out_argmax, out_scaled_dota);
  return check_launch("ua_fuse_logits_f32");
}
