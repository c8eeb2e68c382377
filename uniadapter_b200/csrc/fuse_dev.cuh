// Device-side pieces of the logit fusion (Uni_Adapter.py:491-521), shared by fuse.cu (one launch per step) and by the
// class-sharded sample step (modedota_sample.cu), whose last CTA fuses the gathered rows itself.
#pragma once
#include "common.cuh"

namespace ua {

// s_tmp: >= 32 floats of shared memory; every thread of the CTA must call.
__device__ __forceinline__ float block_sum(float v, float* s_tmp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) s_tmp[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < W; ++w) t += s_tmp[w];
  return t;
}

__device__ __forceinline__ float block_max(float v, float* s_tmp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) s_tmp[warp] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int w = 0; w < W; ++w) t = fmaxf(t, s_tmp[w]);
  return t;
}

// pair versions (same per-value order as block_sum / block_max): s_tmp needs 64 floats
__device__ __forceinline__ void block_sum2(float& a, float& b, float* s_tmp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  a = warp_sum(a), b = warp_sum(b);
  __syncthreads();
  if (lane == 0) s_tmp[warp] = a, s_tmp[32 + warp] = b;
  __syncthreads();
  float ta = 0.f, tb = 0.f;
  for (int w = 0; w < W; ++w) ta += s_tmp[w], tb += s_tmp[32 + w];
  a = ta, b = tb;
}
__device__ __forceinline__ void block_max2(float& a, float& b, float* s_tmp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  a = warp_max(a), b = warp_max(b);
  __syncthreads();
  if (lane == 0) s_tmp[warp] = a, s_tmp[32 + warp] = b;
  __syncthreads();
  float ta = -INFINITY, tb = -INFINITY;
  for (int w = 0; w < W; ++w) ta = fmaxf(ta, s_tmp[w]), tb = fmaxf(tb, s_tmp[32 + w]);
  a = ta, b = tb;
}

// -sum softmax(v) * log(softmax(v) + 1e-10) over K entries produced by `get(k)`
template <typename F>
__device__ __forceinline__ float softmax_entropy(F get, int K, float* s_tmp) {
  float mx = -INFINITY;
  for (int k = threadIdx.x; k < K; k += blockDim.x) mx = fmaxf(mx, get(k));
  mx = block_max(mx, s_tmp);
  float se = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) se += expf(get(k) - mx);
  se = block_sum(se, s_tmp);
  float ent = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float p = __fdiv_rn(expf(get(k) - mx), se);
    ent += p * logf(p + 1e-10f);
  }
  return -block_sum(ent, s_tmp);
}

// The two blend weights of the MODE-DOTA fusion: wc = a/(a+b), wd = b/(wc+b) with a = 1/(H_clip+1e-3),
// b = 1/(H_dota+1e-3) -- the second weight is normalised with the ALREADY-normalised first one (Uni_Adapter.py:512-513).
__device__ __forceinline__ void entropy_weights(float hc, float hd, float& wc, float& wd) {
  const float a = __fdiv_rn(1.f, __fadd_rn(hc, 1e-3f));
  const float b = __fdiv_rn(1.f, __fadd_rn(hd, 1e-3f));
  wc = __fdiv_rn(a, __fadd_rn(a, b));
  wd = __fdiv_rn(b, __fadd_rn(wc, b));
}

// w = min(rho * (sum(c) / count) / batch, eta)   (Uni_Adapter.py:491)
__device__ __forceinline__ float cache_weight(float csum, float count_total, float rho, float batch, float eta) {
  const float cmean = __fdiv_rn(csum, count_total);
  return fminf(__fdiv_rn(__fmul_rn(cmean, rho), batch), eta);
}

}  // namespace ua
