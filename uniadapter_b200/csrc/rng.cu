// Per-stream random inputs of the lock-step engine, drawn on the device (CUDA-graph safe): the Gaussian jitter of the
// augmented view (Uni_Adapter.py:420-421, pc + 0.05 * randn_like(pc)) and the random FPS start indices of both views
// (models/ulip/pointbert/misc.py:52, models/openshape/pointnet_util.py:77).
//
// Counter-based (Philox4x32-10): stream s draws from key = seeds[s] and counter = (element block, step, purpose), so
// what a stream sees depends only on its own seed and on how many steps IT has taken -- not on which other streams
// share the GPU, nor on the world size (the per-stream convention of SURVEY H3). The step counter lives on the device
// and is advanced by the kernel itself, so a captured graph keeps counting.
#include "common.cuh"

namespace ua {
namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0, key.y += W1;
  }
  return ctr;
}

// two uniforms in (0,1] -> two independent N(0,1) (Box-Muller)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
  const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

__global__ void __launch_bounds__(256) stream_rng_kernel(const long long* __restrict__ seeds,
                                                         long long* __restrict__ step_ptr, int S, long long per_stream,
                                                         float* __restrict__ noise, long long* __restrict__ start,
                                                         int n_range, unsigned* __restrict__ done) {
  const long long step = *step_ptr;
  const int s = blockIdx.y;
  const uint2 key = make_uint2((uint32_t)seeds[s], (uint32_t)((unsigned long long)seeds[s] >> 32));
  const long long quads = (per_stream + 3) / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)step, 0u), key);
    const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
    const float v[4] = {a.x, a.y, b.x, b.y};
    float* dst = noise + (size_t)s * per_stream + 4 * q;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * q + i < per_stream) dst[i] = v[i];
  }
  if (start && blockIdx.x == 0 && threadIdx.x == 0) {
    const uint4 r = philox4x32_10(make_uint4(0u, 0u, (uint32_t)step, 1u), key);
    start[s] = (long long)(r.x % (uint32_t)n_range);          // first view
    start[S + s] = (long long)(r.y % (uint32_t)n_range);      // jittered view
  }
  // the last CTA to finish advances the step counter (everybody has read it by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned total = gridDim.x * gridDim.y;
    if (atomicAdd(done, 1u) == total - 1) {
      *done = 0u;
      *step_ptr = step + 1;
    }
  }
}

}  // namespace
}  // namespace ua

extern "C" int ua_stream_rng_f32(const int64_t* seeds, int64_t* step, int S, int64_t per_stream, float* noise,
                                 int64_t* start_idx, int n_range, uint32_t* done_counter, void* stream) {
  using namespace ua;
  UA_REQUIRE(seeds && step && noise && done_counter, "ua_stream_rng_f32: NULL pointer");
  UA_REQUIRE(S >= 1 && S <= 65535 && per_stream >= 1, "ua_stream_rng_f32: bad sizes S=%d per_stream=%lld", S,
             (long long)per_stream);
  UA_REQUIRE(!start_idx || n_range >= 1, "ua_stream_rng_f32: n_range=%d", n_range);
  const long long quads = (per_stream + 3) / 4;
  int gx = (int)((quads + 255) / 256);
  const int cap = (2 * kNumSMs + S - 1) / S;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  stream_rng_kernel<<<dim3((unsigned)gx, (unsigned)S), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)seeds, (long long*)step, S, (long long)per_stream, noise, (long long*)start_idx, n_range,
      done_counter);
  return check_launch("ua_stream_rng_f32");
}
