// Grouping kernels of the point tokenizer for sm_100a.
//
//   knn_group_kernel : k nearest points of every FPS centre under the reference's expanded squared distance,
//                      fused with gather, centre subtraction and channel concat
//                      (models/point_encoder.py:17-49,99-127, models/ulip/pointbert/dvae.py:116-181).
//   ball_group_kernel: first-nsample-in-index-order ball query, fused with the same epilogue
//                      (models/openshape/pointnet_util.py:89-146).
//
// One warp owns one centre. A CTA's warps share point tiles staged in shared memory (coordinates AoS, stride-3
// reads are bank-conflict free, plus the per-point squared norm). The (G,N) distance matrix of the reference is
// never written. kNN selection is a streaming filter: candidates below the running threshold are ballot-compacted
// into a per-warp shared-memory buffer of 64-bit (distance,index) keys; when the buffer fills, a warp-wide bitonic
// sort keeps the k best and tightens the threshold (O(log(N/k)) sorts per centre).
#include "common.cuh"

namespace ua {

int g_knn_warps = 0;  // tuning override (0 = heuristic)

namespace {

constexpr int kTilePoints = 2048;
constexpr uint64_t kKeyMax = ~0ull;

__device__ __forceinline__ void stage_tile(const float* __restrict__ cloud, int t0, int tp, float* s_xyz,
                                           float* s_pn) {
  __syncthreads();  // previous tile fully consumed
  const float* src = cloud + (size_t)3 * t0;
  for (int i = threadIdx.x; i < 3 * tp; i += blockDim.x) s_xyz[i] = __ldg(src + i);
  __syncthreads();
  for (int p = threadIdx.x; p < tp; p += blockDim.x)
    s_pn[p] = sqnorm_nofma(s_xyz[3 * p], s_xyz[3 * p + 1], s_xyz[3 * p + 2]);
  __syncthreads();
}

// Ascending bitonic sort of CAP 64-bit keys in shared memory by one warp.
template <int CAP>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t* buf, int lane) {
#pragma unroll 1
  for (int size = 2; size <= CAP; size <<= 1) {
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int t = lane; t < CAP / 2; t += 32) {
        const int i = 2 * t - (t & (stride - 1));
        const int j = i + stride;
        const bool asc = (i & size) == 0;
        const uint64_t a = buf[i], b = buf[j];
        if ((a > b) == asc) {
          buf[i] = b;
          buf[j] = a;
        }
      }
      __syncwarp();
    }
  }
}

template <int CAP, typename IdxT>
__global__ void __launch_bounds__(256)
    knn_group_kernel(const float* __restrict__ xyz, const float* __restrict__ rgb, const float* __restrict__ centers,
                     int N, int G, int k, IdxT* __restrict__ out_idx, float* __restrict__ out_neigh,
                     float* __restrict__ out_feat) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int W = blockDim.x >> 5;
  uint64_t* s_buf = reinterpret_cast<uint64_t*>(s_raw);                    // [W][CAP]
  float* s_xyz = reinterpret_cast<float*>(s_raw + (size_t)W * CAP * 8);    // [3*kTilePoints]
  float* s_pn = s_xyz + 3 * kTilePoints;                                   // [kTilePoints]

  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.x * W + warp;
  const bool active = g < G;
  const float* cloud = xyz + (size_t)b * N * 3;
  uint64_t* buf = s_buf + (size_t)warp * CAP;

  float cx = 0.f, cy = 0.f, cz = 0.f, cn = 0.f;
  if (active) {
    const float* c = centers + ((size_t)b * G + g) * 3;
    cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
    cn = sqnorm_nofma(cx, cy, cz);
  }
  int count = 0;
  uint64_t thr = kKeyMax;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int t0 = 0; t0 < N; t0 += kTilePoints) {
    const int tp = min(kTilePoints, N - t0);
    stage_tile(cloud, t0, tp, s_xyz, s_pn);
    if (!active) continue;
    for (int base = 0; base < tp; base += 32) {
      const int p = base + lane;
      uint64_t key = kKeyMax;
      if (p < tp) {
        const float d = expanded_sqdist(cx, cy, cz, cn, s_xyz[3 * p], s_xyz[3 * p + 1], s_xyz[3 * p + 2], s_pn[p]);
        key = ((uint64_t)float_to_ordered(d) << 32) | (uint32_t)(t0 + p);
      }
      const bool take = key < thr;  // kKeyMax is never < thr
      const unsigned m = __ballot_sync(kFullMask, take);
      if (m) {
        if (take) buf[count + __popc(m & lt_mask)] = key;
        count += __popc(m);
        if (count > CAP - 32) {  // not enough room for another full batch: keep the k best, tighten the threshold
          __syncwarp();
          for (int i = count + lane; i < CAP; i += 32) buf[i] = kKeyMax;
          __syncwarp();
          warp_bitonic_sort<CAP>(buf, lane);
          count = min(count, k);
          thr = count == k ? buf[k - 1] : kKeyMax;
        }
      }
    }
  }
  if (!active) return;
  __syncwarp();
  for (int i = count + lane; i < CAP; i += 32) buf[i] = kKeyMax;
  __syncwarp();
  warp_bitonic_sort<CAP>(buf, lane);

  // epilogue: nearest-first indices, gather, centre subtraction, concat
  const size_t row0 = ((size_t)b * G + g) * k;
  for (int j = lane; j < k; j += 32) {
    const uint32_t p = (uint32_t)(buf[j] & 0xffffffffull);
    if (out_idx) out_idx[row0 + j] = (IdxT)p;
    const float* src = cloud + (size_t)3 * p;
    const float nx = __fsub_rn(__ldg(src), cx), ny = __fsub_rn(__ldg(src + 1), cy), nz = __fsub_rn(__ldg(src + 2), cz);
    if (out_neigh) {
      float* o = out_neigh + (row0 + j) * 3;
      o[0] = nx, o[1] = ny, o[2] = nz;
    }
    if (out_feat) {
      const float* col = rgb + ((size_t)b * N + p) * 3;
      float2* o = reinterpret_cast<float2*>(out_feat + (row0 + j) * 6);
      o[0] = make_float2(nx, ny);
      o[1] = make_float2(nz, __ldg(col));
      o[2] = make_float2(__ldg(col + 1), __ldg(col + 2));
    }
  }
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
    ball_group_kernel(const float* __restrict__ xyz, const float* __restrict__ feat, int C,
                      const float* __restrict__ centers, int N, int S, float radius2, int nsample,
                      IdxT* __restrict__ out_idx, float* __restrict__ out_new_points) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int W = blockDim.x >> 5;
  int* s_sel = reinterpret_cast<int*>(s_raw);                                   // [W][nsample]
  float* s_xyz = reinterpret_cast<float*>(s_raw + (((size_t)W * nsample * 4 + 15) / 16) * 16);
  float* s_pn = s_xyz + 3 * kTilePoints;

  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.x * W + warp;
  const bool active = g < S;
  const float* cloud = xyz + (size_t)b * N * 3;
  int* sel = s_sel + (size_t)warp * nsample;

  float cx = 0.f, cy = 0.f, cz = 0.f, cn = 0.f;
  if (active) {
    const float* c = centers + ((size_t)b * S + g) * 3;
    cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
    cn = sqnorm_nofma(cx, cy, cz);
  }
  int count = 0;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int t0 = 0; t0 < N; t0 += kTilePoints) {
    // every warp of the CTA must take part in staging even when its own ball is already full
    const int tp = min(kTilePoints, N - t0);
    stage_tile(cloud, t0, tp, s_xyz, s_pn);
    if (!active || count >= nsample) continue;
    for (int base = 0; base < tp && count < nsample; base += 32) {
      const int p = base + lane;
      bool inside = false;
      if (p < tp) {
        const float d = expanded_sqdist(cx, cy, cz, cn, s_xyz[3 * p], s_xyz[3 * p + 1], s_xyz[3 * p + 2], s_pn[p]);
        inside = !(d > radius2);  // the reference marks d > r^2 as outside
      }
      const unsigned m = __ballot_sync(kFullMask, inside);
      if (m) {
        const int pos = count + __popc(m & lt_mask);
        if (inside && pos < nsample) sel[pos] = t0 + p;
        count += __popc(m);
      }
    }
  }
  if (!active) return;
  __syncwarp();
  count = min(count, nsample);
  const int first = count > 0 ? sel[0] : N - 1;
  const size_t row0 = ((size_t)b * S + g) * nsample;
  const int CO = 3 + C;
  for (int j = lane; j < nsample; j += 32) {
    const int p = j < count ? sel[j] : first;
    if (out_idx) out_idx[row0 + j] = (IdxT)p;
    if (out_new_points) {
      const float* src = cloud + (size_t)3 * p;
      float* o = out_new_points + (row0 + j) * CO;
      o[0] = __fsub_rn(__ldg(src), cx);
      o[1] = __fsub_rn(__ldg(src + 1), cy);
      o[2] = __fsub_rn(__ldg(src + 2), cz);
      const float* f = feat + ((size_t)b * N + p) * C;
      for (int ch = 0; ch < C; ++ch) o[3 + ch] = __ldg(f + ch);
    }
  }
}

__global__ void gather_points_kernel(const float* __restrict__ in, const int* __restrict__ idx, int C, int N, int G,
                                     float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
    const int p = idx[(size_t)b * G + g];
    out[((size_t)b * C + c) * G + g] = __ldg(in + ((size_t)b * C + c) * N + p);
  }
}

int pick_warps(int B, int G) {
  if (g_knn_warps > 0) return g_knn_warps;
  // enough CTAs to cover the 148 SMs a few times over, as many centres per staged tile as that allows
  const long long centres = (long long)B * G;
  if (centres >= 8LL * 4 * kNumSMs) return 8;
  if (centres >= 4LL * 2 * kNumSMs) return 4;
  return 2;
}

template <int CAP, typename IdxT>
int launch_knn(const float* xyz, const float* rgb, const float* centers, int B, int N, int G, int k, void* out_idx,
               float* out_neigh, float* out_feat, cudaStream_t st) {
  const int W = pick_warps(B, G);
  const size_t smem = (size_t)W * CAP * 8 + (size_t)4 * kTilePoints * sizeof(float);
  auto kern = knn_group_kernel<CAP, IdxT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("ua_knn_group_f32: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return UA_ERR_CUDA;
    }
  }
  dim3 grid((G + W - 1) / W, B);
  kern<<<grid, W * 32, smem, st>>>(xyz, rgb, centers, N, G, k, (IdxT*)out_idx, out_neigh, out_feat);
  return check_launch("ua_knn_group_f32");
}

template <typename IdxT>
int dispatch_knn(const float* xyz, const float* rgb, const float* centers, int B, int N, int G, int k, void* out_idx,
                 float* out_neigh, float* out_feat, cudaStream_t st) {
  if (k <= 48) return launch_knn<128, IdxT>(xyz, rgb, centers, B, N, G, k, out_idx, out_neigh, out_feat, st);
  if (k <= 112) return launch_knn<256, IdxT>(xyz, rgb, centers, B, N, G, k, out_idx, out_neigh, out_feat, st);
  return launch_knn<512, IdxT>(xyz, rgb, centers, B, N, G, k, out_idx, out_neigh, out_feat, st);
}

}  // namespace
}  // namespace ua

extern "C" int ua_knn_group_f32(const float* xyz, const float* rgb, const float* centers, int B, int N, int G, int k,
                                void* out_idx, int idx_is_i64, float* out_neigh, float* out_feat, void* stream) {
  using namespace ua;
  UA_REQUIRE(xyz && centers, "ua_knn_group_f32: xyz/centers is NULL");
  UA_REQUIRE(B >= 0 && N >= 1 && G >= 1, "ua_knn_group_f32: bad sizes B=%d N=%d G=%d", B, N, G);
  UA_REQUIRE(k >= 1 && k <= N, "ua_knn_group_f32: k=%d must be in [1, N=%d]", k, N);
  UA_UNSUPPORTED(k > 128, "ua_knn_group_f32: k=%d > 128 is not supported", k);
  UA_REQUIRE(out_feat == nullptr || rgb != nullptr, "ua_knn_group_f32: out_feat requires rgb");
  UA_REQUIRE(B <= 65535, "ua_knn_group_f32: B=%d > 65535", B);
  if (B == 0) return UA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  return idx_is_i64 ? dispatch_knn<long long>(xyz, rgb, centers, B, N, G, k, out_idx, out_neigh, out_feat, st)
                    : dispatch_knn<int>(xyz, rgb, centers, B, N, G, k, out_idx, out_neigh, out_feat, st);
}

extern "C" int ua_ball_group_f32(const float* xyz, const float* feat, int C, const float* centers, int B, int N, int S,
                                 float radius2, int nsample, void* out_idx, int idx_is_i64, float* out_new_points,
                                 void* stream) {
  using namespace ua;
  UA_REQUIRE(xyz && centers, "ua_ball_group_f32: xyz/centers is NULL");
  UA_REQUIRE(B >= 0 && N >= 1 && S >= 1 && nsample >= 1, "ua_ball_group_f32: bad sizes B=%d N=%d S=%d nsample=%d", B,
             N, S, nsample);
  UA_REQUIRE(C >= 0 && (C == 0 || feat != nullptr), "ua_ball_group_f32: C=%d needs feat", C);
  UA_UNSUPPORTED(nsample > 1024, "ua_ball_group_f32: nsample=%d > 1024 is not supported", nsample);
  UA_REQUIRE(B <= 65535, "ua_ball_group_f32: B=%d > 65535", B);
  if (B == 0) return UA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int W = pick_warps(B, S);
  const size_t smem = (((size_t)W * nsample * 4 + 15) / 16) * 16 + (size_t)4 * kTilePoints * sizeof(float);
  dim3 grid((S + W - 1) / W, B);
  if (idx_is_i64) {
    auto kern = ball_group_kernel<long long>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, W * 32, smem, st>>>(xyz, feat, C, centers, N, S, radius2, nsample, (long long*)out_idx,
                                     out_new_points);
  } else {
    auto kern = ball_group_kernel<int>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, W * 32, smem, st>>>(xyz, feat, C, centers, N, S, radius2, nsample, (int*)out_idx, out_new_points);
  }
  return check_launch("ua_ball_group_f32");
}

extern "C" int ua_gather_points_f32(const float* in, const int32_t* idx, int B, int C, int N, int G, float* out,
                                    void* stream) {
  using namespace ua;
  UA_REQUIRE(in && idx && out, "ua_gather_points_f32: NULL pointer");
  UA_REQUIRE(B >= 0 && C >= 1 && N >= 1 && G >= 1, "ua_gather_points_f32: bad sizes");
  UA_REQUIRE(B <= 65535 && C <= 65535, "ua_gather_points_f32: B/C > 65535");
  if (B == 0) return UA_OK;
  dim3 grid((G + 255) / 256, C, B);
  gather_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, idx, C, N, G, out);
  return check_launch("ua_gather_points_f32");
}
