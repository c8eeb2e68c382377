// Farthest-point sampling for sm_100a: one cloud per CTA, the cloud's running-min distances and coordinates
// live in registers for the whole G-iteration loop; each iteration costs one block barrier and four REDUX
// warp reductions (max of the distance bits, then min index among the maxima).
//
// Semantics follow models/ulip/pointbert/misc.py:40-60 and models/openshape/pointnet_util.py:64-86 of the
// reference: dist = sum((xyz - centroid)**2, -1) rounded per operation, distance = min(distance, dist) from an
// initial 1e10, farthest = first index of the maximum. The Uni3D path (models/point_encoder.py:7-14, un-vendored
// pointnet2_ops CUDA) maps to start index 0 and has two arithmetic variants: the torch order above (+ optional skip of
// near-origin points), and `PN2`, the published pointnet2_ops kernel's own arithmetic (furthest_point_sampling_kernel,
// pointnet2_ops_lib 3.0.0): distances with the FMA contraction nvcc gives the upstream source,
// d = fma(dz,dz, fma(dx,dx, dy*dy)), points with fma(z,z, fma(x,x, y*y)) <= 1e-3 (compared in double) never take part,
// and ties follow upstream's reduction -- thread `k mod block_size` owns point k, scans its points with a strict '>',
// and the pairwise tree keeps the left operand on ties, which favours the smallest bit-reversed thread id. Here
// ownership is different (registers), so every candidate carries the permuted index
// comp(k) = bitrev(k mod bs) * ceil(N / bs) + k / bs and ties take the lowest comp; bs = largest power of two
// <= min(N, 512) as upstream's opt_n_threads.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace ua {

int g_fps_threads = 0;  // tuning override (0 = heuristic)
int g_fps_cluster = 0;  // tuning: -1 disables the cluster path, N > 0 = smallest cloud that takes it (default 2049)

namespace {

constexpr float kFpsInit = 1e10f;

__device__ __forceinline__ float sqdist_pn2(float px, float py, float pz, float cx, float cy, float cz) {
  const float dx = __fsub_rn(px, cx), dy = __fsub_rn(py, cy), dz = __fsub_rn(pz, cz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
__device__ __forceinline__ bool pn2_skipped(float x, float y, float z) {
  return (double)__fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y))) <= 1e-3;
}
// block / warp keys: 0 = no candidate (every owned point skipped), else distance bits + 1 (distances are >= 0)
__device__ __forceinline__ uint32_t pn2_key(float best) { return best < 0.f ? 0u : __float_as_uint(best) + 1u; }

// pointnet2_ops tie order. Upstream thread t = k mod bs owns point k and keeps its FIRST maximum; the pairwise tree
// (t, t + bs/2), (t, t + bs/4), ... (0, 1) keeps the LEFT operand on ties, so among equal maxima the winner is the thread
// with the smallest BIT-REVERSED id (the last level lets even threads beat odd ones, the one before decides bit 1, ...):
// comp(k) = (bitrev_lg(k mod bs), k / bs) packed into one word with bs = 2^lg; ties go to the lowest comp.
struct Pn2Order {
  int lg, lq;      // bs = 2^lg; 2^lq >= ceil(N / bs): a power of two keeps the (thread, round) order and decodes by shifts
  __device__ __forceinline__ uint32_t rev(uint32_t t) const { return lg ? (__brev(t) >> (32 - lg)) : 0u; }
  __device__ __forceinline__ uint32_t comp(uint32_t k) const { return (rev(k & ((1u << lg) - 1u)) << lq) | (k >> lg); }
  __device__ __forceinline__ uint32_t index(uint32_t c) const { return ((c & ((1u << lq) - 1u)) << lg) + rev(c >> lq); }
};
__device__ __forceinline__ void block_argmax(uint32_t bits, uint32_t idx, uint2 (*s_red)[32], int buf, int lane,
                                             int warp, int nwarps, uint32_t& out_idx) {
  // warp level: maximum of the (non-negative) float bit patterns, lowest index among the maxima
  const uint32_t wmax = __reduce_max_sync(kFullMask, bits);
  const uint32_t widx = __reduce_min_sync(kFullMask, bits == wmax ? idx : 0xffffffffu);
  if (lane == 0) s_red[buf][warp] = make_uint2(wmax, widx);
  __syncthreads();
  const uint2 v = lane < nwarps ? s_red[buf][lane] : make_uint2(0u, 0xffffffffu);
  const uint32_t bmax = __reduce_max_sync(kFullMask, v.x);
  out_idx = __reduce_min_sync(kFullMask, v.x == bmax ? v.y : 0xffffffffu);
}

// the same with SIGNED keys (pointnet2 mode of the register kernel: float bit patterns, -1.0f = no candidate)
__device__ __forceinline__ void block_argmax_signed(int bits, uint32_t idx, uint2 (*s_red)[32], int buf, int lane, int warp,
                                                    int nwarps, uint32_t& out_idx) {
  const int wmax = __reduce_max_sync(kFullMask, bits);
  const uint32_t widx = __reduce_min_sync(kFullMask, bits == wmax ? idx : 0xffffffffu);
  if (lane == 0) s_red[buf][warp] = make_uint2((uint32_t)wmax, widx);
  __syncthreads();
  const uint2 v = lane < nwarps ? s_red[buf][lane] : make_uint2(0x80000000u, 0xffffffffu);
  const int bmax = __reduce_max_sync(kFullMask, (int)v.x);
  out_idx = __reduce_min_sync(kFullMask, (int)v.x == bmax ? v.y : 0xffffffffu);
}

// s_sel holds point indices -- or, in the pointnet2 mode, the tie-order words comp(k), which is also the position of
// point k in the shared-memory copy of the cloud (no decode on the critical path of the sampling loop)
template <typename IdxT, bool PN2 = false>
__device__ __forceinline__ void write_selection(const int* s_sel, const float* cloud_smem, const float* cloud_gmem,
                                                int b, int G, IdxT* out_idx, float* out_centers, Pn2Order ord = Pn2Order()) {
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    const int p = PN2 ? (int)ord.index((uint32_t)s_sel[i]) : s_sel[i];
    if (out_idx) out_idx[(size_t)b * G + i] = (IdxT)p;
  }
  if (out_centers) {
    for (int i = threadIdx.x; i < 3 * G; i += blockDim.x) {
      const int p = s_sel[i / 3];
      const int ch = i - 3 * (i / 3);
      out_centers[(size_t)b * G * 3 + i] = cloud_smem ? cloud_smem[3 * p + ch] : cloud_gmem[3 * (size_t)p + ch];
    }
  }
}

// Register-resident path: N <= PPT * blockDim.x, cloud also staged in shared memory for the centroid fetch.
// (Measured and rejected: thread-blocked point ownership with ballot + find-first-set instead of the second REDUX of
// each reduction level - 111 us instead of 91 us at N = 1024: REDUX.MIN is cheaper than vote + ffs + shuffle.)
template <int PPT, typename IdxT, bool PN2>
__global__ void __launch_bounds__(PPT == 12 ? 896 : 1024, 1)
    fps_reg_kernel(const float* __restrict__ xyz, int N, int G, const long long* __restrict__ start_idx,
                   int skip_small, Pn2Order ord, IdxT* __restrict__ out_idx, float* __restrict__ out_centers) {
  extern __shared__ __align__(16) float s_dyn[];
  const int slots = PN2 ? (1 << (ord.lg + ord.lq)) : N;
  float* s_xyz = s_dyn;                              // [3 * slots]
  int* s_sel = reinterpret_cast<int*>(s_dyn + 3 * slots);  // [G]
  __shared__ uint2 s_red[2][32];

  const int b = blockIdx.x;
  const int T = blockDim.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const float* cloud = xyz + (size_t)b * N * 3;

  if (PN2) {   // the cloud sits in shared memory in tie order: point k at position comp(k)
    for (int i = tid; i < 3 * N; i += T) {
      const int k = i / 3;
      s_xyz[3 * ord.comp((uint32_t)k) + (i - 3 * k)] = __ldg(cloud + i);
    }
  } else {
    for (int i = tid; i < 3 * N; i += T) s_xyz[i] = __ldg(cloud + i);
  }
  __syncthreads();

  float px[PPT], py[PPT], pz[PPT], dmin[PPT];
  uint32_t cj[PN2 ? PPT : 1];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int p = j * T + tid;
    bool valid = p < N;
    const int pos = PN2 ? (int)ord.comp((uint32_t)p) : p;
    if (PN2) cj[j] = (uint32_t)pos;
    px[j] = valid ? s_xyz[3 * pos + 0] : 0.f;
    py[j] = valid ? s_xyz[3 * pos + 1] : 0.f;
    pz[j] = valid ? s_xyz[3 * pos + 2] : 0.f;
    if (PN2) {
      // a point that never takes part keeps the running "distance" -1 (fminf(-1, d) = -1 < any candidate)
      if (valid) valid = !pn2_skipped(px[j], py[j], pz[j]);
      dmin[j] = valid ? kFpsInit : -1.f;
    } else {
      if (skip_small && valid) valid = sqnorm_nofma(px[j], py[j], pz[j]) > 1e-3f;
      // an excluded slot keeps distance 0 forever: it can only tie with exhausted points, and then loses on index
      dmin[j] = valid ? kFpsInit : 0.f;
    }
  }

  long long s0 = start_idx ? start_idx[b] : 0;
  if (s0 < 0) s0 = 0;
  if (s0 >= N) s0 = N - 1;
  uint32_t cur = PN2 ? ord.comp((uint32_t)s0) : (uint32_t)s0;      // PN2: `cur` is a tie-order word = shared-memory position

  for (int i = 0; i < G; ++i) {
    if (tid == 0) s_sel[i] = (int)cur;
    if (i == G - 1) break;
    const float cx = s_xyz[3 * cur + 0], cy = s_xyz[3 * cur + 1], cz = s_xyz[3 * cur + 2];
    float best = -1.f;
    if (PN2) {
      // the thread's maximum and, among its points that share it, the lowest tie-order word: a left-biased pairwise tree
      // over the thread's points (a sequential scan + tie count + rescan sat on the serial path of every iteration)
      float td[PPT];
      uint32_t tc[PPT];
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const float d = sqdist_pn2(px[j], py[j], pz[j], cx, cy, cz);
        dmin[j] = fminf(dmin[j], d);
        td[j] = dmin[j], tc[j] = cj[j];
      }
#pragma unroll
      for (int st = 1; st < PPT; st *= 2)
#pragma unroll
        for (int j = 0; j + st < PPT; j += 2 * st) {
          const bool gt = td[j + st] > td[j], eq = td[j + st] == td[j];
          tc[j] = gt ? tc[j + st] : (eq ? min(tc[j], tc[j + st]) : tc[j]);
          td[j] = gt ? td[j + st] : td[j];
        }
      best = td[0];
      const uint32_t bestc = tc[0];
      // keys are the float bit patterns compared as SIGNED integers: a thread without a candidate holds -1.0f (negative), any
      // distance (>= 0) beats it, and if nobody has a candidate every thread reports word 0 = point 0, as upstream
      block_argmax_signed(__float_as_int(best), best < 0.f ? 0u : bestc, s_red, i & 1, lane, warp, nwarps, cur);
    } else {
      int bestj = 0;
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const float d = sqdist_nofma(px[j], py[j], pz[j], cx, cy, cz);
        const float dm = fminf(dmin[j], d);
        dmin[j] = dm;
        if (dm > best) {
          best = dm;
          bestj = j;
        }
      }
      // (a pairwise tree over the thread's points, as in the pointnet2 branch, was measured here too: no gain at <= 148 clouds,
      // 3 % slower at 1 184)
      block_argmax(__float_as_uint(best), (uint32_t)(bestj * T + tid), s_red, i & 1, lane, warp, nwarps, cur);
    }
  }
  write_selection<IdxT, PN2>(s_sel, s_xyz, nullptr, b, G, out_idx, out_centers, ord);
}


// Cluster path (few large clouds): one thread-block CLUSTER per cloud. A 10 000-point cloud on one SM is issue-bound
// (N x 12 instructions per iteration on four schedulers) and leaves 147 SMs idle at batch 1; here the cloud is cut into
// C contiguous slices, one per CTA of the cluster, each slice register-resident as above (every CTA also keeps the whole
// cloud in shared memory for the centroid fetch). Per iteration every CTA reduces its slice to one candidate and writes
// ONE 64-bit word (tag bit | distance bits, global index) into its OWN shared-memory slot; warp 0 of every CTA polls
// the C slots of the cluster through distributed shared memory (ld.shared::cluster). The word is its own flag: the
// distance bits of a non-negative float leave bit 31 free for a tag that flips on every reuse of the slot, a 64-bit
// access is single-copy atomic, and nothing else has to be ordered with it -- so the loop has no cluster barrier, no
// mbarrier and no fence. The winner is the maximum in rank order (slices are contiguous index ranges, so the lowest
// rank among the maxima holds the lowest index). Slot reuse is safe with two parities: a CTA overwrites its parity
// slot at iteration i+2 only after it has collected every peer's candidate of iteration i+1, and a peer publishes
// that only after it has finished reading iteration i.
// Measured on B200, N = 10 000, 512 samples (one cloud): one CTA 418 us; cluster.sync per iteration 430 us; remote
// st.async completing on the receiver's mbarrier 307 us; plain remote stores + local polling 413-456 us; this pull
// protocol 305 us (319 us when every warp polls; 485 us when every warp publishes and 80 slots are polled).
constexpr int kFpsClusterPpt = 4;

template <typename IdxT, bool PN2>
__global__ void __launch_bounds__(512, 1)
    fps_cluster_kernel(const float* __restrict__ xyz, int N, int G, int chunk, const long long* __restrict__ start_idx,
                       int skip_small, Pn2Order ord, IdxT* __restrict__ out_idx, float* __restrict__ out_centers) {
  constexpr int PPT = kFpsClusterPpt;
  extern __shared__ __align__(16) float s_dyn[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)gridDim.x, rank = (int)blockIdx.x, b = blockIdx.y;
  const int slots = PN2 ? (1 << (ord.lg + ord.lq)) : N;
  float* s_xyz = s_dyn;                                          // [3 * slots] the whole cloud (PN2: in tie order)
  int* s_sel = reinterpret_cast<int*>(s_dyn + 3 * slots);        // [G]
  __shared__ __align__(16) unsigned long long s_mine[2];         // [parity] this CTA's candidate, read by every peer
  __shared__ uint32_t s_cur[2];                                  // [parity] the cluster-wide winner, for the other warps
  __shared__ uint2 s_red[2][32];

  const int T = blockDim.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const float* cloud = xyz + (size_t)b * N * 3;
  const int n0 = rank * chunk, nl = max(0, min(N, n0 + chunk) - n0);

  if (tid < 2) s_mine[tid] = 0ull;                               // tag 0 = nothing published yet
  if (PN2) {
    for (int i = tid; i < 3 * N; i += T) {
      const int k = i / 3;
      s_xyz[3 * ord.comp((uint32_t)k) + (i - 3 * k)] = __ldg(cloud + i);
    }
  } else {
    for (int i = tid; i < 3 * N; i += T) s_xyz[i] = __ldg(cloud + i);
  }
  __syncthreads();

  float px[PPT], py[PPT], pz[PPT], dmin[PPT];
  uint32_t cj[PN2 ? PPT : 1];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int p = j * T + tid;
    bool valid = p < nl;
    const int pos = PN2 ? (int)ord.comp((uint32_t)(n0 + (valid ? p : 0))) : n0 + p;
    if (PN2) cj[j] = (uint32_t)pos;
    px[j] = valid ? s_xyz[3 * pos + 0] : 0.f;
    py[j] = valid ? s_xyz[3 * pos + 1] : 0.f;
    pz[j] = valid ? s_xyz[3 * pos + 2] : 0.f;
    if (PN2) {
      if (valid) valid = !pn2_skipped(px[j], py[j], pz[j]);
      dmin[j] = valid ? kFpsInit : -1.f;
    } else {
      if (skip_small && valid) valid = sqnorm_nofma(px[j], py[j], pz[j]) > 1e-3f;
      dmin[j] = valid ? kFpsInit : 0.f;
    }
  }

  long long s0 = start_idx ? start_idx[b] : 0;
  if (s0 < 0) s0 = 0;
  if (s0 >= N) s0 = N - 1;
  uint32_t cur = PN2 ? ord.comp((uint32_t)s0) : (uint32_t)s0;     // PN2: a tie-order word = shared-memory position
  uint32_t peer_slot = 0;          // lane r < C: cluster address of CTA r's slot (parity 0)
  if (lane < C)
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_slot) : "r"(smem_u32(&s_mine[0])), "r"(lane));
  cluster.sync();                  // every CTA has zeroed its slots before anybody polls them

  for (int i = 0; i < G; ++i) {
    if (tid == 0) s_sel[i] = (int)cur;
    if (i == G - 1) break;
    const float cx = s_xyz[3 * cur + 0], cy = s_xyz[3 * cur + 1], cz = s_xyz[3 * cur + 2];
    float best = -1.f;
    uint32_t bits, cand;           // this thread's candidate: key (distance bits) and tie-order word
    if (PN2) {
      uint32_t bestc = 0u;
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const float d = sqdist_pn2(px[j], py[j], pz[j], cx, cy, cz);
        const float dm = fminf(dmin[j], d);
        dmin[j] = dm;
        const bool gt = dm > best, eq = dm == best;      // maximum + lowest tie-order word among the points that share it
        bestc = gt ? cj[j] : (eq ? min(bestc, cj[j]) : bestc);
        best = gt ? dm : best;
      }
      if (best < 0.f) bestc = 0u;
      bits = pn2_key(best), cand = bestc;
    } else {
      int bestj = 0;
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const float d = sqdist_nofma(px[j], py[j], pz[j], cx, cy, cz);
        const float dm = fminf(dmin[j], d);
        dmin[j] = dm;
        if (dm > best) {
          best = dm;
          bestj = j;
        }
      }
      bits = __float_as_uint(best), cand = (uint32_t)(bestj * T + tid);
    }
    // slice-local argmax (lowest tie-order word among the maxima); out-of-range slots of the torch-order variant hold
    // distance 0 and the highest local indices
    const int par = i & 1;
    const uint32_t wmax = __reduce_max_sync(kFullMask, bits);
    const uint32_t widx = __reduce_min_sync(kFullMask, bits == wmax ? cand : 0xffffffffu);
    if (lane == 0) s_red[par][warp] = make_uint2(wmax, widx);
    __syncthreads();
    if (warp == 0) {
      // publish into THIS CTA's slot: tag | key (low word), tie-order word (high word); empty slice: key 0
      const uint32_t tag = (((uint32_t)i >> 1) & 1u) ^ 1u;       // flips on every reuse of parity slot `par`
      const uint2 v = lane < nwarps ? s_red[par][lane] : make_uint2(0u, 0xffffffffu);
      const uint32_t bmax = __reduce_max_sync(kFullMask, v.x);
      const uint32_t lidx = __reduce_min_sync(kFullMask, v.x == bmax ? v.y : 0xffffffffu);
      if (lane == 0) {
        uint32_t lo, hi;
        if (PN2) {
          lo = bmax | (tag << 31), hi = lidx;                    // key 0 <-> comp 0 (point 0), as upstream
        } else {
          const bool has = (int)lidx < nl;
          lo = (has ? bmax : 0u) | (tag << 31);
          hi = has ? (uint32_t)n0 + lidx : 0xffffffffu;
        }
        const unsigned long long word = ((unsigned long long)hi << 32) | lo;
        asm volatile("st.volatile.shared.b64 [%0], %1;" ::"r"(smem_u32(&s_mine[par])), "l"(word) : "memory");
      }
      // collect (pull): lane r < C polls CTA r's slot through distributed shared memory until it carries this
      // iteration's tag; warp maximum of the keys, lowest tie-order word among the maxima.
      // Only this warp polls: more pollers only slow the peers' shared-memory ports down (measured).
      uint32_t cb = 0, hi = 0xffffffffu;
      if (lane < C) {
        unsigned long long word;
        do {
          asm volatile("ld.volatile.shared::cluster.b64 %0, [%1];" : "=l"(word) : "r"(peer_slot + 8u * (uint32_t)par) : "memory");
        } while ((uint32_t)(((uint32_t)word) >> 31) != tag);
        cb = (uint32_t)word & 0x7fffffffu;
        hi = (uint32_t)(word >> 32);
      }
      __syncwarp();
      const uint32_t gmax = __reduce_max_sync(kFullMask, cb);
      const uint32_t g = __reduce_min_sync(kFullMask, (lane < C && cb == gmax) ? hi : 0xffffffffu);
      if (lane == 0) s_cur[par] = g;
    }
    __syncthreads();
    const uint32_t gidx = s_cur[par];
    cur = PN2 ? gidx : (gidx < (uint32_t)N ? gidx : 0u);     // (all slices empty cannot happen: N >= 1)
  }
  if (rank == 0) write_selection<IdxT, PN2>(s_sel, s_xyz, nullptr, b, G, out_idx, out_centers, ord);   // every CTA holds the same list
  cluster.sync();
}

// Large-cloud path (N > 16384): distances in a caller-provided [B,N] scratch, coordinates re-read through L2.
template <typename IdxT>
__global__ void __launch_bounds__(1024, 1)
    fps_gmem_kernel(const float* __restrict__ xyz, int N, int G, const long long* __restrict__ start_idx,
                    int skip_small, IdxT* __restrict__ out_idx, float* __restrict__ out_centers,
                    float* __restrict__ scratch) {
  extern __shared__ __align__(16) float s_dyn[];
  int* s_sel = reinterpret_cast<int*>(s_dyn);  // [G]
  __shared__ uint2 s_red[2][32];
  const int b = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const float* cloud = xyz + (size_t)b * N * 3;
  float* dist = scratch + (size_t)b * N;
  for (int p = tid; p < N; p += T) {
    bool valid = true;
    if (skip_small) valid = sqnorm_nofma(cloud[3 * p], cloud[3 * p + 1], cloud[3 * p + 2]) > 1e-3f;
    dist[p] = valid ? kFpsInit : 0.f;
  }
  long long s0 = start_idx ? start_idx[b] : 0;
  if (s0 < 0) s0 = 0;
  if (s0 >= N) s0 = N - 1;
  uint32_t cur = (uint32_t)s0;
  for (int i = 0; i < G; ++i) {
    if (tid == 0) s_sel[i] = (int)cur;
    if (i == G - 1) break;
    const float cx = __ldg(cloud + 3 * (size_t)cur), cy = __ldg(cloud + 3 * (size_t)cur + 1),
                cz = __ldg(cloud + 3 * (size_t)cur + 2);
    float best = -1.f;
    uint32_t besti = 0;
    for (int p = tid; p < N; p += T) {  // each thread owns a fixed set of p: no cross-thread hazard on dist[]
      const float d = sqdist_nofma(__ldg(cloud + 3 * (size_t)p), __ldg(cloud + 3 * (size_t)p + 1),
                                   __ldg(cloud + 3 * (size_t)p + 2), cx, cy, cz);
      const float dm = fminf(dist[p], d);
      dist[p] = dm;
      if (dm > best) {
        best = dm;
        besti = (uint32_t)p;
      }
    }
    block_argmax(__float_as_uint(best), besti, s_red, i & 1, lane, warp, nwarps, cur);
  }
  write_selection<IdxT>(s_sel, nullptr, cloud, b, G, out_idx, out_centers);
}

template <int PPT, typename IdxT, bool PN2>
int launch_reg(const float* xyz, int B, int N, int G, const int64_t* start_idx, int skip_small, Pn2Order ord,
               void* out_idx, float* out_centers, int threads, cudaStream_t st) {
  const size_t slots = PN2 ? ((size_t)1 << (ord.lg + ord.lq)) : (size_t)N;     // tie-order positions are padded to a power of two
  const size_t smem = (size_t)3 * slots * sizeof(float) + (size_t)G * sizeof(int);
  UA_UNSUPPORTED(smem > 226 * 1024, "ua_fps_f32: N=%d does not fit in shared memory in the pointnet2 layout", N);
  auto kern = fps_reg_kernel<PPT, IdxT, PN2>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("ua_fps_f32: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return UA_ERR_CUDA;
    }
  }
  kern<<<B, threads, smem, st>>>(xyz, N, G, (const long long*)start_idx, skip_small, ord, (IdxT*)out_idx, out_centers);
  return check_launch("ua_fps_f32");
}

template <typename IdxT, bool PN2>
int dispatch(const float* xyz, int B, int N, int G, const int64_t* start_idx, int skip_small, void* out_idx,
             float* out_centers, float* scratch, cudaStream_t st) {
  Pn2Order ord;
  ord.lg = 0;
  while ((2 << ord.lg) <= N && (2 << ord.lg) <= 512) ++ord.lg;     // upstream opt_n_threads(N)
  const int rounds = (N + (1 << ord.lg) - 1) >> ord.lg;
  ord.lq = 0;
  while ((1 << ord.lq) < rounds) ++ord.lq;
  if (N > UA_FPS_MAX_REG_POINTS) {
    UA_UNSUPPORTED(PN2, "ua_fps_f32: the pointnet2_ops arithmetic is implemented for N <= %d", UA_FPS_MAX_REG_POINTS);
    UA_REQUIRE(scratch != nullptr, "ua_fps_f32: N=%d > %d needs a [B,N] f32 scratch", N, UA_FPS_MAX_REG_POINTS);
    UA_UNSUPPORTED(G > 40000, "ua_fps_f32: G=%d too large for the large-cloud path", G);
    const size_t smem = (size_t)G * sizeof(int);
    auto kern = fps_gmem_kernel<IdxT>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<B, 1024, smem, st>>>(xyz, N, G, (const long long*)start_idx, skip_small, (IdxT*)out_idx, out_centers,
                                scratch);
    return check_launch("ua_fps_f32(gmem)");
  }
  // few large clouds: a cluster of 8 CTAs per cloud (distributed shared memory exchange per iteration)
  {
    const int min_n = g_fps_cluster > 0 ? g_fps_cluster : 2049;
    const int C = 8;
    if (g_fps_cluster >= 0 && N >= min_n && (long long)B * C <= 2 * kNumSMs && B <= 65535) {
      const int chunk = (N + C - 1) / C;
      const int threads = (((chunk + kFpsClusterPpt - 1) / kFpsClusterPpt) + 31) / 32 * 32;
      if (threads <= 512) {
        const size_t slots = PN2 ? ((size_t)1 << (ord.lg + ord.lq)) : (size_t)N;
        const size_t smem = (size_t)3 * slots * sizeof(float) + (size_t)G * sizeof(int);
        auto kern = fps_cluster_kernel<IdxT, PN2>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C, B, 1);
        cfg.blockDim = dim3(threads, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xyz, N, G, chunk, (const long long*)start_idx, skip_small, ord,
                                           (IdxT*)out_idx, out_centers);
        if (e != cudaSuccess) {
          set_error("ua_fps_f32(cluster): launch failed: %s", cudaGetErrorString(e));
          return UA_ERR_CUDA;
        }
        return check_launch("ua_fps_f32(cluster)");
      }
    }
  }
  // threads per cloud: the loop is latency-bound for small clouds (few warps -> cheaper barrier) and
  // issue-bound for large ones (N*12 instructions per iteration on one SM).
  int target = g_fps_threads > 0 ? g_fps_threads : (N <= 2048 ? 256 : 1024);
  static const int kPpt[] = {1, 2, 4, 8, 12, 16};
  int ppt = 16;
  for (int cand : kPpt) {
    const int need = (N + cand - 1) / cand;
    if (need <= target && need <= (cand == 12 ? 896 : 1024)) {
      ppt = cand;
      break;
    }
  }
  int threads = (((N + ppt - 1) / ppt) + 31) / 32 * 32;
  switch (ppt) {
    case 1: return launch_reg<1, IdxT, PN2>(xyz, B, N, G, start_idx, skip_small, ord, out_idx, out_centers, threads, st);
    case 2: return launch_reg<2, IdxT, PN2>(xyz, B, N, G, start_idx, skip_small, ord, out_idx, out_centers, threads, st);
    case 4: return launch_reg<4, IdxT, PN2>(xyz, B, N, G, start_idx, skip_small, ord, out_idx, out_centers, threads, st);
    case 8: return launch_reg<8, IdxT, PN2>(xyz, B, N, G, start_idx, skip_small, ord, out_idx, out_centers, threads, st);
    case 12: return launch_reg<12, IdxT, PN2>(xyz, B, N, G, start_idx, skip_small, ord, out_idx, out_centers, threads, st);
    default: return launch_reg<16, IdxT, PN2>(xyz, B, N, G, start_idx, skip_small, ord, out_idx, out_centers, threads, st);
  }
}

}  // namespace
}  // namespace ua

extern "C" int ua_fps_f32(const float* xyz, int B, int N, int G, const int64_t* start_idx, int skip_small_norm,
                          void* out_idx, int idx_is_i64, float* out_centers, float* scratch, void* stream) {
  using namespace ua;
  UA_REQUIRE(xyz != nullptr, "ua_fps_f32: xyz is NULL");
  UA_REQUIRE(B >= 0 && N >= 1 && G >= 1, "ua_fps_f32: bad sizes B=%d N=%d G=%d", B, N, G);
  if (B == 0) return UA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (skip_small_norm & UA_FPS_POINTNET2)
    return idx_is_i64 ? dispatch<long long, true>(xyz, B, N, G, start_idx, 0, out_idx, out_centers, scratch, st)
                      : dispatch<int, true>(xyz, B, N, G, start_idx, 0, out_idx, out_centers, scratch, st);
  return idx_is_i64 ? dispatch<long long, false>(xyz, B, N, G, start_idx, skip_small_norm & 1, out_idx, out_centers, scratch, st)
                    : dispatch<int, false>(xyz, B, N, G, start_idx, skip_small_norm & 1, out_idx, out_centers, scratch, st);
}
