// Grouping kernels of the point tokenizer for sm_100a.
//
//   knn_group_kernel : k nearest points of every FPS centre under the reference's expanded squared distance,
//                      fused with gather, centre subtraction and channel concat
//                      (models/point_encoder.py:17-49,99-127, models/ulip/pointbert/dvae.py:116-181).
//   ball_group_kernel: first-nsample-in-index-order ball query, fused with the same epilogue
//                      (models/openshape/pointnet_util.py:89-146).
//
// One warp owns one centre. A CTA's warps share point tiles that the TMA unit double-buffers in shared memory
// (coordinates AoS: stride-3 reads are bank-conflict free; plus the per-point squared norm). The (G,N) distance
// matrix of the reference is never written. kNN selection is a streaming filter: candidates below the running
// threshold are ballot-compacted into a per-warp shared-memory buffer; when it fills, a register-resident radix
// select (bitwise binary search, one REDUX per bit) keeps the k best and tightens the threshold -- O(log(N/k))
// selections per centre, no sort anywhere: the buffer stays in ascending point-index order, which is also the
// order the neighbours are emitted in.
#include "common.cuh"

namespace ua {

int g_knn_warps = 0;  // tuning override (0 = heuristic)
int g_knn_hist = 0;   // tuning: -1 = streaming filter only, 2 = candidate-buffer histogram selection only (no register masks)

namespace {

constexpr int kTilePoints = 1024;
constexpr uint64_t kKeyMax = ~0ull;

// ---- point-tile pipeline shared by both grouping kernels ---------------------------------------------------
// Tiles of up to kTilePoints points (AoS xyz, 12 KB) are double-buffered in shared memory. When the cloud is
// 16-byte tileable (N % 4 == 0, aligned base) the TMA unit fills the next tile with one bulk copy while the
// warps scan the current one; otherwise the CTA copies cooperatively.
struct TileSmem {
  float xyz[2][3 * kTilePoints];
  float pn[2][kTilePoints];
  uint64_t bar[2];
};

struct TilePipe {
  TileSmem* sm;
  const float* cloud;
  int N, ntiles, bulk;
  uint32_t phase;

  __device__ __forceinline__ void init(TileSmem* s, const float* c, int n, int use_bulk) {
    sm = s, cloud = c, N = n, bulk = use_bulk, phase = 0;
    ntiles = (n + kTilePoints - 1) / kTilePoints;
    if (bulk && threadIdx.x == 0) {
      mbar_init(&sm->bar[0], 1);
      mbar_init(&sm->bar[1], 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (bulk && threadIdx.x == 0) issue(0);
  }
  __device__ __forceinline__ int tile_points(int t) const { return min(kTilePoints, N - t * kTilePoints); }
  __device__ __forceinline__ void issue(int t) {
    const int st = t & 1;
    const uint32_t bytes = (uint32_t)tile_points(t) * 12u;
    mbar_expect_tx(&sm->bar[st], bytes);
    bulk_g2s(sm->xyz[st], cloud + (size_t)3 * t * kTilePoints, bytes, &sm->bar[st]);
  }
  // Makes tile t resident (coordinates + squared norms) for every thread of the CTA; returns its stage.
  // soa: also lay the coordinates out as three planes X | Y | Z in the OTHER stage's buffer (single-tile clouds only:
  // that stage is never filled), which the register-mask kNN selection reads with 16-byte loads.
  __device__ __forceinline__ int acquire(int t, bool soa = false) {
    const int st = t & 1;
    const int tp = tile_points(t);
    if (bulk) {
      if (threadIdx.x == 0 && t + 1 < ntiles) issue(t + 1);  // stage st^1 was released by the barrier in release()
      mbar_wait(&sm->bar[st], (phase >> st) & 1u);
      phase ^= 1u << st;
    } else {
      const float* src = cloud + (size_t)3 * t * kTilePoints;
      for (int i = threadIdx.x; i < 3 * tp; i += blockDim.x) sm->xyz[st][i] = __ldg(src + i);
      __syncthreads();
    }
    if (soa) {
      float* pl = sm->xyz[st ^ 1];
      for (int p = threadIdx.x; p < tp; p += blockDim.x) {
        const float x = sm->xyz[st][3 * p], y = sm->xyz[st][3 * p + 1], z = sm->xyz[st][3 * p + 2];
        pl[p] = x, pl[kTilePoints + p] = y, pl[2 * kTilePoints + p] = z;
        sm->pn[st][p] = sqnorm_nofma(x, y, z);
      }
    } else {
      for (int p = threadIdx.x; p < tp; p += blockDim.x)
        sm->pn[st][p] = sqnorm_nofma(sm->xyz[st][3 * p], sm->xyz[st][3 * p + 1], sm->xyz[st][3 * p + 2]);
    }
    __syncthreads();
    return st;
  }
  __device__ __forceinline__ void release() { __syncthreads(); }
};

// ---- kNN selection -------------------------------------------------------------------------------------------
// Per warp: a shared-memory buffer of CAP (distance-key, index) pairs, always in ascending point-index order
// (appends follow the scan order and compaction is a stable filter). When the buffer cannot take another
// batch, the warp finds the k-th smallest distance key by a bitwise binary search over register-held keys
// (one REDUX add per bit), keeps the keys below it plus the lowest-index ties, and tightens the threshold.
template <int CAP>
__device__ __forceinline__ void knn_compact(uint32_t* bkey, uint32_t* bidx, int& count, int k, int lane,
                                            uint64_t& thr) {
  constexpr int R = CAP / 32;
  __syncwarp();
  uint32_t key[R], idx[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    key[r] = e < count ? bkey[e] : 0xffffffffu;
    idx[r] = e < count ? bidx[e] : 0xffffffffu;
  }
  // k-th smallest key: the largest T with count(key < T) < k
  uint32_t T = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) c += key[r] < cand;
    c = __reduce_add_sync(kFullMask, c);
    if (c < k) T = cand;
  }
  int nless = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) nless += key[r] < T;
  nless = __reduce_add_sync(kFullMask, nless);
  const int need_eq = k - nless;  // >= 1 by construction of T
  const unsigned lt_mask = (1u << lane) - 1u;
  __syncwarp();
  int eq_seen = 0, out = 0;
  uint32_t last_eq_idx = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool eq = key[r] == T;
    const unsigned meq = __ballot_sync(kFullMask, eq);
    const int eq_rank = eq_seen + __popc(meq & lt_mask);
    const bool keep = key[r] < T || (eq && eq_rank < need_eq);
    // the tie that fills the k-th slot carries the threshold index
    const unsigned mlast = __ballot_sync(kFullMask, eq && eq_rank == need_eq - 1);
    if (mlast) last_eq_idx = __shfl_sync(kFullMask, idx[r], __ffs(mlast) - 1);
    eq_seen += __popc(meq);
    const unsigned mk = __ballot_sync(kFullMask, keep);
    if (keep) {
      const int pos = out + __popc(mk & lt_mask);
      bkey[pos] = key[r];
      bidx[pos] = idx[r];
    }
    out += __popc(mk);
  }
  __syncwarp();
  count = out;  // == k
  thr = ((uint64_t)T << 32) | last_eq_idx;
}


// ---- first tile by histogram selection -----------------------------------------------------------------------
// The streaming filter starts blind: until the first compaction every point is a candidate, and each compaction is a
// 32-step radix select over CAP register-held keys. For the first tile (<= 1024 points = 32 distances per lane, kept
// in registers) the k-th smallest is bracketed with one shared-memory histogram instead:
//   bin(d) = min(255, (bits(max(d,0) + delta) - bits(delta)) >> 18),  delta = max_d / 64
// is monotone in d (linear below delta, 32 bins per octave above: fine where the neighbours are, coarse in the far
// tail); a warp scan over the 256 counters yields the bin b* that holds the k-th smallest. Every point of a bin <= b*
// is emitted to the candidate buffer in ascending index order (k + a handful), and the surplus -- the largest
// (distance bits, index) pairs, all of them in bin b* -- is dropped by repeated warp-wide maxima (two REDUX each).
// Exactness does not depend on the bin shape: bins only pre-partition, the (distance bits, index) order decides.
// Returns false (nothing written) when the surplus is too large for that (e.g. many identical points): the caller then
// streams the tile as before.
constexpr int kMaxDrop = 12;

// Removes the (count - k) largest (key, index) pairs from the buffer (stable), R registers per lane cover `count`.
template <int R>
__device__ __forceinline__ void knn_drop_largest(uint32_t* bkey, uint32_t* bidx, int& count, int k, int lane,
                                                 bool want_thr, uint64_t& thr) {
  __syncwarp();
  uint32_t key[R], idx[R];
  bool alive[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    alive[r] = e < count;
    key[r] = alive[r] ? bkey[e] : 0u;
    idx[r] = alive[r] ? bidx[e] : 0u;
  }
  const int ndrop = count - k;
  for (int it = 0; it <= ndrop; ++it) {          // the last round only reads the k-th pair (the new threshold)
    if (it == ndrop && !want_thr) break;
    uint32_t bk = 0, bi = 0;
    bool any = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool better = alive[r] && (!any || key[r] > bk || (key[r] == bk && idx[r] > bi));
      if (better) bk = key[r], bi = idx[r], any = true;
    }
    const uint32_t mk = __reduce_max_sync(kFullMask, any ? bk : 0u);
    const uint32_t mi = __reduce_max_sync(kFullMask, (any && bk == mk) ? bi : 0u);
    if (it == ndrop) {
      thr = ((uint64_t)mk << 32) | mi;
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (alive[r] && key[r] == mk && idx[r] == mi) alive[r] = false;
    }
  }
  const unsigned lt_mask = (1u << lane) - 1u;
  __syncwarp();
  int out = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const unsigned m = __ballot_sync(kFullMask, alive[r]);
    if (alive[r]) {
      const int pos = out + __popc(m & lt_mask);
      bkey[pos] = key[r];
      bidx[pos] = idx[r];
    }
    out += __popc(m);
  }
  __syncwarp();
  count = out;  // == k
}

template <int CAP, bool FULL>
__device__ __forceinline__ bool knn_first_tile_hist(const float* __restrict__ sx, const float* __restrict__ sp, int tp,
                                                    float cx, float cy, float cz, float cn, int k, int lane,
                                                    uint32_t* bkey, uint32_t* bidx, int* hist, bool more_tiles,
                                                    int& count, uint64_t& thr) {
  constexpr int RD = CAP == 128 ? 2 : (CAP == 256 ? 4 : 5);   // registers per lane that cover k + kMaxDrop candidates
  const unsigned lt_mask = (1u << lane) - 1u;
  float d[32];
  float dm = 0.f;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const int p = r * 32 + lane;
    d[r] = INFINITY;
    if (FULL || p < tp) {
      d[r] = expanded_sqdist(cx, cy, cz, cn, sx[3 * p], sx[3 * p + 1], sx[3 * p + 2], sp[p]);
      dm = fmaxf(dm, d[r]);
    }
  }
  const float dmax = __uint_as_float(__reduce_max_sync(kFullMask, __float_as_uint(dm)));   // dm >= 0: bits are ordered
  const float delta = dmax * 0.015625f;
  const uint32_t dbits = __float_as_uint(delta);
#pragma unroll
  for (int j = 0; j < 8; ++j) hist[lane + 32 * j] = 0;
  __syncwarp();
  const uint32_t hist_s = smem_u32(hist);
#pragma unroll
  for (int r = 0; r < 32; ++r)
    if (FULL || r * 32 + lane < tp) {
      const uint32_t e = __float_as_uint(__fadd_rn(fmaxf(d[r], 0.f), delta)) - dbits;
      const uint32_t bin = min(e >> 18, 255u);
      // plain shared-memory reduction (the compiler would turn atomicAdd into a 30-instruction match-and-aggregate loop)
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist_s + 4u * bin) : "memory");
    }
  __syncwarp();
  // warp scan over the 256 counters (8 per lane)
  int h[8];
  {
    const int4 a = *reinterpret_cast<const int4*>(hist + 8 * lane), b = *reinterpret_cast<const int4*>(hist + 8 * lane + 4);
    h[0] = a.x, h[1] = a.y, h[2] = a.z, h[3] = a.w, h[4] = b.x, h[5] = b.y, h[6] = b.z, h[7] = b.w;
  }
  int ssum = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) ssum += h[j];
  int incl = ssum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFullMask, incl, o);
    if (lane >= o) incl += t;
  }
  const unsigned mb = __ballot_sync(kFullMask, incl >= k);      // tp >= k, so some lane crosses
  const int L = __ffs(mb) - 1;
  int bstar = 0, upto = 0;                                      // upto = points in bins <= b*
  if (lane == L) {
    int c = incl - ssum;
    bool found = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c += h[j];
      if (!found && c >= k) bstar = 8 * lane + j, upto = c, found = true;
    }
  }
  bstar = __shfl_sync(kFullMask, bstar, L);
  upto = __shfl_sync(kFullMask, upto, L);
  if (upto - k > kMaxDrop) return false;
  // every point of a bin <= b*, in index order, into the candidate buffer: bits(max(d,0) + delta) < ub
  const uint32_t ub = bstar >= 255 ? 0xffffffffu : dbits + ((uint32_t)(bstar + 1) << 18);
  int out = 0;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const int p = r * 32 + lane;
    const bool take = (FULL || p < tp) && __float_as_uint(__fadd_rn(fmaxf(d[r], 0.f), delta)) < ub;
    const unsigned m = __ballot_sync(kFullMask, take);
    if (take) {
      const int pos = out + __popc(m & lt_mask);
      bkey[pos] = float_to_ordered(d[r]);
      bidx[pos] = (uint32_t)p;
    }
    out += __popc(m);
  }
  count = out;  // == upto
  if (count > k || more_tiles) knn_drop_largest<RD>(bkey, bidx, count, k, lane, more_tiles, thr);
  __syncwarp();
  return true;
}


// ---- single-tile clouds: selection on register masks ---------------------------------------------------------------
// For clouds of one tile (N <= 1024: the ULIP / Uni3D 1024-point shapes) the candidate buffer disappears. A lane owns the
// points p = 128 q + 4 lane + j (q < 8, j < 4; bit b = 4 q + j of its masks), read from the coordinate planes with four
// 16-byte shared-memory loads per four points; the six operations of the reference's expanded distance run as packed
// f32x2 instructions (two points per instruction, each half rounded exactly like the scalar operation), and the 32
// distances stay in registers as 16 packed pairs.
//   1. 256-bin histogram over the top bits of the distance (32 bins per octave, the 8 octaves below the farthest point;
//      everything nearer falls into bin 0), warp scan -> the bin b* that holds the k-th smallest;
//   2. one pass builds two masks per lane: "below b*" (all selected) and "in b*" (at most 32 points in the warp, else the
//      caller falls back to the candidate-buffer path);
//   3. the points of b* go one per lane, are ranked by (distance bits, index) with one shuffle round per candidate, and
//      the first k - below of them set their bit in the owner's mask through shared memory;
//   4. the positions in ascending point-index order come from the masks alone: per-q counts packed as bytes, two warp
//      scans, and each lane stores the indices of its own set bits (k / 32 on average) -- no ballot per point anywhere.
// Bit-identical to the other paths: bins only pre-partition, the exact (distance bits, index) order decides.
__device__ __forceinline__ int float_to_skey(float f) {
  const int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}
constexpr int kBinShift = 18;   // 32 bins per octave of the squared distance

template <bool FULL>
__device__ __forceinline__ bool knn_single_tile_masks(const float* __restrict__ planes, const float* __restrict__ sp, int tp,
                                                      float cx, float cy, float cz, float cn, int k, int lane,
                                                      uint32_t* scratch, uint32_t* bidx, int* hist) {
  const float* X = planes;
  const float* Y = planes + kTilePoints;
  const float* Z = planes + 2 * kTilePoints;
  u64 dp[16];                   // the 32 distances as packed pairs: dp[2q] = points 4q, 4q+1; dp[2q+1] = points 4q+2, 4q+3
  int kmax = INT_MIN;           // maximum of the float bit patterns as signed ints (>= 0 unless every distance is negative)
  uint32_t vm = 0xffffffffu;    // this lane's points that exist (ragged tiles)
  {
    const u64 cx2 = pack2(cx, cx), cy2 = pack2(cy, cy), cz2 = pack2(cz, cz), cn2 = pack2(cn, cn), m2 = pack2(-2.0f, -2.0f);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int p0 = 128 * q + 4 * lane;
      const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(X + p0);
      const ulonglong2 y = *reinterpret_cast<const ulonglong2*>(Y + p0);
      const ulonglong2 z = *reinterpret_cast<const ulonglong2*>(Z + p0);
      const ulonglong2 n = *reinterpret_cast<const ulonglong2*>(sp + p0);
      // ((-2 * fma(cz,pz, fma(cy,py, cx*px))) + |c|^2) + |p|^2, two points per instruction. ptxas contracts a
      // mul.rn.f32x2 feeding an add.rn.f32x2 into one FFMA2 (tools/ubench/f32x2_contract.cu) -- here that pair is
      // (-2 * dot) + |c|^2, and a multiplication by -2 is exact, so the contraction cannot change a bit
      dp[2 * q] = add2(add2(mul2(m2, fma2(cz2, z.x, fma2(cy2, y.x, mul2(cx2, x.x)))), cn2), n.x);
      dp[2 * q + 1] = add2(add2(mul2(m2, fma2(cz2, z.y, fma2(cy2, y.y, mul2(cx2, x.y)))), cn2), n.y);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int bits = (int)(uint32_t)(dp[2 * q + (j >> 1)] >> (32 * (j & 1)));
        if (!FULL && p0 + j >= tp) vm &= ~(1u << (4 * q + j));
        else kmax = max(kmax, bits);
      }
    }
  }
  // The distances stay floats: for d >= 0 the bit pattern is monotone, and the (rare, tiny) negative values of the expanded
  // form all fall into bin 0 / below every threshold, so histogram and masks need no order-preserving integer image; only
  // the <= 32 candidates of the k-th neighbour's bin are ranked on exact (sign-folded bits, index) pairs further down.
  kmax = __reduce_max_sync(kFullMask, kmax);
  const int off = (max(kmax, 0) >> kBinShift) - 255;     // bin(d) = max((bits(d) >> 18) - off, 0) <= 255
  if (off < 0) return false;                             // farthest point nearer than 2^-119: leave it to the general path
#pragma unroll
  for (int j = 0; j < 8; ++j) hist[lane + 32 * j] = 0;
  __syncwarp();
  const uint32_t hist_s = smem_u32(hist);
#pragma unroll
  for (int b = 0; b < 32; ++b)
    if (FULL || ((vm >> b) & 1u)) {
      const int bits = (int)(uint32_t)(dp[b >> 1] >> (32 * (b & 1)));
      const int bin = max((bits >> kBinShift) - off, 0);
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist_s + 4u * (uint32_t)bin) : "memory");
    }
  __syncwarp();
  // warp scan over the 256 counters (8 per lane)
  int h[8];
  {
    const int4 a = *reinterpret_cast<const int4*>(hist + 8 * lane), b = *reinterpret_cast<const int4*>(hist + 8 * lane + 4);
    h[0] = a.x, h[1] = a.y, h[2] = a.z, h[3] = a.w, h[4] = b.x, h[5] = b.y, h[6] = b.z, h[7] = b.w;
  }
  int pre[8];                                                   // inclusive prefix of this lane's 8 bins (lane-local)
  pre[0] = h[0];
#pragma unroll
  for (int j = 1; j < 8; ++j) pre[j] = pre[j - 1] + h[j];
  int incl = pre[7];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFullMask, incl, o);
    if (lane >= o) incl += t;
  }
  const int excl = incl - pre[7];
  const unsigned mb = __ballot_sync(kFullMask, incl >= k);      // tp >= k, so some lane crosses
  const int L = __ffs(mb) - 1;
  // in lane L: the first bin whose inclusive prefix reaches k (branch-free: every lane evaluates its own 8 bins)
  int jstar = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) jstar += (excl + pre[j] < k) ? 1 : 0;
  int nb_l = h[0], below_l = excl;
#pragma unroll
  for (int j = 1; j < 8; ++j)
    if (jstar >= j) nb_l = h[j], below_l = excl + pre[j - 1];
  const int bstar = 8 * L + __shfl_sync(kFullMask, jstar, L);  // the bin of the k-th neighbour
  const int below = __shfl_sync(kFullMask, below_l, L);         // points in bins < b*
  const int nb = __shfl_sync(kFullMask, nb_l, L);               // points in b*
  if (nb > 32) return false;
  const int need = k - below;                                   // 1 .. nb
  // d-range of bin b*: [lo, hi) (bin 0 is open below). d < t  <=>  sign(d - t): one packed add per two points and one
  // funnel shift per point and mask (the first point inserted ends up in bit 31: pairs are walked from the top)
  const float lo_f = bstar == 0 ? -INFINITY : __int_as_float((bstar + off) << kBinShift);
  const int hib = (bstar + off + 1) << kBinShift;               // <= 0x7f800000
  const float hi_f = hib >= 0x7f800000 ? 3.402823466e38f : __int_as_float(hib);
  uint32_t m_below = 0, m_lt_hi = 0;
  {
    const u64 nlo2 = pack2(-lo_f, -lo_f), nhi2 = pack2(-hi_f, -hi_f);
#pragma unroll
    for (int i = 15; i >= 0; --i) {
      const u64 tl = add2(dp[i], nlo2), th = add2(dp[i], nhi2);
      m_below = __funnelshift_l((uint32_t)(tl >> 32), m_below, 1);
      m_below = __funnelshift_l((uint32_t)tl, m_below, 1);
      m_lt_hi = __funnelshift_l((uint32_t)(th >> 32), m_lt_hi, 1);
      m_lt_hi = __funnelshift_l((uint32_t)th, m_lt_hi, 1);
    }
    if (!FULL) m_below &= vm, m_lt_hi &= vm;
  }
  const uint32_t m_in = m_lt_hi & ~m_below;
  // the points of b*, one per lane: slots from a warp scan of the per-lane counts
  uint32_t* cand = scratch;          // [32] point indices
  uint32_t* selw = scratch + 32;     // [32] bits the ranked candidates add to their owner's mask
  {
    const int cnt = __popc(m_in);
    int pos = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFullMask, pos, o);
      if (lane >= o) pos += t;
    }
    pos -= cnt;
    selw[lane] = 0u;
    for (uint32_t mm = m_in; mm; mm &= mm - 1) {
      const int b = __ffs(mm) - 1;
      cand[pos++] = (uint32_t)(128 * (b >> 2) + 4 * lane + (b & 3));
    }
  }
  __syncwarp();
  {
    int ck = INT_MAX, ci = INT_MAX;
    if (lane < nb) {
      ci = (int)cand[lane];
      ck = float_to_skey(expanded_sqdist(cx, cy, cz, cn, X[ci], Y[ci], Z[ci], sp[ci]));
    }
    int rank = 0;
    for (int i = 0; i < nb; ++i) {
      const int ok = __shfl_sync(kFullMask, ck, i), oi = __shfl_sync(kFullMask, ci, i);
      rank += (ok < ck || (ok == ck && oi < ci)) ? 1 : 0;
    }
    if (lane < nb && rank < need) atomicOr(&selw[(ci >> 2) & 31], 1u << (((ci >> 7) << 2) | (ci & 3)));
  }
  __syncwarp();
  const uint32_t m = m_below | selw[lane];                      // this lane's selected points; k bits in the warp
  // positions in ascending point order (q, lane, j): per-q counts as packed bytes, scanned over the lanes
  uint32_t nib = m - ((m >> 1) & 0x55555555u);
  nib = (nib & 0x33333333u) + ((nib >> 2) & 0x33333333u);       // nibble q = number of selected points of row q
  const uint32_t c_even = nib & 0x0f0f0f0fu, c_odd = (nib >> 4) & 0x0f0f0f0fu;   // byte a: q = 2a / 2a + 1
  uint32_t s_even = c_even, s_odd = c_odd;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t te = __shfl_up_sync(kFullMask, s_even, o), to = __shfl_up_sync(kFullMask, s_odd, o);
    if (lane >= o) s_even += te, s_odd += to;
  }
  const uint32_t t_even = __shfl_sync(kFullMask, s_even, 31), t_odd = __shfl_sync(kFullMask, s_odd, 31);
  {
    // sanity: exactly k selected (every byte sum stays below 256 because the total is k <= 128)
    const uint32_t tot = t_even + t_odd;
    const int total = (int)((tot & 0xff) + ((tot >> 8) & 0xff) + ((tot >> 16) & 0xff) + (tot >> 24));
    if (total != k) return false;
  }
  const uint32_t pair = t_even + t_odd;
  const uint32_t E = (pair << 8) + (pair << 16) + (pair << 24);         // byte a: selected points of rows < 2a
  const uint32_t w_even = E + (s_even - c_even);                        // byte a: first position of this lane in row 2a
  const uint32_t w_odd = E + t_even + (s_odd - c_odd);                  //                                   row 2a + 1
  {
    int prevq = -1, run = 0;                                            // selected points of this lane seen in the current row
    const uint32_t lane4 = 4u * (uint32_t)lane;
    for (uint32_t mm = m; mm; mm &= mm - 1) {
      const int b = __ffs(mm) - 1, q = b >> 2;
      run = q == prevq ? run + 1 : 0;
      prevq = q;
      const uint32_t w = (b & 4) ? w_odd : w_even;
      const int pos = (int)__byte_perm(w, 0u, 0x4440u + (uint32_t)(b >> 3)) + run;   // byte q/2 of w + rank inside the row
      bidx[pos] = ((uint32_t)q << 7) + lane4 + (uint32_t)(b & 3);
    }
  }
  __syncwarp();
  return true;
}

template <int CAP, typename IdxT>
__global__ void __launch_bounds__(256, 3)   // 80 registers: 3 CTAs per SM (4 = 64 registers spills and is slower: 52.6 vs 43.8 us)
    knn_group_kernel(const float* __restrict__ xyz, const float* __restrict__ rgb, const float* __restrict__ centers,
                     int N, int G, int k, int use_bulk, int use_hist, IdxT* __restrict__ out_idx,
                     float* __restrict__ out_neigh, float* __restrict__ out_feat) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  TileSmem* tiles = reinterpret_cast<TileSmem*>(s_raw);
  uint32_t* s_key = reinterpret_cast<uint32_t*>(s_raw + sizeof(TileSmem));  // [W][CAP]
  const int W = blockDim.x >> 5;
  uint32_t* s_idx = s_key + (size_t)W * CAP;                                // [W][CAP]
  int* s_hist = reinterpret_cast<int*>(s_idx + (size_t)W * CAP);            // [W][256]

  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.x * W + warp;
  const bool active = g < G;
  const float* cloud = xyz + (size_t)b * N * 3;
  uint32_t* bkey = s_key + (size_t)warp * CAP;
  uint32_t* bidx = s_idx + (size_t)warp * CAP;

  int count = 0;
  uint64_t thr = kKeyMax;
  const unsigned lt_mask = (1u << lane) - 1u;

  TilePipe pipe;
  pipe.init(tiles, cloud, N, use_bulk);
  // (the centre is fetched AFTER the first tile has been requested: the two global-memory round trips overlap)
  float cx = 0.f, cy = 0.f, cz = 0.f, cn = 0.f;
  if (active) {
    const float* c = centers + ((size_t)b * G + g) * 3;
    cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
    cn = sqnorm_nofma(cx, cy, cz);
  }
  const bool masks = use_hist == 1 && pipe.ntiles == 1;   // register-mask selection (single-tile clouds)
  bool from_planes = false;
  for (int t = 0; t < pipe.ntiles; ++t) {
    const int st = pipe.acquire(t, masks);
    if (active) {
      const float* sx = tiles->xyz[st];
      const float* sp = tiles->pn[st];
      const int tp = pipe.tile_points(t), t0 = t * kTilePoints;
      bool done = false;
      if (masks) {
        int* hist = s_hist + warp * 256;
        done = tp == kTilePoints
                   ? knn_single_tile_masks<true>(tiles->xyz[st ^ 1], sp, tp, cx, cy, cz, cn, k, lane, bkey, bidx, hist)
                   : knn_single_tile_masks<false>(tiles->xyz[st ^ 1], sp, tp, cx, cy, cz, cn, k, lane, bkey, bidx, hist);
        if (done) count = k;
        from_planes = done;
      }
      if (!done && t == 0 && use_hist) {
        int* hist = s_hist + warp * 256;
        const bool more = pipe.ntiles > 1;
        done = tp == kTilePoints
                   ? knn_first_tile_hist<CAP, true>(sx, sp, tp, cx, cy, cz, cn, k, lane, bkey, bidx, hist, more, count, thr)
                   : knn_first_tile_hist<CAP, false>(sx, sp, tp, cx, cy, cz, cn, k, lane, bkey, bidx, hist, more, count, thr);
      }
      for (int base = 0; base < tp && !done; base += 32) {
        const int p = base + lane;
        uint64_t key = kKeyMax;
        if (p < tp) {
          const float d = expanded_sqdist(cx, cy, cz, cn, sx[3 * p], sx[3 * p + 1], sx[3 * p + 2], sp[p]);
          key = ((uint64_t)float_to_ordered(d) << 32) | (uint32_t)(t0 + p);
        }
        const bool take = key < thr;  // kKeyMax is never < thr
        const unsigned m = __ballot_sync(kFullMask, take);
        if (m) {
          if (take) {
            const int pos = count + __popc(m & lt_mask);
            bkey[pos] = (uint32_t)(key >> 32);
            bidx[pos] = (uint32_t)key;
          }
          count += __popc(m);
          if (count > CAP - 32) knn_compact<CAP>(bkey, bidx, count, k, lane, thr);
        }
      }
    }
    pipe.release();
  }
  if (!active) return;
  if (count > k) knn_compact<CAP>(bkey, bidx, count, k, lane, thr);
  __syncwarp();

  // epilogue: ascending point index, gather, centre subtraction, concat
  const size_t row0 = ((size_t)b * G + g) * k;
  for (int j = lane; j < k; j += 32) {
    const uint32_t p = bidx[j];
    if (out_idx) out_idx[row0 + j] = (IdxT)p;
    float px, py, pz;
    if (from_planes) {   // the tile is still resident: coordinates from shared memory
      const float* pl = tiles->xyz[1];
      px = pl[p], py = pl[kTilePoints + p], pz = pl[2 * kTilePoints + p];
    } else {
      const float* src = cloud + (size_t)3 * p;
      px = __ldg(src), py = __ldg(src + 1), pz = __ldg(src + 2);
    }
    const float nx = __fsub_rn(px, cx), ny = __fsub_rn(py, cy), nz = __fsub_rn(pz, cz);
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
                      const float* __restrict__ centers, int N, int S, float radius2, int nsample, int use_bulk,
                      IdxT* __restrict__ out_idx, float* __restrict__ out_new_points) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  TileSmem* tiles = reinterpret_cast<TileSmem*>(s_raw);
  int* s_sel = reinterpret_cast<int*>(s_raw + sizeof(TileSmem));  // [W][nsample]
  const int W = blockDim.x >> 5;

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

  TilePipe pipe;
  pipe.init(tiles, cloud, N, use_bulk);
  for (int t = 0; t < pipe.ntiles; ++t) {
    // every warp of the CTA takes part in the tile hand-over even when its own ball is already full
    const int st = pipe.acquire(t);
    if (active && count < nsample) {
      const float* sx = tiles->xyz[st];
      const float* sp = tiles->pn[st];
      const int tp = pipe.tile_points(t), t0 = t * kTilePoints;
      for (int base = 0; base < tp && count < nsample; base += 32) {
        const int p = base + lane;
        bool inside = false;
        if (p < tp) {
          const float d = expanded_sqdist(cx, cy, cz, cn, sx[3 * p], sx[3 * p + 1], sx[3 * p + 2], sp[p]);
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
    pipe.release();
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
  if (centres >= 8LL * 2 * kNumSMs) return 8;
  return 4;
}

template <int CAP, typename IdxT>
int launch_knn(const float* xyz, const float* rgb, const float* centers, int B, int N, int G, int k, void* out_idx,
               float* out_neigh, float* out_feat, cudaStream_t st) {
  const int W = pick_warps(B, G);
  const size_t smem = sizeof(TileSmem) + (size_t)W * CAP * 8 + (size_t)W * 256 * sizeof(int);
  const int use_bulk = (N % 4 == 0) && ((uintptr_t)xyz % 16 == 0);
  auto kern = knn_group_kernel<CAP, IdxT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("ua_knn_group_f32: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return UA_ERR_CUDA;
    }
  }
  dim3 grid((G + W - 1) / W, B);
  kern<<<grid, W * 32, smem, st>>>(xyz, rgb, centers, N, G, k, use_bulk, g_knn_hist < 0 ? 0 : (g_knn_hist == 2 ? 2 : 1), (IdxT*)out_idx, out_neigh,
                                   out_feat);
  return check_launch("ua_knn_group_f32");
}

template <typename IdxT>
int dispatch_knn(const float* xyz, const float* rgb, const float* centers, int B, int N, int G, int k, void* out_idx,
                 float* out_neigh, float* out_feat, cudaStream_t st) {
  // buffer capacity: room for at least k fresh candidates (plus one 32-wide batch) between compactions
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
  const size_t smem = sizeof(TileSmem) + (size_t)W * nsample * 4;
  const int use_bulk = (N % 4 == 0) && ((uintptr_t)xyz % 16 == 0);
  dim3 grid((S + W - 1) / W, B);
  if (idx_is_i64) {
    auto kern = ball_group_kernel<long long>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, W * 32, smem, st>>>(xyz, feat, C, centers, N, S, radius2, nsample, use_bulk, (long long*)out_idx,
                                     out_new_points);
  } else {
    auto kern = ball_group_kernel<int>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, W * 32, smem, st>>>(xyz, feat, C, centers, N, S, radius2, nsample, use_bulk, (int*)out_idx,
                                     out_new_points);
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
