/*
 * TEST INFRASTRUCTURE — CPU oracle for the point tokenizer. Not part of the product: only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load this.
 *
 * Plain-C restatement of the reference's tokenizer arithmetic (paths relative to the reference repo):
 *   oracle_fps        models/ulip/pointbert/misc.py:40-60 fps(), models/openshape/pointnet_util.py:64-86
 *                     farthest_point_sample(); with start index 0 it also stands for models/point_encoder.py:7-14
 *                     (pointnet2_ops.furthest_point_sample, un-vendored CUDA) in torch arithmetic.
 *   oracle_fps_pointnet2  the PUBLISHED algorithm of that un-vendored dependency (erikwijmans/Pointnet2_PyTorch,
 *                     pointnet2_ops_lib 3.0.0, _ext-src/src/sampling_gpu.cu furthest_point_sampling_kernel +
 *                     sampling.cpp: temp = 1e10, first sample 0), call site models/point_encoder.py:7-14: strided point
 *                     ownership, per-thread strict '>' scan, pairwise tree reduction keeping the LEFT operand on ties,
 *                     distances with the FMA contraction nvcc applies to the upstream expression (probed with nvcc
 *                     12.9, oracle/pointnet2_contraction_probe.cu), '|p|^2 <= 1e-3' evaluated in double. The upstream
 *                     binary cannot run here (parity unpinned against it); this pins the kernel to the published source.
 *   oracle_sqdist     square_distance(): models/point_encoder.py:30-49, dvae.py:130-149, pointnet_util.py:20-41
 *   oracle_knn        knn_point(): models/point_encoder.py:17-28, dvae.py:116-127  (topk largest=False)
 *   oracle_ball       query_ball_point(): models/openshape/pointnet_util.py:89-110
 *
 * Rounding is made explicit: build with -ffp-contract=off so that only the fmaf() calls below fuse.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* torch.sum((xyz - c) ** 2, -1): three separately rounded squares added left to right (misc.py:57). */
static inline float sqdist3(const float* p, const float* c) {
  const float dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2];
  const float sx = dx * dx, sy = dy * dy, sz = dz * dz;
  return (sx + sy) + sz;
}

static inline float sqnorm3(const float* p) {
  const float sx = p[0] * p[0], sy = p[1] * p[1], sz = p[2] * p[2];
  return (sx + sy) + sz;
}

/* square_distance: dist = -2 * matmul(src, dst^T); dist += sum(src**2); dist += sum(dst**2)
 * (point_encoder.py:46-48). The K=3 matmul is the fma chain fma(a2,b2,fma(a1,b1,a0*b0)). */
static inline float expanded_sqdist(const float* c, float cn, const float* p, float pn) {
  const float dot = fmaf(c[2], p[2], fmaf(c[1], p[1], c[0] * p[0]));
  const float t = -2.0f * dot;
  const float u = t + cn;
  return u + pn;
}

/* xyz [B,N,3]; start [B] or NULL (-> 0); out_idx [B,G]; clouds [b0,b1) (callers thread over clouds) */
void oracle_fps(const float* xyz, int b0, int b1, int N, int G, const int64_t* start, int skip_small,
                int64_t* out_idx) {
  for (int b = b0; b < b1; ++b) {
    const float* cloud = xyz + (size_t)b * N * 3;
    float* dist = (float*)malloc(sizeof(float) * (size_t)N);
    unsigned char* skip = (unsigned char*)calloc((size_t)N, 1);
    for (int p = 0; p < N; ++p) {
      dist[p] = 1e10f; /* misc.py:51 */
      if (skip_small && !(sqnorm3(cloud + 3 * p) > 1e-3f)) { skip[p] = 1; dist[p] = 0.0f; }
    }
    int64_t far = start ? start[b] : 0; /* misc.py:52 (random) / pointnet2_ops (0) */
    for (int i = 0; i < G; ++i) {
      out_idx[(size_t)b * G + i] = far; /* misc.py:55 */
      const float* c = cloud + 3 * far;
      float best = -1.0f;
      int64_t besti = 0;
      for (int p = 0; p < N; ++p) {
        if (!skip[p]) {
          const float d = sqdist3(cloud + 3 * p, c); /* misc.py:57 */
          if (d < dist[p]) dist[p] = d;              /* misc.py:58 torch.min */
        }
        if (dist[p] > best) { best = dist[p]; besti = p; } /* misc.py:59 torch.max -> first maximal index */
      }
      far = besti;
    }
    free(dist);
    free(skip);
  }
}

/* pointnet2_ops furthest_point_sampling_kernel<block_size> restated thread by thread.
 * block_size = opt_n_threads(N): the largest power of two <= N, clipped to [1, 512] (cuda_utils.h). */
static int pn2_block_size(int n) {
  int p = 1;
  while (p * 2 <= n && p * 2 <= 512) p *= 2;
  return p;
}

void oracle_fps_pointnet2(const float* xyz, int b0, int b1, int N, int G, int64_t* out_idx) {
  const int bs = pn2_block_size(N);
  float* dists = (float*)malloc(sizeof(float) * (size_t)bs);
  int* dists_i = (int*)malloc(sizeof(int) * (size_t)bs);
  float* temp = (float*)malloc(sizeof(float) * (size_t)N);
  for (int b = b0; b < b1; ++b) {
    const float* dataset = xyz + (size_t)b * N * 3;
    for (int p = 0; p < N; ++p) temp[p] = 1e10f;            /* sampling.cpp: torch::full({B, N}, 1e10) */
    int old = 0;
    out_idx[(size_t)b * G] = 0;
    for (int j = 1; j < G; ++j) {
      const float x1 = dataset[old * 3 + 0], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
      for (int tid = 0; tid < bs; ++tid) {
        int besti = 0;
        float best = -1.0f;
        for (int k = tid; k < N; k += bs) {
          const float x2 = dataset[k * 3 + 0], y2 = dataset[k * 3 + 1], z2 = dataset[k * 3 + 2];
          /* nvcc: mag = fma(z2,z2, fma(x2,x2, y2*y2)); compared in double against 1e-3 */
          const float mag = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
          if ((double)mag <= 1e-3) continue;
          const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
          const float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));   /* nvcc's contraction of the upstream sum */
          const float d2 = d < temp[k] ? d : temp[k];
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      for (int half = bs / 2; half >= 1; half /= 2) {         /* __update(dists, dists_i, tid, tid + half) */
        for (int tid = 0; tid < half; ++tid) {
          const float v1 = dists[tid], v2 = dists[tid + half];
          const int i1 = dists_i[tid], i2 = dists_i[tid + half];
          dists[tid] = v1 > v2 ? v1 : v2;
          dists_i[tid] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      out_idx[(size_t)b * G + j] = old;
    }
  }
  free(dists);
  free(dists_i);
  free(temp);
}

/* full (G,N) matrix for one cloud, for spot checks */
void oracle_sqdist(const float* centers, int G, const float* xyz, int N, float* out) {
  for (int g = 0; g < G; ++g) {
    const float cn = sqnorm3(centers + 3 * g);
    for (int p = 0; p < N; ++p) out[(size_t)g * N + p] = expanded_sqdist(centers + 3 * g, cn, xyz + 3 * p, sqnorm3(xyz + 3 * p));
  }
}

typedef struct { float d; int32_t i; } pair_t;
static int cmp_pair(const void* a, const void* b) {
  const pair_t* x = (const pair_t*)a; const pair_t* y = (const pair_t*)b;
  if (x->d < y->d) return -1;
  if (x->d > y->d) return 1;
  return (x->i > y->i) - (x->i < y->i);
}

/* k smallest (distance, index) pairs per centre, ascending; ties -> lower index. out_idx [B,G,k], out_d optional;
 * centres [g0,g1) of every cloud (callers thread over centre ranges) */
void oracle_knn(const float* xyz, const float* centers, int B, int N, int G, int g0, int g1, int k, int64_t* out_idx,
                float* out_d) {
  for (int b = 0; b < B; ++b) {
    for (int g = g0; g < g1; ++g) {
      const float* cloud = xyz + (size_t)b * N * 3;
      const float* c = centers + ((size_t)b * G + g) * 3;
      const float cn = sqnorm3(c);
      pair_t* arr = (pair_t*)malloc(sizeof(pair_t) * (size_t)N);
      for (int p = 0; p < N; ++p) {
        arr[p].d = expanded_sqdist(c, cn, cloud + 3 * p, sqnorm3(cloud + 3 * p));
        arr[p].i = p;
      }
      qsort(arr, (size_t)N, sizeof(pair_t), cmp_pair);
      for (int j = 0; j < k; ++j) {
        out_idx[((size_t)b * G + g) * k + j] = arr[j].i;
        if (out_d) out_d[((size_t)b * G + g) * k + j] = arr[j].d;
      }
      free(arr);
    }
  }
}

/* first nsample indices (ascending) with !(d > r2); padded with the first; out_idx [B,S,nsample].
 * A centre without any hit is filled with N (the reference then fails on the gather). */
void oracle_ball(const float* xyz, const float* centers, int B, int N, int S, int g0, int g1, float r2, int nsample,
                 int64_t* out_idx) {
  for (int b = 0; b < B; ++b) {
    for (int g = g0; g < g1; ++g) {
      const float* cloud = xyz + (size_t)b * N * 3;
      const float* c = centers + ((size_t)b * S + g) * 3;
      const float cn = sqnorm3(c);
      int64_t* row = out_idx + ((size_t)b * S + g) * nsample;
      int cnt = 0;
      for (int p = 0; p < N && cnt < nsample; ++p) {
        const float d = expanded_sqdist(c, cn, cloud + 3 * p, sqnorm3(cloud + 3 * p));
        if (!(d > r2)) row[cnt++] = p; /* pointnet_util.py:105 */
      }
      const int64_t first = cnt > 0 ? row[0] : N;
      for (int j = cnt; j < nsample; ++j) row[j] = first; /* pointnet_util.py:107-109 */
    }
  }
}
