/*
 * ua_b200.h — C ABI of libua_b200.so, the sm_100a implementation of the Uni-Adapter
 * per-sample test-time hot path (point tokenizer, zero-shot head, DOTA / MODE-DOTA cache step).
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the library never allocates, never synchronises and never copies to the host: the caller owns
 *     every buffer (including scratch) and the stream;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: UA_OK (0) or a negative UA_ERR_* code; ua_last_error() gives a message;
 *   - all tensors are dense row-major fp32 unless the parameter says otherwise.
 *
 * Each entry point cites the reference interface (file:line under the reference repo) it replaces.
 */
#ifndef UA_B200_H
#define UA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UA_OK 0
#define UA_ERR_INVALID_ARG (-1)
#define UA_ERR_UNSUPPORTED (-2)
#define UA_ERR_CUDA (-3)

#define UA_ABI_VERSION 1

/* Library identity / diagnostics ------------------------------------------------------------ */
int ua_abi_version(void);
const char* ua_last_error(void);
/* Number of kernel launches issued by this library since load (or since ua_reset_launch_count). */
int64_t ua_launch_count(void);
void ua_reset_launch_count(void);
/* Tuning knobs for experiments (0 = built-in heuristic): "fps_threads", "knn_warps", "modedota_threads",
 * "modedota_v" (float4 per lane of the single-sample cache kernel; -1 disables that kernel), "modedota_groups",
 * "modedota_logprod" (1 = product-form log-determinant, the default; 0 = one logf per element). */
int ua_set_tuning(const char* key, int value);

/* ------------------------------------------------------------------------------------------
 * Tokenizer
 * ---------------------------------------------------------------------------------------- */

/* Farthest-point sampling, one cloud per thread block, running-min distances in registers.
 * Replaces: models/ulip/pointbert/misc.py:40-60 fps(xyz,npoint)            (random start, returns points)
 *           models/openshape/pointnet_util.py:64-86 farthest_point_sample  (random start, returns idx)
 *           models/point_encoder.py:7-14 fps() -> pointnet2_ops.furthest_point_sample + gather_operation
 *                                                                           (start index 0)
 * Arithmetic: dist = ((dx*dx)+(dy*dy))+(dz*dz), every operation rounded separately (no FMA),
 * running min initialised to 1e10, argmax takes the FIRST maximal index.
 *   xyz        [B,N,3] f32
 *   start_idx  [B] i64, or NULL for start index 0 in every cloud
 *   skip_small_norm  bit 0: points with x*x+y*y+z*z <= 1e-3 are never selected/updated (pointnet2_ops quirk, in the
 *              torch arithmetic above); UA_FPS_POINTNET2 (bit 1): the published pointnet2_ops kernel's own arithmetic
 *              (erikwijmans/Pointnet2_PyTorch pointnet2_ops_lib 3.0.0, sampling_gpu.cu; the un-vendored extension behind
 *              models/point_encoder.py:12): d = fma(dz,dz, fma(dx,dx, dy*dy)) (nvcc's contraction of the upstream
 *              expression), points with fma(z,z, fma(x,x, y*y)) <= 1e-3 (double compare) never take part, ties follow
 *              upstream's left-biased reduction tree (lowest bit-reversed owner thread k mod bs, then lowest k), bs = largest power of two <= min(N, 512); N <= UA_FPS_MAX_REG_POINTS.
 *   out_idx    [B,G] i32 or i64 (idx_is_i64), may be NULL
 *   out_centers[B,G,3] f32, may be NULL
 *   scratch    [B,N] f32, only required when N > UA_FPS_MAX_REG_POINTS (else may be NULL)
 */
#define UA_FPS_MAX_REG_POINTS 16384
#define UA_FPS_POINTNET2 2
int ua_fps_f32(const float* xyz, int B, int N, int G, const int64_t* start_idx, int skip_small_norm,
               void* out_idx, int idx_is_i64, float* out_centers, float* scratch, void* stream);

/* kNN grouping fused with gather + centre subtraction + channel concat; the (G,N) distance matrix is
 * never materialised.
 * Replaces: models/point_encoder.py:17-49,99-127 (knn_point/square_distance/Group.forward, Uni3D)
 *           models/ulip/pointbert/dvae.py:116-181 (same, ULIP: no colour)
 * Ranking value: d = ((-2*fma(cz,pz,fma(cy,py,cx*px))) + ((cx*cx+cy*cy)+cz*cz)) + ((px*px+py*py)+pz*pz)
 * (the reference's expanded form, fp32); the k smallest (d, index) pairs in lexicographic order are kept,
 * i.e. ties go to the lower point index. Neighbours are emitted in ascending point-index order (the reference's
 * topk(sorted=False) order is unspecified; everything downstream is permutation-invariant over a group).
 *   xyz      [B,N,3]      rgb [B,N,3] or NULL      centers [B,G,3]
 *   out_idx  [B,G,k] i32/i64 or NULL
 *   out_neigh[B,G,k,3] (xyz - centre) or NULL
 *   out_feat [B,G,k,6] (xyz - centre, rgb) or NULL (requires rgb)
 *   1 <= k <= 128, k <= N
 */
int ua_knn_group_f32(const float* xyz, const float* rgb, const float* centers, int B, int N, int G, int k,
                     void* out_idx, int idx_is_i64, float* out_neigh, float* out_feat, void* stream);

/* Ball-query grouping (first `nsample` points, in index order, with d <= radius2; padded with the first hit).
 * Replaces: models/openshape/pointnet_util.py:89-146 (query_ball_point + sample_and_group).
 * A point is REJECTED iff d > radius2 with d computed as in ua_knn_group_f32.
 *   feat [B,N,C] or NULL (C == 0)
 *   out_idx [B,S,nsample] i32/i64 or NULL
 *   out_new_points [B,S,nsample,3+C] = (xyz - centre, feat) or NULL
 * A centre with no point inside the ball (impossible when centres are cloud points) yields index N-1
 * clamped gathers; the reference raises an index error there.
 */
int ua_ball_group_f32(const float* xyz, const float* feat, int C, const float* centers, int B, int N, int S,
                      float radius2, int nsample, void* out_idx, int idx_is_i64, float* out_new_points,
                      void* stream);

/* out[b,c,g] = in[b,c,idx[b,g]] — pointnet2_ops.gather_operation (models/point_encoder.py:13). */
int ua_gather_points_f32(const float* in, const int32_t* idx, int B, int C, int N, int G, float* out,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * Zero-shot cosine-logit head
 * Replaces: Uni_Adapter.py:21-26,53-75 (softmax_entropy, get_logits_wrapper after the encoder)
 *   x [B,D] raw encoder output; text [num_text,K,D] unit-norm rows (clip_weights = text^T). num_text = 1 is the
 *   reference's single text matrix; num_text = S gives every block of B/S consecutive rows its own matrix
 *   (independent streams whose text residuals are learned separately).
 *   out_xnorm [B,D] = x/||x||; out_logits [B,K] = (scale*xnorm) @ text^T ; out_prob = softmax(logits);
 *   out_entropy [B] = -sum p*log(p+1e-10); out_argmax [B] i32 (first maximal index).
 * out_prob / out_entropy / out_argmax may be NULL.
 * ---------------------------------------------------------------------------------------- */
int ua_head_f32(const float* x, int B, int D, const float* text, int num_text, int K, float scale, float* out_xnorm,
                float* out_logits, float* out_prob, float* out_entropy, int32_t* out_argmax, void* stream);
/* Batched head on the tensor cores (B >= 64): ua_head_prepare_f32 writes xnorm and the (hi, lo) pair of scale*xnorm,
 * ua_gemm_tf32x3_f32 contracts it with the (hi, lo) text rows (K padded to a multiple of 128 with zero rows, ld = Kpad),
 * ua_row_stats_f32 finishes softmax / entropy / first-index argmax over the first K columns of each row. */
int ua_head_prepare_f32(const float* x, int B, int D, float scale, float* out_xnorm, float* out_hi, float* out_lo,
                        void* stream);
int ua_row_stats_f32(const float* logits, int B, int K, long long ld, float* out_prob, float* out_entropy,
                     int32_t* out_argmax, void* stream);

/* ------------------------------------------------------------------------------------------
 * MODE-DOTA (diagonal Gaussian mixture cache)
 * Replaces: dota_mixture.py:117-156 (_get_var/_log_likelihood), :162-234 (fit), :236-267 (predict).
 * State of S independent streams: mu,var [S,K,M,D]; pi,c [S,K,M]; class_counts [S,K] (S = 1 for the
 * reference's single adapter object; S > 1 runs S adapters in lock-step in one launch).
 * One launch evaluates  predict(x_pred) on the CURRENT state (if x_pred != NULL) and then applies
 * fit(x_fit, gamma_class) in place (if x_fit != NULL); each class's (M,D) state tile is read from HBM once
 * and written once (TMA bulk loads through a shared-memory ring; register-resident M-step and direct warp stores
 * on the single-sample path Bp <= 1, B <= 1, D % 128 == 0; bulk stores on the general path).
 *   x_pred [S,Bp,D] -> out_logits [S,Bp,ldo] written at columns [k_out_offset, k_out_offset+K)
 *   x_fit [S,B,D], gamma_class [S,B,ldg] read at columns [k_gamma_offset, k_gamma_offset+K)
 * (the offsets / leading dimensions let a class-sharded rank use the full-width logits / prob_map buffers).
 * Limits: M <= 16, Bp + B <= 160, 2*M*D*4 bytes + 27 KB <= 227 KB.
 * ---------------------------------------------------------------------------------------- */
int ua_modedota_step_f32(const float* x_pred, int Bp, const float* x_fit, const float* gamma_class, int B,
                         int ldg, int k_gamma_offset, float* mu, float* var, float* pi, float* c,
                         float* class_counts, int S, int K, int M, int D, float eps, float* out_logits, int ldo,
                         int k_out_offset, void* stream);

/* Per-stream random inputs of a lock-step step, drawn on the device (CUDA-graph safe): the N(0,1) jitter of the
 * augmented view (Uni_Adapter.py:420-421) and the random FPS start indices of both views (models/ulip/pointbert/
 * misc.py:52, models/openshape/pointnet_util.py:77). Counter-based Philox4x32-10 keyed by seeds[s], counter = (element,
 * *step, purpose): a stream's draws depend only on its own seed and its own step count, not on the co-resident
 * streams or the world size. The kernel advances *step (device memory) itself; done_counter is a zeroed u32 scratch.
 *   noise [S, per_stream] f32; start_idx [2,S] i64 in [0, n_range) (first view, jittered view) or NULL
 */
int ua_stream_rng_f32(const int64_t* seeds, int64_t* step, int S, int64_t per_stream, float* noise,
                      int64_t* start_idx, int n_range, uint32_t* done_counter, void* stream);

/* Fusion of zero-shot and cache logits, Uni_Adapter.py:491-521 (MODE-DOTA, mode=1) or
 * dota_mixture.py:289-293 (DOTA, mode=0: final = clip + w*dota).
 *   w = min(rho * mean(c) / batch, eta) computed on device from c[count_c]; row r reads c + r*c_row_stride
 *   (stride 0: one adapter for all rows; stride K*M: one adapter per row / stream)
 *   (c_sum_override >= 0 replaces sum(c): used by class-sharded ranks, closed form K + fits*B)
 *   dota_logits may be fp16 (dota_is_f16) as DOTA.predict returns half.
 *   out_final [R,K], out_argmax [R] i32, out_scaled_dota [R,K] or NULL
 */
int ua_fuse_logits_f32(const float* clip_logits, const void* dota_logits, int dota_is_f16, int R, int K,
                       const float* c, int count_c, int c_row_stride, float c_sum_override, float c_count_total,
                       float rho, float eta, float batch, int mode, float* out_final, int32_t* out_argmax,
                       float* out_scaled_dota, void* stream);

/* ------------------------------------------------------------------------------------------
 * Residual text-feature learning
 * Replaces: Uni_Adapter.py:191-270 (compute_text_alignment_loss, forward + autograd backward) and the inner loop
 * Uni_Adapter.py:443-476 (zero_grad / backward / Adam.step x10 on text_residuals, torch.optim.Adam defaults).
 *   text0 [K,D] (text0_stream_stride = 0) or [S,K,D] (= K*D): the fixed initial text features
 *   residual, adam_m, adam_v [S,K,D] updated in place; adam_t [S] i32 on the device: Adam steps taken so far,
 *   advanced by `iters` (on the device, so a captured CUDA graph keeps counting)
 *   mu,var [S,K,M,D], pi [S,K,M]: the MODE-DOTA state (read only)
 *   out_text [S,K,D] = normalize(text0 + residual) after the updates (the next sample's clip_weights^T)
 *   out_loss [S,iters] or NULL: the alignment loss before each step
 *   scratch: ua_residual_scratch_floats(S,K,M,D) floats, 16-byte aligned.
 * ua_align_loss_grad_f32 evaluates one loss / likelihood matrix [S,K,K] / gradient w.r.t. the residual [S,K,D]
 * (any of them NULL to skip) without touching the residual; out_emb [S,K,D] receives the normalised embeddings.
 * Limits: D % 128 == 0, M in {4,8,12,16}, K*K*M < 2^31 (any realistic class count: past K ~ 160 the likelihood matrix is
 * read from global memory, past K ~ 230 P lives in the scratch too, past K*M = 880 the backward kernel walks the columns in
 * chunks; the reference's own (K,K,M,D) broadcast is 44 GB at K = 1156).
 * ---------------------------------------------------------------------------------------- */
long long ua_residual_scratch_floats(int S, int K, int M, int D);
int ua_residual_learn_f32(const float* text0, long long text0_stream_stride, float* residual, float* adam_m,
                          float* adam_v, int32_t* adam_t, const float* mu, const float* var, const float* pi, int S,
                          int K, int M, int D, float eps, double lr, double beta1, double beta2, double adam_eps,
                          int iters, float* out_text, float* out_loss, float* scratch, long long scratch_floats,
                          void* stream);
int ua_align_loss_grad_f32(const float* text0, long long text0_stream_stride, const float* residual, const float* mu,
                           const float* var, const float* pi, int S, int K, int M, int D, float eps, float* out_emb,
                           float* out_loss, float* out_lm, float* out_grad, float* scratch, long long scratch_floats,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * fp32-accurate GEMM on tcgen05 tensor cores (3xTF32), the dense contraction of the path:
 *   C[M,N] = epilogue(A[M,K] . W[N,K]^T)   — the layout of nn.Linear / 1x1 Conv1d weights
 * Replaces: the zero-shot head contraction at batch >= 64 (Uni_Adapter.py:61-62), the mini-PointNet group encoder's
 * 1x1 convolutions (models/ulip/pointbert/dvae.py:201-215, models/point_encoder.py:145-159).
 * Operands come as (hi, lo) pairs: hi = tf32(x), lo = x - hi (ua_split_tf32_f32, or the OUT_SPLIT epilogue of the
 * producing GEMM). Epilogue: + bias[N], + group_bias[row/32, N], + residual[M,ldo], act (0 none, 1 ReLU, 2 erf-GELU),
 * then any of
 *   out [M,ldo] fp32;  (out_hi, out_lo) [M,ldo];  gmax [M/32, N] = max over each 32 consecutive rows (one point
 *   group; needs M % 32 == 0) with an optional (gmax_hi, gmax_lo) copy.
 * Limits: N % 128 == 0, K % 32 == 0, 16-byte aligned operands, lda/ldw/ldo multiples of 4.
 * ---------------------------------------------------------------------------------------- */
int ua_split_tf32_f32(const float* x, float* hi, float* lo, long long n, void* stream);
/* (hi, lo) of relu?(x[M,C] . w[N,C]^T + b[N]) for the C = 3 / 6 input channels of the group encoder's first conv. */
int ua_pointwise_linear_split_f32(const float* x, const float* w, const float* b, int relu, long long M, int C, int N,
                                  float* out_hi, float* out_lo, void* stream);
int ua_gemm_tf32x3_f32(const float* a_hi, const float* a_lo, long long lda, const float* w_hi, const float* w_lo,
                       long long ldw, int M, int N, int K, const float* bias, const float* group_bias,
                       const float* residual, int act, float* out, float* out_hi, float* out_lo, long long ldo,
                       float* gmax, float* gmax_hi, float* gmax_lo, void* stream);
/* (hi, lo) of LayerNorm(x (+ pos)) over the last dimension C (C % 128 == 0, C <= 1024); out_sum = x + pos (optional):
 * the operand producer in front of a transformer block's GEMMs (one warp per row). */
int ua_layernorm_split_f32(const float* x, const float* pos, const float* gamma, const float* beta, float eps,
                           long long rows, int C, float* out_sum, float* out_hi, float* out_lo, void* stream);

/* ------------------------------------------------------------------------------------------
 * fp32-accurate multi-head self-attention on tcgen05 (3xTF32), head dimension 64:
 *   O = softmax(Q K^T / 8) V  per (batch, head) — the encoder blocks' attention (F.scaled_dot_product_attention in
 *   models/ulip/pointbert/point_encoder.py's Attention / timm's blocks), kept at fp32-class accuracy.
 * ua_attn_prepare_f32: qkv [B*N, 3*H*64] (the qkv Linear's output, columns ordered (q|k|v, head, d)) -> (hi, lo) pairs
 *   Q, K [B*H, N, 64] (Q pre-scaled by log2(e)/8: base-2 softmax) and V^T [B*H, 64, Npad], Npad = ua_attn_padded_tokens(N) (zero padded).
 * ua_attention_f32: -> (out_hi, out_lo) [B*N, H*64], the A operand of the projection GEMM.
 * ---------------------------------------------------------------------------------------- */
long long ua_attn_padded_tokens(int N);
int ua_attn_prepare_f32(const float* qkv, int B, int N, int H, float* q_hi, float* q_lo, float* k_hi, float* k_lo,
                        float* vt_hi, float* vt_lo, void* stream);
int ua_attention_f32(const float* q_hi, const float* q_lo, const float* k_hi, const float* k_lo, const float* vt_hi,
                     const float* vt_lo, int B, int N, int H, float* out_hi, float* out_lo, void* stream);

/* ------------------------------------------------------------------------------------------
 * DOTA (full covariance)
 * Replaces: dota.py:41-63 (fit), :66-69 (update), :72-87 (predict).
 *   fit:  mu [K,D], c [K], Sigma [K,D,D], overall [D,D] updated in place from x [B,D], y [B,K].
 *   predict: x_h [R,D] f16, Lambda_h [D,D] f16, mu [K,D] f32 -> out_scores_h [R,K] f16, rounding to fp16
 *            at the reference's rounding points (M, W, M*W, 0.5*sum, X@W, difference).
 * ---------------------------------------------------------------------------------------- */
int ua_dota_fit_f32(const float* x, const float* y, int B, float* mu, float* c, float* Sigma, float* overall,
                    int K, int D, void* stream);
int ua_dota_predict_f16(const void* x_h, int R, const void* Lambda_h, const float* mu, int K, int D,
                        void* out_scores_h, void* stream);
/* A = (1-eps)*overall + eps*I  (the matrix dota.py:67-68 inverts), written to out [D,D]. */
int ua_dota_regularize_f32(const float* overall, int D, float eps, float* out, void* stream);
/* update (dota.py:66-69): Lambda = inverse((1-eps)*overall + eps*I).half(), one cooperative launch (register-resident
 *   block Gauss-Jordan, no pivoting: the matrix is SPD). D % 16 == 0, D <= 1536 (else UA_ERR_UNSUPPORTED: the caller
 *   keeps a library inverse for such sizes). workspace: ua_dota_update_workspace_bytes(D) bytes of device scratch,
 *   256-byte aligned. out_lambda_h [D,D] f16 and/or out_lambda_f32 [D,D] (either may be NULL, not both). */
long long ua_dota_update_workspace_bytes(int D);
int ua_dota_update_f32(const float* overall, int D, float eps, void* workspace, void* out_lambda_h,
                       float* out_lambda_f32, void* stream);

/* One cache pass for a whole batch-1 sample step of the reference loop (Uni_Adapter.py:416-430):
 *     logits = predict(x.half())  ->  fit(x, prob_map)  ->  fit(x_aug, prob_map)
 * Each class's (M,D) tile is read from HBM once and written once for all three operations (16*K*M*D bytes per sample
 * step instead of 32 with ua_modedota_step_f32 called twice); bit-identical to that two-call sequence.
 *   x_fit [S,D] the normalised sample (predict uses its fp16 rounding, Uni_Adapter.py:416), x_fit2 [S,D] the normalised
 *   jittered view or NULL (one fit), gamma_class [S,ldg] prob_map read at columns [k_gamma_offset, +K),
 *   out_logits [S,ldo] written at columns [k_out_offset, +K) or NULL (no predict).
 * Limits: D % 128 == 0, M <= 16, 16-byte aligned pointers, 4*M*D*4 bytes of tiles must fit in shared memory
 * (UA_ERR_UNSUPPORTED otherwise: the caller falls back to ua_modedota_step_f32).
 */
int ua_modedota_sample_step_f32(const float* x_fit, const float* x_fit2, const float* gamma_class, int ldg,
                                int k_gamma_offset, float* mu, float* var, float* pi, float* c, float* class_counts,
                                int S, int K, int M, int D, float eps, float* out_logits, int ldo, int k_out_offset,
                                void* stream);

/* ------------------------------------------------------------------------------------------
 * Class-sharded sample step (BASELINE cfg 4: the Objaverse-LVIS cache split over P GPUs by class):
 * the pass above over this rank's classes with both logit exchanges and the fusion in the SAME persistent kernel.
 * The reference has no multi-GPU path; this is the exchange step the north star names (logit all-gather per step),
 * done with stores into NVLink-mapped peer memory from the kernel itself instead of a collective call.
 * Class ranges are contiguous, the first K mod P ranks own one class more. Every rank passes the same sample.
 *   per rank (struct below, a HOST array of n_ranks entries, copied into the kernel parameters):
 *     text_local [K_local, D] != NULL: x_fit / x_fit2 [D] are the RAW encoder outputs of the sample and of its jittered
 *       view (x_fit2 may be NULL); the kernel L2-normalises them and forms the zero-shot logits of its classes itself
 *       (the arithmetic of ua_head_f32, Uni_Adapter.py:56-57) -- the whole adapter step is this one launch;
 *     text_local == NULL: x_fit / x_fit2 are already normalised and clip_local [K_local] holds the zero-shot logits;
 *     the state shard (mu, var, pi, c, class_counts over K_local classes), and
 *     peer_recv: device array [P] of float* -- every rank's symmetric receive buffer, 2*P*2*K_pad floats
 *                ([parity][rank][0: zero-shot | 1: cache][K_pad]), mapped into this process;
 *     peer_flag: device array [P] of int* -- every rank's 2*P flags ([exchange][rank]), zero-initialised;
 *     seq (device int, step counter, advanced by the kernel; all ranks start from 0), err (device int: 1 = a peer's
 *     zero-shot logits did not arrive, cache untouched; 2 = a peer's cache logits did not arrive or the peer aborted;
 *     outputs are NaN / -1 then), done (TWO zeroed u32 of scratch), c_sum (device float: sum of the soft counts so far, K at
 *     start; the kernel adds the fits of the step -- closed form);
 *     out_final [K], out_argmax [1], out_clip / out_dota [K] or NULL: replicated results of the step.
 *   n_ranks = 1: this process' rank (ONE struct). n_ranks = P <= 8: single-GPU emulation of all P ranks in one
 *   cooperative launch (tests); all "peer" buffers then live on the one device.
 * The launch is cooperative (all CTAs co-resident: they wait for one another through the peers).
 * ---------------------------------------------------------------------------------------- */
typedef struct ua_shard_rank {
  const float* x_fit;
  const float* x_fit2;
  const float* text_local;
  const float* clip_local;
  float* mu;
  float* var;
  float* pi;
  float* c;
  float* class_counts;
  float* const* peer_recv;
  int* const* peer_flag;
  int* seq;
  int* err;
  uint32_t* done;
  float* c_sum;
  float* out_final;
  int* out_argmax;
  float* out_clip;
  float* out_dota;
  int rank;
  int reserved;
} ua_shard_rank;
/* diagnosis: phase time stamps (ns) of the last sharded step launched with ua_set_tuning("sample_trace", 1) */
int ua_debug_sample_trace(int64_t* host_out16);
int ua_modedota_sharded_step_f32(const ua_shard_rank* ranks, int n_ranks, int P, int K, int K_pad, int M, int D,
                                 float eps, float rho, float eta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UA_B200_H */
