/* b200rec.h - C ABI of libb200rec.so, the B200 (sm_100a) BPR-MF training and
 * scoring engine that drops in behind yoongi0428/RecSys_PyTorch's hot path.
 *
 * Two groups of entry points:
 *
 *  (1) HOST-buffer drop-ins with the exact argument lists of the reference's own
 *      native layer (what its Cython shims bind today).  Buffers are caller-owned
 *      host memory; copies happen inside the call.
 *  (2) DEVICE-pointer engine calls used by the host-side mirror of the
 *      reference's model / generator / evaluator interface
 *      (recsys_pytorch_b200/*.py).  Pointers come from torch.Tensor.data_ptr();
 *      the caller owns all memory; `stream` is a cudaStream_t (0 = default).
 *
 * Conventions: every function returns 0 on success, a negative B200REC_E* code on
 * failure, and never aborts; b200rec_last_error() returns a thread-local message.
 * (The reference's native functions are `void` with no error reporting and UB on
 * bad sizes - SURVEY section 8(b); the codes are an addition, the happy path is
 * identical.)  One host thread per device at a time.  No CPU fallback exists:
 * without a usable CUDA device every call returns B200REC_ECUDA.
 *
 * Embedding tables are row-major fp32 `[rows, ld]` with `ld >= d`, `ld % 4 == 0`,
 * base 16-byte aligned, and columns d..ld-1 zero (they stay zero under every
 * update).  Item / user ids are int32.  CSR = int64 indptr + int32 sorted indices.
 */
#ifndef B200REC_H
#define B200REC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200REC_OK 0
#define B200REC_EINVAL (-1)  /* bad argument (null pointer, size, alignment)   */
#define B200REC_ECUDA (-2)   /* CUDA runtime / launch failure, or no device    */
#define B200REC_ENOMEM (-3)  /* workspace too small / allocation failed        */
#define B200REC_EUNSUPPORTED (-4)

const char *b200rec_last_error(void);
int b200rec_version(void);            /* 100*major + minor */
/* number of kernels this library has launched in the calling process */
int64_t b200rec_launch_count(void);

/* ------------------------------------------------------------------------- *
 * (1) drop-ins for the reference's native evaluation layer (HOST buffers)
 * ------------------------------------------------------------------------- */

/* replaces c_top_k_array_index, evaluation/backend/cython/include/func.h:22-31
 * (bound by evaluation/backend/cython/func.pyx:8-10,22).  scores: fp32
 * [rows_num, columns_num] C-contiguous; rankings: int32 [rows_num, max_k], ids of
 * the max_k largest scores, descending; ties -> smaller id first (upstream:
 * unspecified).  Requires 1 <= max_k <= columns_num and max_k <= 1024. */
int b200rec_top_k_array_index(const float *scores_pt, int columns_num, int rows_num,
                              int max_k, int *rankings_pt);

/* replaces evaluate_holdout, evaluation/backend/cython/include/holdout.h:20-103
 * (bound by holdout_func.pyx:6-10,40).  results: fp32 [users_num, 3*K_len] laid
 * out [Prec@Ks.., Recall@Ks.., NDCG@Ks..] per user. */
int b200rec_evaluate_holdout(int users_num, const int *rankings, int max_k, const int *Ks,
                             int K_len, int **ground_truths, const int *ground_truths_num,
                             float *results);

/* replaces evaluate_loo, evaluation/backend/cython/include/loo.h:20-85 (bound by
 * loo_func.pyx:6-9,35).  results: fp32 [users_num, 2*K_len] = [HR@Ks.., NDCG@Ks..] */
int b200rec_evaluate_loo(int users_num, const int *rankings, int max_k, const int *Ks,
                         int K_len, int **ground_truths, float *results);

/* ------------------------------------------------------------------------- *
 * (2) device-pointer engine
 * ------------------------------------------------------------------------- */

/* models/MF.py:32-42 MF.forward: out[b] = sum_k U[users[b],k] * V[items[b],k] */
int b200rec_mf_forward(const float *U, const float *V, int ld, int d, const int32_t *users,
                       const int32_t *items, int n, float *out, void *stream);

/* Sink of the fused BPR step */
#define B200REC_SINK_UPDATE 0 /* tables updated in place: W += -lr*(grad)  (one fused kernel) */
#define B200REC_SINK_STAGE 1  /* per-triple delta rows -> `stage` [B,3,ld]; apply with b200rec_bpr_apply */
#define B200REC_SINK_GRAD 2   /* raw gradient rows accumulated into dense gU/gV (embedding_dense_backward) */
#define B200REC_SINK_NONE 3   /* forward only: loss_sum / x_out (models/MF.py:99-107 without backward) */

/* flags */
#define B200REC_F_USERS_UNIQUE 1 /* no user id repeats inside the batch: user rows use plain vector stores */
#define B200REC_F_TMA_GATHER 2   /* rows gathered with cp.async.bulk (TMA) into shared memory */
#define B200REC_F_ITEM_DELTA 4   /* SINK_UPDATE: item-row deltas accumulate into dense gV (user-sharded layout) */
#define B200REC_F_GENERIC 8      /* force the general kernel where the lean d=128/256 fast path would be taken */
#define B200REC_F_ASYNC_GATHER 16 /* fast path with the deep cp.async (LDGSTS) per-warp row ring */
#define B200REC_F_ITEM_DELTA_BF16 32 /* with F_ITEM_DELTA: gV is a bf16 [num_items, ld] buffer (REDG.ADD.BF16x4) */
#define B200REC_F_L2_HINTS 64    /* fast path: user rows evict-first, item rows / item deltas evict-last in L2 */
/* b200rec_p2p_step only */
#define B200REC_F_P2P_ROUND_ROBIN 256 /* measurement only: visit ALL source ranks round-robin by chunk (slow beyond 2 ranks) */
#define B200REC_F_P2P_PURE_SEQUENTIAL 2048 /* measurement only: no local chunks interleaved with the remote stream        */
#define B200REC_F_P2P_NO_UWRITE 512  /* measurement only: do not write the user row back (results are then wrong)    */
#define B200REC_F_P2P_NO_UREAD 1024  /* measurement only: do not read the user row (a constant is used instead)      */

typedef struct b200rec_bpr_args {
    float *U;              /* [num_users, ld]  user_embedding.weight (models/MF.py:23)   */
    float *V;              /* [num_items, ld]  item_embedding.weight (models/MF.py:24)   */
    int32_t ld, d;
    int32_t num_users, num_items;
    const int32_t *users;  /* [B]  required                                               */
    const int32_t *pos;    /* [B]  or NULL -> sampled on device from the CSR row          */
    const int32_t *neg;    /* [B]  or NULL -> sampled on device (uniform over non-positives) */
    int32_t B;
    /* CSR of train positives, required when pos or neg is NULL (data/generators.py:168-201) */
    const int64_t *csr_indptr;
    const int32_t *csr_indices;
    uint64_t seed, step;   /* counter RNG key (seed, step, triple index, draw) */
    int32_t *out_pos, *out_neg; /* optional [B]: the triples actually used */
    float lr;              /* SGD step size                                   */
    float reg;             /* per-occurrence L2 (0 = the reference's loss)    */
    int32_t sink;          /* B200REC_SINK_*                                  */
    int32_t flags;         /* B200REC_F_*                                     */
    float *stage;          /* SINK_STAGE: [B,3,ld] fp32                       */
    float *gU, *gV;        /* SINK_GRAD: dense [num_users,ld], [num_items,ld] (caller zeroes) */
    double *loss_sum;      /* optional device scalar, += sum_b -log sigmoid(x_b) */
    float *x_out;          /* optional [B]: x_b = s(u,i) - s(u,j)             */
    /* multi-GPU layouts (0 / NULL = single device) */
    int32_t item_lo, item_hi; /* item-sharded: V holds rows [item_lo,item_hi); a rank processes only the
                                 triples whose sampled positive is in range and samples negatives from it */
    float *udelta;         /* item-sharded: user delta rows go to udelta[t] ([B,ld]) instead of U        */
    float inv_batch;       /* >0: overrides 1/B in g = -sigmoid(-x)/B (global batch of a sharded step)   */
} b200rec_bpr_args;

/* models/MF.py:63-68 (zero_grad -> process_one_batch -> backward -> step) as ONE
 * kernel: [sample] -> gather 3 rows -> warp dot pair -> g=-sigmoid(-x)/B ->
 * scatter.  With SINK_UPDATE duplicates inside a batch see each other's partial
 * updates (Hogwild inside a step); SINK_STAGE + b200rec_bpr_apply reproduces
 * autograd's "all gradients from pre-step weights" semantics exactly. */
int b200rec_bpr_step(const b200rec_bpr_args *args, void *stream);
/* diagnostics: which kernel the last b200rec_bpr_step of this process dispatched to, e.g.
 * "bpr_step_group_kernel<8,32,0,1,1,0>" (G, float4 per row, prefetch sets, users-unique, loss, item-delta) */
const char *b200rec_last_step_kernel(void);

/* Pointwise MF step (models/MF.py:63-68 with the pointwise branch MF.py:101-102):
 *   x_b = U[users[b]] . V[items[b]];  loss_kind 0: F.binary_cross_entropy_with_logits(x, ratings) ('ce', MF.py:21),
 *   loss_kind 1: F.mse_loss(x, ratings), both reduction='mean' (scale inv_batch, 0 = 1/B).
 * sink = B200REC_SINK_UPDATE (rows updated in place with -lr * (grad + reg/B * row), vector atomics),
 *        B200REC_SINK_GRAD (raw gradient rows accumulated into dense gU/gV, for the reference's dense Adam) or
 *        B200REC_SINK_NONE (loss only).  loss_sum (device double, may be NULL) receives the SUM of per-sample losses.
 * Pinned to the reference's outputs (tests/golden/tiny_pointwise.npz, ml100k_pointwise.npz) through
 * oracle/bpr_oracle.py::pointwise_loss/_grads; device tests tests/test_gpu_pointwise.py, tests/test_gpu_parity.py. */
int b200rec_pointwise_step(float *U, float *V, int ld, int d, const int32_t *users, const int32_t *items,
                           const float *ratings, int B, int loss_kind, float lr, float reg, int sink,
                           float *gU, float *gV, double *loss_sum, float inv_batch, void *stream);

/* data/generators.py:168-201 on the device, sampling only: for each users[t] draw
 * (pos, neg) with the same counter RNG the fused step uses for (seed, step, t).
 * out_pos[t] = out_neg[t] = -1 for users without positives. */
int b200rec_sample_triples(const int32_t *users, int B, const int64_t *csr_indptr,
                           const int32_t *csr_indices, int num_items, uint64_t seed, uint64_t step,
                           int32_t *out_pos, int32_t *out_neg, void *stream);

/* second phase of the exact step: W[row] += stage rows (vector atomics). */
int b200rec_bpr_apply(float *U, float *V, int ld, const int32_t *users, const int32_t *pos,
                      const int32_t *neg, int B, const float *stage, void *stream);

/* W[ids[t]] += scale * delta[t] for t < n (vector atomics): applies the exchanged user-gradient
 * rows of the item-sharded layout to every replica. */
int b200rec_rows_add(float *W, int ld, const int32_t *ids, int n, const float *delta, int ld_delta,
                     float scale, void *stream);

/* W[0..n) += float(delta_bf16[0..n)) - applies the all-reduced bf16 item-delta buffer (n % 4 == 0). */
int b200rec_add_bf16(float *W, const void *delta_bf16, int64_t n, void *stream);

/* dense SGD / Adam sweeps: torch.optim.SGD(lr) and torch.optim.Adam(lr, betas,
 * eps, weight_decay=0) as constructed at models/MF.py:30 - every element moves. */
int b200rec_sgd_dense(float *param, const float *grad, int64_t n, float lr, void *stream);
/* "update in place, exchange the difference" (user-sharded multi-GPU layout; no reference counterpart):
 *   delta_diff:  d_wire = d_own = W - snapshot      delta_apply:  W += d_sum - d_own      (n a multiple of 4) */
int b200rec_delta_diff(const float *W, const float *snapshot, float *d_wire, float *d_own, int64_t n, void *stream);
int b200rec_delta_apply(float *W, const float *d_sum, const float *d_own, int64_t n, void *stream);
/* W = snapshot + scale * d_sum (n a multiple of 4): combines all-reduced per-rank differences of a replicated table */
int b200rec_snap_apply(float *W, const float *snapshot, const float *d_sum, float scale, int64_t n, void *stream);
/* W += delta; delta = 0 (n a multiple of 4): applies an all-reduced delta buffer and clears it in one pass */
int b200rec_add_clear(float *W, float *delta, int64_t n, void *stream);
int b200rec_adam_dense(float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                       int64_t n, float lr, float beta1, float beta2, float eps, int step,
                       void *stream);

/* ------------------------------------------------------------------------- *
 * multi-GPU: item table sharded by item-id range (north_star), user table sharded by user-id range, user rows
 * exchanged through NVSwitch peer memory INSIDE the fused step (csrc/p2p.cu).  The reference is single-device; the
 * per-step semantics are those of b200rec_bpr_step.  One process per GPU; pointers of other ranks are peer-mapped
 * (b200rec_peer_import) or, for single-process emulation in the tests, plain local pointers.
 * ------------------------------------------------------------------------- */
#define B200REC_MAX_RANKS 16
#define B200REC_PEER_HANDLE_BYTES 64

typedef struct b200rec_p2p_route_args {
    const int32_t *users;      /* [B] LOCAL row ids of this rank's batch users                                  */
    const int32_t *pos, *neg;  /* [B] GLOBAL item ids, or NULL -> sampled on device (neg from owner(pos)'s range) */
    int32_t B;
    const int64_t *csr_indptr; /* CSR shard: rows = this rank's users (local ids), columns = GLOBAL item ids     */
    const int32_t *csr_indices;
    uint64_t seed, step;
    int32_t world, rank;
    int32_t item_bounds[B200REC_MAX_RANKS + 1]; /* rank r holds items [item_bounds[r], item_bounds[r+1])         */
    int32_t head;              /* items [0, head) are REPLICATED on every rank (item_bounds[0] == head): a triple whose
                                  positive is a head item stays on the routing rank; negatives are drawn from
                                  head U owner's range.  0 = pure range sharding                                 */
    /* outbox in this rank's exported memory: triples for owner d at out_*[d*cap + k], k < out_cnt[d]            */
    int32_t *out_u, *out_i, *out_j, *out_cnt;
    int32_t cap;               /* >= B                                                                           */
    int32_t *dbg_pos, *dbg_neg;/* optional [B]: the (pos, neg) drawn for users[t], -1 = skipped                  */
} b200rec_p2p_route_args;

/* sample + bucket this rank's batch by owner(pos) into its outbox (local kernel; zeroes out_cnt first) */
int b200rec_p2p_route(const b200rec_p2p_route_args *args, void *stream);

typedef struct b200rec_p2p_step_args {
    int32_t world, rank, ld, d;
    float *U_peer[B200REC_MAX_RANKS];  /* user-table shard base of every rank ([rank] = this rank's own)         */
    float *V_peer[B200REC_MAX_RANKS];  /* item-table shard base of every rank                                    */
    int32_t item_bounds[B200REC_MAX_RANKS + 1];
    int32_t head;                      /* replicated head rows [0, head): read from Vh, updates reduced into dVh  */
    float *Vh, *dVh;                   /* [head, ld] local replica and where its updates go: dVh == Vh updates the
                                          replica in place (Hogwild inside the rank; the caller exchanges the
                                          difference to a snapshot between steps), a separate zeroed buffer keeps Vh
                                          at its pre-step value; NULL when head == 0                              */
    /* for every source rank s: the outbox segment s routed to THIS rank (already offset by rank*cap) + its count */
    const int32_t *in_u[B200REC_MAX_RANKS], *in_i[B200REC_MAX_RANKS], *in_j[B200REC_MAX_RANKS];
    const int32_t *in_cnt[B200REC_MAX_RANKS];
    float lr, reg, inv_batch;  /* inv_batch = 1 / GLOBAL batch                                                   */
    int32_t flags;             /* B200REC_F_USERS_UNIQUE: user rows written back with plain stores               */
    double *loss_sum;          /* optional local device scalar                                                   */
    int32_t *n_processed;      /* optional local device int: triples this rank processed                         */
} b200rec_p2p_step_args;

/* fused step over the triples routed to this rank: user row <- home rank (peer load), item rows local (peer only
 * when a given negative lives elsewhere), item updates by vector atomics, user row -> home rank (peer store). */
int b200rec_p2p_step(const b200rec_p2p_step_args *args, void *stream);

/* peer memory: cudaMalloc'd (IPC-exportable) allocation, 64-byte handle, import on another process of the box */
int b200rec_peer_alloc(int64_t bytes, void **dev_ptr);
int b200rec_peer_free(void *dev_ptr);
int b200rec_peer_export(void *dev_ptr, void *handle64);
int b200rec_peer_import(const void *handle64, void **dev_ptr);
int b200rec_peer_close(void *dev_ptr);
/* dst[0..bytes) = src[0..bytes) on `stream`; either side may be a peer-mapped pointer (UVA device-to-device copy) */
int b200rec_peer_copy(void *dst, const void *src, int64_t bytes, void *stream);

/* L2 residency control for a reused table (no reference counterpart; the item table of the BPR step): reserves up
 * to `bytes` of persisting L2 and installs an access-policy window [base, base+bytes) on `stream`; bytes == 0
 * removes it.  Affects kernels launched on `stream` afterwards. */
int b200rec_l2_persist(const void *base, int64_t bytes, float hit_ratio, void *stream);

/* Row-wise (lazy) Adam = torch.optim.SparseAdam semantics: only the rows listed in
 * ids[0..n) move.  `grad` holds the per-row gradient sums produced by b200rec_bpr_step(SINK_GRAD) and is zeroed
 * again row by row; `stamp` (int32 per table row, zero-initialised once) de-duplicates ids within a step; ids < 0
 * are skipped.  `step` is the optimiser's global step count (>= 1, bias correction as in SparseAdam). */
int b200rec_adam_rows(float *W, float *grad, float *exp_avg, float *exp_avg_sq, int32_t *stamp, int ld,
                      const int32_t *ids, int n, float lr, float beta1, float beta2, float eps, int step,
                      void *stream);

/* Scoring algorithms */
#define B200REC_SCORE_EXACT 0 /* fp32 FMA chain in k order on CUDA cores (bit-exact vs oracle) */
#define B200REC_SCORE_TC 1    /* tcgen05 bf16 candidate pass + exact fp32 re-rank               */

/* models/MF.py:109-132 + func.h:12-31 fused: for each of n_users users,
 * S = U[user] . V^T over all items, -inf at the user's mask row (train
 * positives), top-k by (score desc, id asc).  out_idx int32 [n_users,k],
 * out_score fp32 [n_users,k] (may be NULL).  mask CSR is indexed by user id and
 * may be NULL.  workspace: device scratch of at least
 * b200rec_score_topk_workspace() bytes. */
int64_t b200rec_score_topk_workspace(int n_users, int num_items, int d, int k, int algo);
int b200rec_score_topk(const float *U, const float *V, int ld, int d, const int32_t *users,
                       int n_users, int num_items, const int64_t *mask_indptr,
                       const int32_t *mask_indices, int k, int32_t *out_idx, float *out_score,
                       void *workspace, int64_t workspace_bytes, int algo, void *stream);

/* Test hooks (not part of the reference surface): the raw fp16 tensor-core scores of the
 * B200REC_SCORE_TC candidate pass, dense fp32 [roundup(n_users,256), roundup(num_items,128)], items in
 * the kernel's descending-norm order and rescaled domain; debug_tc_layout gives the workspace offsets
 * (relative to the 1024-aligned workspace base) of that permutation and of {scale_v, scale_u}. */
int b200rec_debug_tc_layout(int n_users, int num_items, int d, int64_t *off_perm, int64_t *off_scales);
int b200rec_debug_tc_scores(const float *U, const float *V, int ld, int d, const int32_t *users,
                            int n_users, int num_items, float *dump, void *workspace,
                            int64_t workspace_bytes, void *stream);

/* models/MF.py:109-130 for the dense predict() contract (small U only):
 * out fp32 [n_users, num_items] = U[users] @ V^T with -inf at mask nonzeros. */
int b200rec_predict_dense(const float *U, const float *V, int ld, int d, const int32_t *users,
                          int n_users, int num_items, const int64_t *mask_indptr,
                          const int32_t *mask_indices, float *out, void *stream);

/* device-resident form of c_top_k_array_index (func.h:22-31) */
int b200rec_topk_rows(const float *scores, int64_t row_stride, int rows, int cols, int k,
                      int32_t *out_idx, void *stream);

/* device-resident forms of evaluate_holdout / evaluate_loo.  truth CSR rows are
 * addressed by row_ids[r] (or r when row_ids is NULL); indices need not be sorted.
 * Ks: HOST array.  out: device fp32 [n, 3*K_len] / [n, 2*K_len]. */
int b200rec_holdout_metrics(const int32_t *topk, int n, int max_k, const int32_t *row_ids,
                            const int64_t *truth_indptr, const int32_t *truth_indices,
                            const int *Ks, int K_len, float *out, void *stream);
int b200rec_loo_metrics(const int32_t *topk, int n, int max_k, const int32_t *row_ids,
                        const int64_t *truth_indptr, const int32_t *truth_indices,
                        const int *Ks, int K_len, float *out, void *stream);
/* utils/stats.py:29-32 column means of an fp32 [n, cols] device matrix -> out_host[cols] */
int b200rec_column_means(const float *mat, int64_t n, int cols, double *out_host, void *stream);

/* models/LightGCN.py:196 one propagation layer Y = A X for CSR A (fp32 values),
 * optionally accumulating acc += scale * Y (the running layer mean of :198-200);
 * acc_init != 0 starts the mean instead: acc = scale * X + scale * Y (first layer). */
int b200rec_spmm_csr(const int64_t *indptr, const int32_t *indices, const float *values,
                     int n_rows, const float *X, int ldx, int d, float *Y, int ldy, float *acc,
                     int ldacc, float acc_scale, int acc_init, void *stream);

/* Same product with a plan for LONG rows (popular items have up to ~1e6 neighbours): rows with more than seg_len
 * nonzeros are cut into segments [seg_begin[s], seg_end[s]) of at most seg_len nonzeros (n_seg of them, grouped by
 * row: long_rows[r] owns segments long_seg_ptr[r] .. long_seg_ptr[r+1]); `partial` is scratch of
 * n_seg * roundup(d,4) floats.  Partials are added in segment order: results are deterministic.  n_seg == 0 is
 * b200rec_spmm_csr.  The plan depends on indptr only (recsys_pytorch_b200.engine.spmm_plan builds it once per graph). */
int b200rec_spmm_csr_split(const int64_t *indptr, const int32_t *indices, const float *values, int n_rows,
                           const float *X, int ldx, int d, float *Y, int ldy, float *acc, int ldacc,
                           float acc_scale, int acc_init, int64_t seg_len, const int64_t *seg_begin,
                           const int64_t *seg_end, int n_seg, const int32_t *long_rows,
                           const int32_t *long_seg_ptr, int n_long, float *partial, void *stream);

/* models/NGCF.py:198-212, one propagation layer after side = A_hat @ ego (b200rec_spmm_csr):
 *   z = side W_gc + b_gc + (ego * side) W_bi + b_bi;  ego_next = dropout(leaky_relu(z, 0.2), mess_dropout);
 *   nrm = max(|ego_next|_2, 1e-12);  acc += acc_scale * ego_next / nrm      (the running mean of :216-218)
 * W_*: [d,d] row-major (in x out) as nn.Parameter stores them, b_*: [d].  d <= 64.  mess_dropout > 0 draws its mask
 * from the library's counter RNG keyed by (seed, step, layer, row, column) - pass 0 for evaluation. */
int b200rec_ngcf_layer_forward(const float *ego, const float *side, const float *W_gc, const float *b_gc,
                               const float *W_bi, const float *b_bi, int n_rows, int ld, int d, int layer,
                               float mess_dropout, uint64_t seed, uint64_t step, float *ego_next, float *nrm,
                               float *acc, float acc_scale, void *stream);
/* autograd backward of that layer.  g_out = dL/d(acc) (dense [n_rows, ld]); g_next = gradient reaching ego_next from
 * the layer above (NULL for the top layer).  Writes g_z (scratch), g_side = d(side), g_ego = the direct part of d(ego)
 * (the caller adds A_hat @ g_side with b200rec_spmm_csr), and ACCUMULATES dW_gc, dW_bi [d,d] and db [d] (= d b_gc =
 * d b_bi; caller zeroes them). */
int b200rec_ngcf_layer_backward(const float *g_out, const float *g_next, const float *ego, const float *side,
                                const float *ego_next, const float *nrm, const float *W_gc, const float *W_bi,
                                int n_rows, int ld, int d, int layer, float mess_dropout, uint64_t seed, uint64_t step,
                                float acc_scale, float *g_z, float *g_side, float *g_ego, float *dW_gc, float *dW_bi,
                                float *db, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200REC_H */
