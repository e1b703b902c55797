/*
 * gmr.h -- C ABI of libgmr.so, the sm_100a implementation of GenMMRec's hot path.
 *
 * The reference (orangeai-research/Generative-Multimodal-Recommendation) is pure Python and has no
 * FFI of its own; on this path it calls four PyTorch library operators.  Each entry point below
 * replaces one of those call sites (cited per function, paths relative to the reference root) and
 * is what a ctypes stub added to the reference would bind (INTEGRATION.md shows that stub).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the boundary.
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls only enqueue
 *     work on that stream; they never synchronise unless documented.
 *   - return value: 0 on success, a negative GMR_ERR_* code otherwise; gmr_last_error() returns a
 *     thread-local description of the last failure.  No C++ exception crosses the ABI.
 *   - the caller owns every buffer, including workspaces (sizes are queried with the matching
 *     *_workspace_bytes call); plan handles own only their private schedule arrays.
 *   - indices are int32 (N, nnz < 2^31), values/embeddings fp32, row-major with explicit leading
 *     dimensions in ELEMENTS.
 */
#ifndef GMR_H_
#define GMR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMR_OK 0
#define GMR_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, unsupported K ...) */
#define GMR_ERR_CUDA (-2)        /* a CUDA runtime call failed; see gmr_last_error() */
#define GMR_ERR_UNSUPPORTED (-3) /* shape/mode not implemented by this build */
#define GMR_ERR_WORKSPACE (-4)   /* workspace too small */

#define GMR_ABI_VERSION 1

/* Largest K the fused top-K kernels accept. */
#define GMR_MAX_TOPK 256

const char* gmr_last_error(void);
int gmr_abi_version(void);
/* SM count, compute capability and L2 size of the current device. */
int gmr_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* l2_bytes);

/* ---------------------------------------------------------------------------------------------
 * K1  CSR SpMM:  Y = alpha * A * X + beta * Y
 * replaces  torch.sparse.mm / torch.spmm(A_coo, X)
 *   GenMMRec/src/models/diffmm.py:136,139,142,146,149,152,284-285   (forward_MM, GCNLayer)
 *   GenMMRec/src/models/gume.py:216,226,241,247                      (conv_ui, conv_ii, R)
 *   GenMMRec/src/models/genrecv1.py:259,279,284,289,294              (user_item_GCN, item_item_GCN)
 *   GenMMRec/src/models/lightgcn.py:122   GenMMRec/src/models/ld4mrec.py:206
 *
 * A is CSR (rowptr int32[n_rows+1], col int32[nnz], val fp32[nnz]); duplicates allowed (they add,
 * as in an uncoalesced COO).  X is [n_cols, D] with leading dimension ldx, Y is [n_rows, D] with
 * ldy.  beta == 0 means Y is write-only (it may hold NaNs).  Every D >= 1 is accepted; D % 4 == 0
 * with 16-byte aligned X/Y rows takes the 128-bit gather path.
 *
 * A plan holds the row schedule of one sparsity pattern (rows cut into chunks of at most
 * `chunk_nnz` nonzeros so that one warp never owns a long row; chunk partial sums of split rows go
 * through the workspace and are reduced in a fixed order -> results are run-to-run deterministic).
 * Plan creation reads rowptr back to the host and synchronises `stream`; do it once per graph.
 * ------------------------------------------------------------------------------------------- */
typedef struct gmr_spmm_plan gmr_spmm_plan_t;

int gmr_spmm_plan_create(gmr_spmm_plan_t** plan, const int32_t* rowptr, int64_t n_rows, int64_t n_cols,
                         int32_t chunk_nnz /* 0 = default */, void* stream);
int gmr_spmm_plan_destroy(gmr_spmm_plan_t* plan);
/* number of virtual rows (chunks) and of split (multi-chunk) rows, for diagnostics */
int gmr_spmm_plan_stats(const gmr_spmm_plan_t* plan, int64_t* n_chunks, int64_t* n_split_rows, int64_t* nnz);
int64_t gmr_spmm_workspace_bytes(const gmr_spmm_plan_t* plan, int32_t D);

int gmr_spmm_csr_f32(const gmr_spmm_plan_t* plan, const int32_t* rowptr, const int32_t* col, const float* val,
                     const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t D, float alpha, float beta,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1b  column-blocked, nonzero-centric SpMM (same operator and call sites as K1) for graphs whose
 * gathered operand does not fit the L2.  The blocked plan SNAPSHOTS the matrix: it stably sorts the
 * nonzeros by column block (blocks of `block_cols` columns; 0 = one block) and keeps its own
 * (row, col, val) arrays, so the product takes no CSR arguments.  One pass per block runs back to
 * back on the stream (each gathers from an L2-resident slice of X), Y is accumulated across
 * passes.  Per-row summation order is canonical (pieces of 64 nonzeros counted from the start of a
 * row's block segment, blocks ascending): results are run-to-run deterministic and do not depend
 * on which other rows share the matrix (row sharding keeps bits).  D % 4 == 0, 16-byte aligned
 * X / Y rows.  gmr_spmm_blocked_plan_set_values re-reads `val` (same sparsity pattern, new values).
 * stats[8] = {blocks, block_cols, tiles, segments, long segments, slots, small / big reduce entries}.
 * ------------------------------------------------------------------------------------------- */
typedef struct gmr_spmm_bplan gmr_spmm_bplan_t;

int gmr_spmm_blocked_plan_create(gmr_spmm_bplan_t** plan, const int32_t* rowptr, const int32_t* col, const float* val,
                                 int64_t n_rows, int64_t n_cols, int64_t block_cols, void* stream);
int gmr_spmm_blocked_plan_set_values(gmr_spmm_bplan_t* plan, const float* val, void* stream);
int gmr_spmm_blocked_plan_destroy(gmr_spmm_bplan_t* plan);
int gmr_spmm_blocked_plan_stats(const gmr_spmm_bplan_t* plan, int64_t* stats /* [8] */);
int64_t gmr_spmm_blocked_workspace_bytes(const gmr_spmm_bplan_t* plan, int32_t D);
int gmr_spmm_blocked_f32(const gmr_spmm_bplan_t* plan, const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t D,
                         float alpha, float beta, void* workspace, int64_t workspace_bytes, void* stream);

/* Same product, with the result rows additionally PUSHED to peer buffers (fused all-gather of the
 * row-sharded layer output over NVLink peer memory): rank-local rows [0, n_rows) of this shard are
 * written to y_peers[p] + (row_offset + r) * ldy for every p in [0, n_peers).  y_peers is a
 * DEVICE array of n_peers device pointers (peer-mapped); the caller synchronises the ranks. */
int gmr_spmm_csr_f32_push(const gmr_spmm_plan_t* plan, const int32_t* rowptr, const int32_t* col, const float* val,
                          const float* X, int64_t ldx, float* const* y_peers, int32_t n_peers, int64_t row_offset,
                          int64_t ldy, int32_t D, float alpha, void* workspace, int64_t workspace_bytes,
                          void* stream);

/* Dense all-gather of a row block by peer stores: rows [0, n_rows) of `src` (leading dimension ld) are copied to
 * y_peers[p] + (row_offset + r) * ldy for every p (own replica included).  D, ld, ldy multiples of 4, 16-byte
 * aligned buffers.  The caller synchronises the ranks (a stream-ordered barrier) before anyone reads. */
int gmr_rows_push_f32(const float* src, int64_t ld, int64_t n_rows, int32_t D, float* const* y_peers, int32_t n_peers,
                      int64_t row_offset, int64_t ldy, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  fused score + train-history mask + top-K (the [B, I] score matrix is never written)
 * replaces  torch.matmul(u_e[user], i_e.T)  ->  scores[mask] = -1e10  ->  torch.topk(scores, K)
 *   GenMMRec/src/models/diffmm.py:276-278, gume.py:420-428, genrecv1.py:417-427, vbpr.py:100-106,
 *   lightgcn.py:156-164, ld4mrec.py:36,54 (output_proj: bias != NULL)
 *   GenMMRec/src/common/trainer.py:381-386
 *
 * score(b, i) = bias[i] + sum_d Eu[users[b], d] * Ei[i, d]  evaluated as the fp32 chain
 * s = bias[i]; for d = 0..D-1: s = fmaf(u[d], e[d], s)   (bias NULL -> s starts at +0.0f).
 * Masked pairs score exactly -1e10f.  Results are ordered by (score descending, item id
 * ascending) -- a total order, so the output is unique.
 *
 * users:        int64[B] row ids into Eu, or NULL for rows 0..B-1
 * mask_rowptr:  int64[B+1] (or NULL: no mask), mask_items int32, ascending within each row
 * out_ids:      int32[B, K]   out_scores: fp32[B, K] or NULL
 * precision:    GMR_SCORE_FP32      CUDA-core fp32 FMA scoring
 *               GMR_SCORE_TC        one fp16 tcgen05 pass used as a certified screen (error bound
 *                                   2^-10 |u| max|e|), then exact fp32 re-scoring of every item that
 *                                   can still reach the top K; rows the screen cannot decide (buffer
 *                                   overflow under massive near-ties, fewer than K unmasked items,
 *                                   non-finite scores) are redone on the fp32 path.  Same ids and
 *                                   scores as GMR_SCORE_FP32.  Requires D % 64 == 0, D <= 256.
 *               GMR_SCORE_TC_SPLIT  first-generation path: 3-term split-bf16 tcgen05 product + exact
 *                                   re-scoring of the K' best candidates, certified per row.  Same
 *                                   results; kept for cross-checks.  D % 64 == 0, D <= 192, K <= 248.
 * 1 <= K <= GMR_MAX_TOPK; if fewer than K items exist the tail is (-1, -inf).
 * ------------------------------------------------------------------------------------------- */
#define GMR_SCORE_FP32 0
#define GMR_SCORE_TC 1
#define GMR_SCORE_TC_SPLIT 2

int64_t gmr_score_topk_workspace_bytes(int32_t B, int32_t I, int32_t D, int32_t K, int32_t precision);

int gmr_score_mask_topk_f32(const float* Eu, int64_t lde_u, const int64_t* users, int32_t B, const float* Ei,
                            int64_t lde_i, const float* bias, int32_t I, int32_t D, const int64_t* mask_rowptr,
                            const int32_t* mask_items, int32_t K, int32_t precision, int32_t* out_ids,
                            float* out_scores, void* workspace, int64_t workspace_bytes, void* stream);

/* Diagnostics of the last tensor-core call that used `workspace` with the same shape and precision:
 * number of rows that were redone on the fp32 path.  Copies one int to the host and synchronises
 * `stream`. */
int gmr_score_tc_fallback_rows(const void* workspace, int32_t B, int32_t I, int32_t D, int32_t K, int32_t precision,
                               int32_t* count_host, void* stream);
/* Counters of the last GMR_SCORE_TC call made with the environment variable GMR_SCREEN_STATS set
 * (zeros otherwise): stats_host[8] = {chunks that took the append path, appended candidates, cheap
 * prunes, exact prunes, -, exactly re-scored candidates, -, -}.  Synchronises `stream`. */
int gmr_score_tc_stats(const void* workspace, int32_t B, int32_t I, int32_t D, int32_t K, uint64_t* stats_host,
                       void* stream);

/* Dense variant for API parity with full_sort_predict (GenMMRec/src/models/diffmm.py:277): writes the
 * fp32 [B, I] score matrix (leading dimension ldo) with the same fmaf chain.  Not used by the fused
 * evaluation. */
int gmr_scores_f32(const float* Eu, int64_t lde_u, const int64_t* users, int32_t B, const float* Ei, int64_t lde_i,
                   const float* bias, int32_t I, int32_t D, float* out, int64_t ldo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4  hit matrix + Recall / NDCG / Precision / MAP prefix sums
 * replaces  the Python membership loop and the numpy metric kernels
 *   GenMMRec/src/utils/topk_evaluator.py:107-120,299-313   GenMMRec/src/utils/metrics.py:12-105
 *
 * topk int32[U, K]; ground truth as CSR over eval users (gt_rowptr int64[U+1], gt_items int32
 * ascending within a row, length >= 1 per user).  Writes hit uint8[U, K] (or NULL) and
 * sums double[4, K] = per-position SUMS over users of recall, ndcg, precision, map (the caller
 * divides by the user count, after an all-reduce when users are sharded).  Deterministic.
 * ------------------------------------------------------------------------------------------- */
int64_t gmr_hits_metrics_workspace_bytes(int32_t U, int32_t K);
int gmr_hits_metrics(const int32_t* topk, const int64_t* gt_rowptr, const int32_t* gt_items, int32_t U, int32_t K,
                     uint8_t* hit, double* sums, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row-wise glue of the propagation step (a7 in SURVEY.md section 8): fused so that a layer costs
 * one pass over [N, D] instead of ~10 (F.normalize / axpy / layer sums of
 * GenMMRec/src/models/diffmm.py:138-167).
 *   out[r, :] = a * x[r, :] + b * y[r, :] + c * z[r, :] / max(||z[r, :]||_2, eps)
 * y and/or z may be NULL (their terms vanish).  out may alias x or y.
 * ------------------------------------------------------------------------------------------- */
int gmr_rows_axpby_norm_f32(const float* x, int64_t ldx, const float* y, int64_t ldy, const float* z, int64_t ldz,
                            float* out, int64_t ldo, int64_t n_rows, int32_t D, float a, float b, float c,
                            float eps, void* stream);

/* Modality mix of the propagation step (GenMMRec/src/models/diffmm.py:131-133,138,145: leaky-ReLU'd
 * projections, F.normalize, softmax-weighted sum) in one pass:
 *   out[r, 0:D]  = w1 * n(act(x1[r])) + w2 * n(act(x2[r])),   n(v) = v / max(||v||_2, eps),
 *   out[r, D:2D] = out[r, 0:D] + y[r]                         (only when y != NULL; then ldo >= 2 D)
 * act = leaky ReLU with negative slope `slope` (1.0 = identity).  D <= 256. */
int gmr_rows_normalize_mix_f32(const float* x1, int64_t ld1, const float* x2, int64_t ld2, const float* y, int64_t ldy,
                               float* out, int64_t ldo, int64_t n_rows, int32_t D, float w1, float w2, float slope,
                               float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense modality projection  C[M, N] = A[M, K] * W[K, N]  (fp32 in, fp32 out, fp32-level accuracy)
 * replaces  torch.mm(image_embedding.weight, image_trans) / torch.mm(text_embedding.weight, text_trans)
 *   GenMMRec/src/models/diffmm.py:115-127 (getImageFeats / getTextFeats, before the leaky ReLU)
 *
 * tcgen05.mma.kind::tf32 with both operands split as hi + lo (4 product terms) and two-level accumulation
 * (TMEM per 128-column K chunk, fp32 registers across chunks); TMA-fed, persistent, HBM-bound on reading A once.
 * Requires N == 64, K % 32 == 0, 16-byte aligned rows of A and C.  GMR_ERR_UNSUPPORTED otherwise (the host layer
 * then falls back to the library GEMM, which is what the reference calls).
 * ------------------------------------------------------------------------------------------- */
int64_t gmr_dense_proj_workspace_bytes(int32_t K, int32_t N);
int gmr_dense_proj_f32(const float* A, int64_t lda, int32_t M, int32_t K, const float* W, int64_t ldw, int32_t N, float* C,
                       int64_t ldc, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-side fusion: the BPR gathers of calculate_loss
 * replaces  anc = usr[users]; pos = itm[pos_items]; neg = itm[neg_items];
 *           (anc * pos).sum(-1), (anc * neg).sum(-1)        and their index_put_(accumulate) backward
 *   GenMMRec/src/models/diffmm.py:211-219, vbpr.py:84-93, gume.py:281-292, genrecv1.py:362-372, lightgcn.py:136-146
 *
 * pos_score[t] = <Eu[users[t]], Ei[pos[t]]>,  neg_score[t] = <Eu[users[t]], Ei[neg[t]]>   (fp32 fmaf chains).
 * The backward ACCUMULATES (+=) into dEu / dEi:
 *   dEu[users[t]] += g_pos[t] Ei[pos[t]] + g_neg[t] Ei[neg[t]];  dEi[pos[t]] += g_pos[t] Eu[users[t]];
 *   dEi[neg[t]] += g_neg[t] Eu[users[t]]
 * with fp32 atomic adds (repeated indices inside a batch), i.e. reproducible to summation-order rounding.
 * ------------------------------------------------------------------------------------------- */
int gmr_bpr_scores_f32(const float* Eu, int64_t ldu, const float* Ei, int64_t ldi, const int64_t* users, const int64_t* pos,
                       const int64_t* neg, int32_t B, int32_t D, float* pos_score, float* neg_score, void* stream);
int gmr_bpr_scores_backward_f32(const float* Eu, int64_t ldu, const float* Ei, int64_t ldi, const int64_t* users,
                                const int64_t* pos, const int64_t* neg, int32_t B, int32_t D, const float* g_pos,
                                const float* g_neg, float* dEu, int64_t lddu, float* dEi, int64_t lddi, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Column-sharded propagation (the multi-GPU form of the same torch.sparse.mm call sites: rank g owns
 * columns [g * dc, (g + 1) * dc) of every embedding row; A X is column-separable, so the SpMM layers
 * need no collective).
 *
 * gmr_spmm_narrow_f32: K1 for rows of D = 8 / 16 / 32 / 64 floats.  A warp walks 32 / (D/4 * chains)
 * virtual rows at once (taken in descending-length order, gmr_spmm_plan_enable_narrow builds that
 * order once per plan and synchronises `stream`).  Per-column operation order is the wide
 * kernels': chains = 2 -> even / odd nonzeros in two sequential fmaf chains added at the end (what
 * gmr_spmm_csr_f32 does for D <= 64), chains = 1 -> one chain (what it does for 64 < D <= 128).
 * A column slice of the narrow product therefore equals the wide product's columns bit for bit.
 * X / Y rows 16-byte aligned, ldx / ldy multiples of 4; workspace as gmr_spmm_workspace_bytes(plan, D).
 * ------------------------------------------------------------------------------------------- */
int gmr_spmm_plan_enable_narrow(gmr_spmm_plan_t* plan, const int32_t* rowptr, void* stream);
int gmr_spmm_narrow_f32(const gmr_spmm_plan_t* plan, const int32_t* rowptr, const int32_t* col, const float* val,
                        const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t D, int32_t chains, float alpha,
                        float beta, void* workspace, int64_t workspace_bytes, void* stream);

/* Sliced-ELL snapshot for the narrow product: the (col, val) pairs of the 32 / (D/4 * chains) rows a warp walks
 * together, interleaved per iteration of 8 entries, rows in descending-length order -- one coalesced load per lane
 * and iteration instead of scattered CSR reads (the gathers of a column shard are bound by the L1 request rate, so
 * every request the index stream does not take is a gather gained).  The snapshot copies the matrix (col, val;
 * gmr_spmm_sell_set_values re-reads them after an in-place value change), is tied to one (D, chains) lane layout
 * and keeps a pointer to `plan`, which must outlive it.  Same results, bit for bit, as gmr_spmm_narrow_f32. */
typedef struct gmr_spmm_sell gmr_spmm_sell_t;
int gmr_spmm_sell_create(gmr_spmm_sell_t** sell, gmr_spmm_plan_t* plan, const int32_t* rowptr, const int32_t* col,
                         const float* val, int32_t D, int32_t chains, void* stream);
int gmr_spmm_sell_set_values(gmr_spmm_sell_t* sell, const int32_t* rowptr, const int32_t* col, const float* val, void* stream);
int gmr_spmm_sell_destroy(gmr_spmm_sell_t* sell);
int64_t gmr_spmm_sell_bytes(const gmr_spmm_sell_t* sell);
int gmr_spmm_sell_f32(const gmr_spmm_sell_t* sell, const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t D,
                      int32_t chains, float alpha, float beta, void* workspace, int64_t workspace_bytes, void* stream);

/* Layout changes between row blocks and column blocks by peer stores.  For every peer p (y_peers is a DEVICE
 * array of n_peers <= 16 peer-mapped base pointers), every source row r and every segment s < n_seg:
 *   y_peers[p][(dst_row_offset + r - first_p) * ldy + dst_col0 + s * dst_seg_step + c]
 *       = src[r * ld + src_col0 + s * src_seg_step + p * src_col_step + c],  c < dc
 * row_bounds == NULL: every row goes to every peer (first_p = 0) -- an all-to-all of column slices when
 * src_col_step = dc, an all-gather of one slice when it is 0.  row_bounds = DEVICE int64[n_peers + 1]: row r goes
 * only to the peer with row_bounds[p] <= r < row_bounds[p + 1] (first_p = row_bounds[p]).  All widths and
 * offsets multiples of 4 floats, 16-byte aligned buffers.  The caller synchronises the ranks. */
int gmr_cols_push_f32(const float* src, int64_t ld, int64_t n_rows, int32_t dc, int64_t src_col0, int64_t src_col_step,
                      float* const* y_peers, int32_t n_peers, const int64_t* row_bounds, int64_t dst_row_offset,
                      int64_t ldy, int64_t dst_col0, int32_t n_seg, int64_t src_seg_step, int64_t dst_seg_step, void* stream);

/* out[r, p * dc + c] = slabs[p * slab_stride + r * dc + c]: n_slabs column slabs of [n_rows, dc] (what the peers stored,
 * each contiguously -- NVLink moves full lines ~3x faster than 32-byte granules at a row pitch) side by side as rows. */
int gmr_slabs_to_rows_f32(const float* slabs, int32_t n_slabs, int64_t slab_stride, int64_t n_rows, int32_t dc, float* out,
                          int64_t ldo, void* stream);

/* out[r] = sum_d z[r, d]^2 over D = 4 / 8 / 16 / 32 / 64 columns: a rank's part of the squared row norm of
 * F.normalize (GenMMRec/src/models/diffmm.py:166), summed as the subtree of gmr_rows_axpby_norm_f32's (D = 64)
 * reduction tree that covers an aligned block of D columns. */
int gmr_rows_sumsq_f32(const float* z, int64_t ldz, int64_t n_rows, int32_t D, float* out, void* stream);

/* out[r, :] = a * x[r, :] + b * y[r, :] + c * z[r, :] / max(sqrt(ss[r]), eps),  ss[r] = the n_parts (a power of two
 * <= 16) values ss_parts[k * ss_stride + r] added as a balanced tree, neighbours first -- together with
 * gmr_rows_sumsq_f32 the same bits as gmr_rows_axpby_norm_f32 on the full 64-column row.  y, z may be NULL;
 * out may alias x, y or z; D % 4 == 0, 16-byte aligned rows. */
int gmr_rows_axpby_ss_f32(const float* x, int64_t ldx, const float* y, int64_t ldy, const float* z, int64_t ldz,
                          const float* ss_parts, int32_t n_parts, int64_t ss_stride, float* out, int64_t ldo,
                          int64_t n_rows, int32_t D, float a, float b, float c, float eps, void* stream);

/* Stream-ordered cross-rank barrier through peer memory: flags_peers is a DEVICE array of n_ranks pointers to
 * every rank's uint32[n_ranks] flag array (zero-initialised, peer-mapped), state a LOCAL device uint32[2]
 * (zero-initialised): state[0] counts the barriers passed, state[1] the times a peer did not arrive within ~4 s
 * (the kernel then gives up instead of hanging the GPU; the host should check it).  Everything this rank's earlier
 * kernels on `stream` stored into peer buffers is visible to a peer once it has passed the barrier. */
int gmr_peer_barrier(uint32_t* const* flags_peers, int32_t my_rank, int32_t n_ranks, uint32_t* state, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Peer-memory plumbing for the fused SpMM + all-gather (CUDA IPC, one process per GPU).
 * gmr_peer_alloc allocates `bytes` of device memory suitable for export; gmr_peer_export fills a
 * 64-byte handle; another process maps it with gmr_peer_open.  The Python host exchanges the
 * handles through torch.distributed.
 * ------------------------------------------------------------------------------------------- */
#define GMR_PEER_HANDLE_BYTES 64
int gmr_peer_alloc(void** ptr, int64_t bytes);
int gmr_peer_free(void* ptr);
int gmr_peer_export(void* ptr, uint8_t* handle_host /* [GMR_PEER_HANDLE_BYTES] */);
int gmr_peer_open(const uint8_t* handle_host, void** ptr);
int gmr_peer_close(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* GMR_H_ */
