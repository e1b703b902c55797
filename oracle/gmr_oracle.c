/*
 * gmr_oracle.c -- CPU restatement (plain C) of the arithmetic on GenMMRec's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under generative-multimodal-recommendation_b200/ links,
 * loads or calls this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may (as the checker / the timed CPU baseline, never as the product).
 *
 * The reference is pure Python; its arithmetic on this path is four PyTorch library calls.
 * Each function below restates one of them at the call site named in its comment, with a
 * FIXED, documented evaluation order so that integer/index results are reproducible bit for bit:
 *
 *   oracle_spmm_*            torch.sparse.mm / torch.spmm(A_coo, X)
 *                            GenMMRec/src/models/diffmm.py:136-152,284-285; gume.py:211-227,241,247;
 *                            genrecv1.py:255-306; lightgcn.py:121-123; ld4mrec.py:206
 *   oracle_score_mask_topk   torch.matmul(u_e, i_e.T) -> scores[mask] = -1e10 -> torch.topk
 *                            GenMMRec/src/models/diffmm.py:276-278 (and the five sibling
 *                            full_sort_predict bodies), GenMMRec/src/common/trainer.py:379-387
 *   oracle_hits              [[i in m for i in n] ...]   GenMMRec/src/utils/topk_evaluator.py:107-112
 *   oracle_metrics           recall_/ndcg_/precision_/map_  GenMMRec/src/utils/metrics.py:12-105,
 *                            mean over users as in topk_evaluator.py:299-313
 *
 * Parity pin: tests/test_oracle_golden.py checks every function here against vectors produced by
 * the reference itself (tests/golden/*.npz, generator tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- SpMM ---------------------------------------------------------------------------------- */

/* COO, entries visited in storage order, fp32 accumulate: the order ATen's CPU kernel uses for
 * an uncoalesced COO operand (one axpy per stored entry; duplicates simply add up). */
void oracle_spmm_coo_f32(const int64_t* rows, const int64_t* cols, const float* vals, int64_t nnz,
                         const float* X, int64_t ldx, int32_t D, float* Y, int64_t ldy, int64_t n_rows)
{
    for (int64_t r = 0; r < n_rows; ++r) memset(Y + r * ldy, 0, sizeof(float) * (size_t)D);
    for (int64_t e = 0; e < nnz; ++e) {
        const float v = vals[e];
        const float* x = X + cols[e] * ldx;
        float* y = Y + rows[e] * ldy;
        for (int32_t d = 0; d < D; ++d) y[d] += v * x[d];
    }
}

/* CSR, fp32 accumulate in column-storage order per row; Y = alpha*A*X + beta*Y. */
void oracle_spmm_csr_f32(const int32_t* rowptr, const int32_t* col, const float* val, const float* X,
                         int64_t ldx, int32_t D, float* Y, int64_t ldy, int64_t n_rows, float alpha,
                         float beta)
{
    float* acc = (float*)malloc(sizeof(float) * (size_t)D);
    for (int64_t r = 0; r < n_rows; ++r) {
        for (int32_t d = 0; d < D; ++d) acc[d] = 0.0f;
        for (int32_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
            const float v = val[e];
            const float* x = X + (int64_t)col[e] * ldx;
            for (int32_t d = 0; d < D; ++d) acc[d] += v * x[d];
        }
        float* y = Y + r * ldy;
        for (int32_t d = 0; d < D; ++d) y[d] = (beta == 0.0f) ? alpha * acc[d] : alpha * acc[d] + beta * y[d];
    }
    free(acc);
}

/* Same product with fp64 accumulation: the yardstick both fp32 orders are measured against. */
void oracle_spmm_csr_f64(const int32_t* rowptr, const int32_t* col, const float* val, const float* X,
                         int64_t ldx, int32_t D, double* Y, int64_t ldy, int64_t n_rows)
{
    for (int64_t r = 0; r < n_rows; ++r) {
        double* y = Y + r * ldy;
        for (int32_t d = 0; d < D; ++d) y[d] = 0.0;
        for (int32_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
            const double v = (double)val[e];
            const float* x = X + (int64_t)col[e] * ldx;
            for (int32_t d = 0; d < D; ++d) y[d] += v * (double)x[d];
        }
    }
}

/* ---- score + mask + top-K ------------------------------------------------------------------- */

/* The scoring order fixed for this project: s = bias; for d = 0..D-1: s = fmaf(u[d], e[d], s)
 * (bias absent -> s starts at +0.0f).  cuBLAS / MKL use other orders; the difference is the
 * tie tolerance stated in the parity tests (|ds| <= 2^-20 * sum|u_d e_d|). */
static inline float score_one(const float* u, const float* e, int32_t D, float b)
{
    float s = b;
    for (int32_t d = 0; d < D; ++d) s = fmaf(u[d], e[d], s);
    return s;
}

/* total order: higher score first, then lower item id */
static inline int better(float sa, int32_t ia, float sb, int32_t ib)
{
    return (sa > sb) || (sa == sb && ia < ib);
}

/*
 * Eu [*, D] row-major (lde_u), users[B] row ids into Eu (NULL -> identity), Ei [I, D] (lde_i),
 * bias[I] or NULL, mask CSR over the B batch rows (rowptr int64[B+1], items int32, any order,
 * duplicates allowed).  Masked entries score -1e10 exactly as trainer.py:384 writes them.
 * out_ids [B, K] int32, out_scores [B, K] fp32, sorted by the total order above.
 */
void oracle_score_mask_topk_f32(const float* Eu, int64_t lde_u, const int64_t* users, int32_t B,
                                const float* Ei, int64_t lde_i, const float* bias, int32_t I, int32_t D,
                                const int64_t* mask_rowptr, const int32_t* mask_items, int32_t K,
                                int32_t* out_ids, float* out_scores)
{
    float* s = (float*)malloc(sizeof(float) * (size_t)I);
    for (int32_t b = 0; b < B; ++b) {
        const float* u = Eu + (users ? users[b] : (int64_t)b) * lde_u;
        for (int32_t i = 0; i < I; ++i) s[i] = score_one(u, Ei + (int64_t)i * lde_i, D, bias ? bias[i] : 0.0f);
        if (mask_rowptr)
            for (int64_t e = mask_rowptr[b]; e < mask_rowptr[b + 1]; ++e) s[mask_items[e]] = -1e10f;
        /* insertion into a sorted list of K: O(I*K) worst case, fine at test sizes */
        int32_t n = 0;
        int32_t* ids = out_ids + (int64_t)b * K;
        float* sc = out_scores + (int64_t)b * K;
        for (int32_t i = 0; i < I; ++i) {
            if (n == K && !better(s[i], i, sc[K - 1], ids[K - 1])) continue;
            int32_t p = (n < K) ? n : K - 1;
            while (p > 0 && better(s[i], i, sc[p - 1], ids[p - 1])) {
                sc[p] = sc[p - 1];
                ids[p] = ids[p - 1];
                --p;
            }
            sc[p] = s[i];
            ids[p] = i;
            if (n < K) ++n;
        }
        for (int32_t p = n; p < K; ++p) { ids[p] = -1; sc[p] = -INFINITY; }
    }
    free(s);
}

/* ---- hit matrix + metrics ------------------------------------------------------------------- */

/* hit[u, j] = topk[u, j] in gt_items[gt_rowptr[u] : gt_rowptr[u+1]]  (linear scan, any order) */
void oracle_hits(const int32_t* topk, const int64_t* gt_rowptr, const int32_t* gt_items, int32_t U,
                 int32_t K, uint8_t* hit)
{
    for (int32_t u = 0; u < U; ++u)
        for (int32_t j = 0; j < K; ++j) {
            const int32_t id = topk[(int64_t)u * K + j];
            uint8_t h = 0;
            for (int64_t e = gt_rowptr[u]; e < gt_rowptr[u + 1]; ++e)
                if (gt_items[e] == id) { h = 1; break; }
            hit[(int64_t)u * K + j] = h;
        }
}

/*
 * Per-position means over users (length K each), float64 throughout as numpy does:
 *   recall[k]    = mean_u c[u,k] / n_u                                 metrics.py:12-15
 *   precision[k] = mean_u c[u,k] / (k+1)                               metrics.py:92-105
 *   ndcg[k]      = mean_u DCG_u(k) / IDCG_u(k), IDCG cut at min(n_u,K) metrics.py:30-63
 *   map[k]       = mean_u [sum_{j<=k} hit*c/(j+1)] / min(k+1, n_u)     metrics.py:66-89
 * with c[u,k] = sum_{j<=k} hit[u,j], n_u = gt_len[u].  Users are summed in index order.
 */
void oracle_metrics(const uint8_t* hit, const int64_t* gt_len, int32_t U, int32_t K, double* recall,
                    double* ndcg, double* precision, double* map)
{
    for (int32_t k = 0; k < K; ++k) recall[k] = ndcg[k] = precision[k] = map[k] = 0.0;
    for (int32_t u = 0; u < U; ++u) {
        const double n = (double)gt_len[u];
        const int64_t lim = gt_len[u] < K ? gt_len[u] : K;
        double c = 0.0, dcg = 0.0, idcg = 0.0, sp = 0.0;
        for (int32_t k = 0; k < K; ++k) {
            const double h = (double)hit[(int64_t)u * K + k];
            const double disc = 1.0 / log2((double)k + 2.0);
            c += h;
            dcg += h * disc;
            if (k < lim) idcg += disc;
            sp += h * (c / (double)(k + 1));
            recall[k] += c / n;
            precision[k] += c / (double)(k + 1);
            /* gt_len == 0 (never produced by the reference's loaders): metrics.py:54-55,86-87 index with -1 and wrap to
             * the full-length normalisers, so NDCG and MAP are 0 there while recall is 0 / 0 = NaN */
            ndcg[k] += (lim > 0) ? dcg / idcg : 0.0;
            map[k] += (lim > 0) ? sp / (double)((k + 1) < lim ? (k + 1) : lim) : 0.0;
        }
    }
    for (int32_t k = 0; k < K; ++k) {
        recall[k] /= U; ndcg[k] /= U; precision[k] /= U; map[k] /= U;
    }
}
