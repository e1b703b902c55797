// score_topk_simt.cu -- K2 (fp32 path): fused score + train-history mask + top-K on CUDA cores.
//
// Replaces  torch.matmul(u_e[user], i_e.T) -> scores[mask] = -1e10 -> torch.topk(scores, K)
// (GenMMRec/src/models/diffmm.py:276-278 and siblings; GenMMRec/src/common/trainer.py:381-386)
// without ever writing the [B, I] score matrix.
//
// A persistent CTA owns a tile of 128 users and sweeps all items in tiles of 128; the 128x128 fp32
// score tile lives in registers (8x8 per thread) and is computed as the chain
// s = bias; s = fmaf(u[d], e[d], s), d ascending -- the order the oracle fixes, so ids AND scores
// are bit-identical to oracle_score_mask_topk_f32.  Scores that beat the row's running threshold
// are appended (after the mask lookup) to the row's key slots; see topk_select.cuh.
//
// This is also the exact fallback of the tcgen05 path (score_topk_tc.cu) for rows whose candidate
// margin cannot certify exactness.
#include "common.cuh"
#include "topk_select.cuh"

namespace gmr {

constexpr int kBU = 128;   // users per CTA tile
constexpr int kBI = 128;   // items per tile
constexpr int kDK = 16;    // reduction chunk
constexpr int kLd = 132;   // padded shared-memory row (floats), 16-byte multiple
constexpr int kThreads = 256;

struct ScoreArgs {
    const float* Eu;
    int64_t lde_u;
    const int64_t* users;   // or null
    const int32_t* row_map; // or null: tile row -> batch row (exact re-run of selected rows)
    int32_t B;              // rows processed by this launch (length of row_map when given)
    const int32_t* B_dev;   // or null: the row count lives on the device (rows queued by the tensor-core path)
    const float* Ei;
    int64_t lde_i;
    const float* bias;
    int32_t I, D;
    const int64_t* mask_rowptr;
    const int32_t* mask_items;
    int32_t K;
    int32_t* out_ids;
    float* out_scores;
    uint64_t* slots;  // [gridDim.x][kBU][CAP]
};

template <int NPL, bool VEC4>
__global__ void __launch_bounds__(kThreads, 2) score_topk_simt_kernel(ScoreArgs a)
{
    constexpr int CAP = 32 * NPL;
    constexpr int KP = CAP / 4;
    constexpr int TRIG = KP + (CAP - KP) / 2;

    __shared__ __align__(16) float As[2][kDK][kLd];
    __shared__ __align__(16) float Bs[2][kDK][kLd];
    __shared__ RowState rows[kBU];
    __shared__ int32_t row_b[kBU];  // batch row of each tile row (-1: padding)

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int tx = t & 15, ty = t >> 4;
    const int n_rows = a.B_dev ? min(*a.B_dev, a.B) : a.B;
    const int n_utiles = (n_rows + kBU - 1) / kBU;
    const int n_itiles = (a.I + kBI - 1) / kBI;
    const int n_chunks = (a.D + kDK - 1) / kDK;
    uint64_t* my_slots = a.slots + (int64_t)blockIdx.x * kBU * CAP;
    const int kth = a.K - 1;

    // the two (row, quad) pairs this thread moves per chunk: idx = t and t + 256
    const int lrow0 = t >> 2, lq = t & 3;
    const int lrow1 = lrow0 + 64;

    for (int ut = blockIdx.x; ut < n_utiles; ut += gridDim.x) {
        const int u0 = ut * kBU;
        __syncthreads();  // previous tile fully retired
        if (t < kBU) {
            const int r = u0 + t;
            row_b[t] = (r < n_rows) ? (a.row_map ? a.row_map[r] : r) : -1;
            rows[t].cnt = 0;
            rows[t].thr_score = -INFINITY;
            rows[t].thr_key = 0ull;
        }
        __syncthreads();
        const float* arow[2];
        {
            const int b0 = row_b[lrow0], b1 = row_b[lrow1];
            arow[0] = b0 < 0 ? nullptr : a.Eu + (a.users ? a.users[b0] : (int64_t)b0) * a.lde_u;
            arow[1] = b1 < 0 ? nullptr : a.Eu + (a.users ? a.users[b1] : (int64_t)b1) * a.lde_u;
        }

        for (int it = 0; it < n_itiles; ++it) {
            const int i0 = it * kBI;
            const float* brow[2];
            brow[0] = (i0 + lrow0 < a.I) ? a.Ei + (int64_t)(i0 + lrow0) * a.lde_i : nullptr;
            brow[1] = (i0 + lrow1 < a.I) ? a.Ei + (int64_t)(i0 + lrow1) * a.lde_i : nullptr;

            float acc[8][8];
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
                const int item = i0 + (ii >> 2) * 64 + tx * 4 + (ii & 3);
                const float bv = (a.bias != nullptr && item < a.I) ? a.bias[item] : 0.f;
#pragma unroll
                for (int ui = 0; ui < 8; ++ui) acc[ui][ii] = bv;
            }

            float4 ra[2], rb[2];
            auto fetch = [&](int c) {
                const int d = c * kDK + lq * 4;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (VEC4) {
                        ra[h] = (arow[h] && d < a.D) ? *reinterpret_cast<const float4*>(arow[h] + d)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                        rb[h] = (brow[h] && d < a.D) ? *reinterpret_cast<const float4*>(brow[h] + d)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                    } else {
                        float va[4], vb[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            va[k] = (arow[h] && d + k < a.D) ? arow[h][d + k] : 0.f;
                            vb[k] = (brow[h] && d + k < a.D) ? brow[h][d + k] : 0.f;
                        }
                        ra[h] = make_float4(va[0], va[1], va[2], va[3]);
                        rb[h] = make_float4(vb[0], vb[1], vb[2], vb[3]);
                    }
                }
            };
            auto stash = [&](int buf) {
                const int r0 = lrow0, r1 = lrow1, d = lq * 4;
                As[buf][d + 0][r0] = ra[0].x; As[buf][d + 1][r0] = ra[0].y; As[buf][d + 2][r0] = ra[0].z; As[buf][d + 3][r0] = ra[0].w;
                As[buf][d + 0][r1] = ra[1].x; As[buf][d + 1][r1] = ra[1].y; As[buf][d + 2][r1] = ra[1].z; As[buf][d + 3][r1] = ra[1].w;
                Bs[buf][d + 0][r0] = rb[0].x; Bs[buf][d + 1][r0] = rb[0].y; Bs[buf][d + 2][r0] = rb[0].z; Bs[buf][d + 3][r0] = rb[0].w;
                Bs[buf][d + 0][r1] = rb[1].x; Bs[buf][d + 1][r1] = rb[1].y; Bs[buf][d + 2][r1] = rb[1].z; Bs[buf][d + 3][r1] = rb[1].w;
            };

            fetch(0);
            stash(0);
            __syncthreads();
            for (int c = 0; c < n_chunks; ++c) {
                const int buf = c & 1;
                if (c + 1 < n_chunks) fetch(c + 1);
#pragma unroll
                for (int d = 0; d < kDK; ++d) {
                    const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][d][ty * 4]);
                    const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][d][64 + ty * 4]);
                    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][d][tx * 4]);
                    const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][d][64 + tx * 4]);
                    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int ui = 0; ui < 8; ++ui)
#pragma unroll
                        for (int ii = 0; ii < 8; ++ii) acc[ui][ii] = fmaf(av[ui], bv[ii], acc[ui][ii]);
                }
                if (c + 1 < n_chunks) stash(buf ^ 1);
                __syncthreads();
            }

            // ---- selection: append what beats the row threshold --------------------------------
            uint64_t todo = 0ull;  // bit (ui * 8 + ii): candidate still to be appended
#pragma unroll
            for (int ui = 0; ui < 8; ++ui) {
                const int lr = (ui >> 2) * 64 + ty * 4 + (ui & 3);
                const float thr = rows[lr].thr_score;
#pragma unroll
                for (int ii = 0; ii < 8; ++ii)
                    if (acc[ui][ii] >= thr) todo |= 1ull << (ui * 8 + ii);
            }
            // drop padding rows / items
#pragma unroll
            for (int ui = 0; ui < 8; ++ui) {
                const int lr = (ui >> 2) * 64 + ty * 4 + (ui & 3);
                if (row_b[lr] < 0) todo &= ~(0xFFull << (ui * 8));
            }
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
                const int item = i0 + (ii >> 2) * 64 + tx * 4 + (ii & 3);
                if (item >= a.I) todo &= ~(0x0101010101010101ull << ii);
            }
            // the train-history mask is applied when a row is compacted (topk_select.cuh)
            while (true) {
                if (todo != 0ull) {
#pragma unroll
                    for (int ui = 0; ui < 8; ++ui) {
                        if (((todo >> (ui * 8)) & 0xFFull) == 0ull) continue;
                        const int lr = (ui >> 2) * 64 + ty * 4 + (ui & 3);
                        const uint64_t thr_key = rows[lr].thr_key;
#pragma unroll
                        for (int ii = 0; ii < 8; ++ii) {
                            const uint64_t bit = 1ull << (ui * 8 + ii);
                            if (!(todo & bit)) continue;
                            const int item = i0 + (ii >> 2) * 64 + tx * 4 + (ii & 3);
                            const uint64_t key = make_key(acc[ui][ii], item);
                            if (key <= thr_key) {
                                todo &= ~bit;
                                continue;
                            }
                            const int slot = atomicAdd(&rows[lr].cnt, 1);
                            if (slot < CAP) {
                                my_slots[(int64_t)lr * CAP + slot] = key;
                                todo &= ~bit;
                            }
                        }
                    }
                }
                const int left = __syncthreads_or(todo != 0ull);
                // compaction: each warp looks after 16 rows
                {
                    const int lr = warp * 16 + (lane & 15);
                    const bool need = (lane < 16) && (rows[lr].cnt >= TRIG);
                    unsigned m = __ballot_sync(0xffffffffu, need);
                    while (m) {
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        const int r = warp * 16 + l;
                        int64_t mlo = 0, mhi = 0;
                        if (a.mask_rowptr != nullptr) {
                            mlo = a.mask_rowptr[row_b[r]];
                            mhi = a.mask_rowptr[row_b[r] + 1];
                        }
                        compact_row<NPL>(my_slots + (int64_t)r * CAP, &rows[r], kth, lane, a.mask_items, mlo, mhi);
                    }
                }
                __syncthreads();
                if (!left) break;
            }
        }

        // ---- final compaction + output ----------------------------------------------------------
        for (int r = warp; r < kBU; r += kThreads / 32) {
            const int b = row_b[r];
            if (b < 0) continue;
            uint64_t* s = my_slots + (int64_t)r * CAP;
            int64_t mlo = 0, mhi = 0;
            if (a.mask_rowptr != nullptr) {
                mlo = a.mask_rowptr[b];
                mhi = a.mask_rowptr[b + 1];
            }
            compact_row<NPL>(s, &rows[r], kth, lane, a.mask_items, mlo, mhi);
            const int cnt = rows[r].cnt;
            for (int j = lane; j < a.K; j += 32) {
                const bool ok = j < cnt;
                const uint64_t key = ok ? s[j] : 0ull;
                a.out_ids[(int64_t)b * a.K + j] = ok ? key_id(key) : -1;
                if (a.out_scores) a.out_scores[(int64_t)b * a.K + j] = ok ? key_score(key) : -INFINITY;
            }
        }
    }
}

// ---- dense scores (API parity with full_sort_predict; not on the fused evaluation path) -------
// out[b, i] = bias[i] + sum_d Eu[users[b], d] * Ei[i, d], same fmaf chain as above.
__global__ void __launch_bounds__(256)
    scores_dense_kernel(const float* __restrict__ Eu, int64_t lde_u, const int64_t* __restrict__ users, int32_t B,
                        const float* __restrict__ Ei, int64_t lde_i, const float* __restrict__ bias, int32_t I,
                        int32_t D, float* __restrict__ out, int64_t ldo)
{
    constexpr int T = 64, DKc = 16;
    __shared__ float As[DKc][T + 1];
    __shared__ float Bs[DKc][T + 1];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int u0 = blockIdx.y * T, i0 = blockIdx.x * T;
    float acc[4][4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        const int item = i0 + tx + 16 * ii;
        const float bv = (bias != nullptr && item < I) ? bias[item] : 0.f;
#pragma unroll
        for (int ui = 0; ui < 4; ++ui) acc[ui][ii] = bv;
    }
    for (int d0 = 0; d0 < D; d0 += DKc) {
        __syncthreads();
        for (int idx = t; idx < T * DKc; idx += 256) {
            const int r = idx / DKc, d = idx % DKc;
            const int b = u0 + r, item = i0 + r;
            float av = 0.f, bv = 0.f;
            if (d0 + d < D) {
                if (b < B) av = Eu[(users ? users[b] : (int64_t)b) * lde_u + d0 + d];
                if (item < I) bv = Ei[(int64_t)item * lde_i + d0 + d];
            }
            As[d][r] = av;
            Bs[d][r] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < DKc; ++d) {
            float a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = As[d][ty + 16 * q];
                b[q] = Bs[d][tx + 16 * q];
            }
#pragma unroll
            for (int ui = 0; ui < 4; ++ui)
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) acc[ui][ii] = fmaf(a[ui], b[ii], acc[ui][ii]);
        }
    }
#pragma unroll
    for (int ui = 0; ui < 4; ++ui) {
        const int b = u0 + ty + 16 * ui;
        if (b >= B) continue;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int item = i0 + tx + 16 * ii;
            if (item < I) out[(int64_t)b * ldo + item] = acc[ui][ii];
        }
    }
}

int scores_dense_launch(const float* Eu, int64_t lde_u, const int64_t* users, int32_t B, const float* Ei,
                        int64_t lde_i, const float* bias, int32_t I, int32_t D, float* out, int64_t ldo,
                        cudaStream_t st)
{
    if (B == 0) return GMR_OK;
    dim3 grid((I + 63) / 64, (B + 63) / 64);
    scores_dense_kernel<<<grid, 256, 0, st>>>(Eu, lde_u, users, B, Ei, lde_i, bias, I, D, out, ldo);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

static int simt_grid(int32_t B)
{
    const int tiles = (B + kBU - 1) / kBU;
    return tiles < 2 * sm_count() ? tiles : 2 * sm_count();
}

static thread_local const int32_t* g_dynamic_rows = nullptr;
void score_simt_set_dynamic_rows(const int32_t* n_rows_dev) { g_dynamic_rows = n_rows_dev; }

static int npl_for(int32_t K) { return K <= 64 ? 8 : (K <= 128 ? 16 : 32); }

int64_t score_simt_workspace_bytes(int32_t B, int32_t K)
{
    return (int64_t)simt_grid(B) * kBU * (32 * npl_for(K)) * (int64_t)sizeof(uint64_t);
}

// Shared with the tensor-core path: run the exact fp32 kernel on `n_rows` batch rows (all rows when
// row_map is null).
int score_topk_simt_launch(const float* Eu, int64_t lde_u, const int64_t* users, const int32_t* row_map,
                           int32_t n_rows, const float* Ei, int64_t lde_i, const float* bias, int32_t I, int32_t D,
                           const int64_t* mask_rowptr, const int32_t* mask_items, int32_t K, int32_t* out_ids,
                           float* out_scores, void* workspace, cudaStream_t st, int grid_override)
{
    if (n_rows == 0) return GMR_OK;
    ScoreArgs a;
    a.Eu = Eu; a.lde_u = lde_u; a.users = users; a.row_map = row_map; a.B = n_rows; a.B_dev = g_dynamic_rows;
    a.Ei = Ei; a.lde_i = lde_i; a.bias = bias; a.I = I; a.D = D;
    a.mask_rowptr = mask_rowptr; a.mask_items = mask_items; a.K = K;
    a.out_ids = out_ids; a.out_scores = out_scores; a.slots = (uint64_t*)workspace;
    const int grid = grid_override > 0 ? grid_override : simt_grid(n_rows);
    const bool vec4 = (D % 4 == 0) && (lde_u % 4 == 0) && (lde_i % 4 == 0) && ((uintptr_t)Eu % 16 == 0) &&
                      ((uintptr_t)Ei % 16 == 0);
    const int npl = npl_for(K);
#define GMR_SIMT_LAUNCH(NPL)                                                         \
    do {                                                                             \
        if (vec4)                                                                    \
            score_topk_simt_kernel<NPL, true><<<grid, kThreads, 0, st>>>(a);         \
        else                                                                         \
            score_topk_simt_kernel<NPL, false><<<grid, kThreads, 0, st>>>(a);        \
    } while (0)
    if (npl == 8)
        GMR_SIMT_LAUNCH(8);
    else if (npl == 16)
        GMR_SIMT_LAUNCH(16);
    else
        GMR_SIMT_LAUNCH(32);
#undef GMR_SIMT_LAUNCH
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

}  // namespace gmr
