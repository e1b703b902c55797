// metrics.cu -- K4: hit matrix + Recall / NDCG / Precision / MAP prefix sums on the device.
//
// Replaces the pure-Python membership loop of GenMMRec/src/utils/topk_evaluator.py:107-112 and
// the numpy kernels of GenMMRec/src/utils/metrics.py:12-105 (mean over users taken by the caller,
// topk_evaluator.py:299-313).  Integer work (hits, cumulative hit counts) is bit-exact; the fp64
// sums are accumulated in a FIXED order (warp tree -> warps in order -> blocks strided/tree), so the
// result is run-to-run deterministic and within a few ulp of numpy's pairwise sum.  The warp tree is evaluated
// column-wise through a shared-memory tile (same additions in the same order as the shuffle butterfly it replaced).
#include <cmath>

#include <mutex>

#include "common.cuh"
#include "topk_select.cuh"

namespace gmr {

constexpr int kMU = 128;  // users per CTA (one thread each)
constexpr int kKC = 16;   // top-K positions staged per pass
constexpr int kKS = 4;    // positions whose per-user values are reduced together (16 columns = 4 metrics x 4 positions)
constexpr int kVP = 4 * kKS + 1;   // row pitch of the value tile in doubles (odd: conflict-free 64-bit stores)

__constant__ double c_discount[GMR_MAX_TOPK];  // 1 / log2(k + 2)

// block_sums layout: [4 * K][n_blocks]  (metric-major, then position, then block)
__global__ void __launch_bounds__(kMU)
    hits_metrics_kernel(const int32_t* __restrict__ topk, const int64_t* __restrict__ gt_rowptr,
                        const int32_t* __restrict__ gt_items, int32_t U, int32_t K, uint8_t* __restrict__ hit,
                        double* __restrict__ block_sums)
{
    __shared__ int32_t tile[kMU][kKC + 1];
    __shared__ double wsum[kMU / 32][4][kKC];
    __shared__ double vals[kMU][kVP];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t u0 = (int64_t)blockIdx.x * kMU;
    const int64_t u = u0 + t;
    const bool valid = u < U;
    int64_t lo = 0, hi = 0;
    if (valid) {
        lo = gt_rowptr[u];
        hi = gt_rowptr[u + 1];
    }
    const double n = (double)(hi - lo);
    const int lim = (int)min((int64_t)K, hi - lo);
    double c = 0.0, dcg = 0.0, idcg = 0.0, sp = 0.0;
    // c / n, dcg / idcg, sp / min(pos + 1, lim) as of the last change.  A user WITHOUT ground truth (the reference's
    // loaders never produce one) gets what GenMMRec/src/utils/metrics.py computes for pos_len == 0: recall = 0 / 0 = NaN,
    // but NDCG and MAP = 0, because `idcg[row, 0:] = idcg[row, -1]` (:54-55) and `ranges[0:] = ranges[-1]` (:86-87) wrap
    // around to the full-length normalisers.
    double q_recall = (hi > lo) ? 0.0 : nan(""), q_ndcg = 0.0, q_map = 0.0;

    for (int k0 = 0; k0 < K; k0 += kKC) {
        const int kc = min(kKC, K - k0);
        __syncthreads();
        // coalesced stage-in of topk[u0 : u0+128, k0 : k0+kc]
        for (int idx = t; idx < kMU * kc; idx += kMU) {
            const int r = idx / kc, k = idx - r * kc;
            tile[r][k] = (u0 + r < U) ? topk[(u0 + r) * K + k0 + k] : -1;
        }
        __syncthreads();
        for (int ks = 0; ks < kc; ks += kKS) {
#pragma unroll
            for (int kk = 0; kk < kKS; ++kk) {
                const int k = ks + kk;
                double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
                if (k < kc) {
                    const int pos = k0 + k;
                    int h = 0;
                    if (valid) h = sorted_contains(gt_items, lo, hi, tile[t][k]) ? 1 : 0;
                    tile[t][k] = h;
                    if (valid) {
                        // fp64 divides are the cost of the arithmetic (5 per user and position when written literally):
                        // the hit count c changes on a hit only, the ideal DCG only while pos < lim, so quotients are
                        // refreshed when their operands change and reused otherwise -- same expressions, same results
                        const double disc = c_discount[pos];
                        if (h) {
                            c += 1.0;
                            dcg += disc;
                            sp += c / (double)(pos + 1);
                            q_recall = c / n;
                        }
                        if (pos < lim) {
                            idcg += disc;
                            q_ndcg = dcg / idcg;
                            q_map = sp / (double)(pos + 1);
                        } else if (h) {
                            q_ndcg = dcg / idcg;
                            q_map = sp / (double)lim;
                        }
                        v0 = q_recall;
                        v1 = q_ndcg;
                        v2 = (c != 0.0) ? c / (double)(pos + 1) : 0.0;
                        v3 = q_map;
                    }
                }
                vals[t][0 * kKS + kk] = v0;
                vals[t][1 * kKS + kk] = v1;
                vals[t][2 * kKS + kk] = v2;
                vals[t][3 * kKS + kk] = v3;
            }
            __syncwarp();
            {
                // Column sums over this warp's 32 users in the order of the xor-butterfly (16, 8, 4, 2, 1) the kernel
                // used to run with shuffles -- 40 SHFL + 20 DADD per warp and position -- so the sums keep their bits:
                // lane = (h, column) adds the users of parity h as that tree does, the two parities meet in one shuffle.
                const int col = lane & 15, h = lane >> 4;
                const double* vw = &vals[warp * 32][col];
                double a8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a8[i] = vw[(h + 2 * i) * kVP] + vw[(h + 2 * i + 16) * kVP];   // a_j, j = h + 2 i
                const double b0 = a8[0] + a8[4], b1 = a8[1] + a8[5], b2 = a8[2] + a8[6], b3 = a8[3] + a8[7];   // b_j = a_j + a_{j+8}
                const double c0 = b0 + b2, c1 = b1 + b3;                                                   // c_j = b_j + b_{j+4}
                double dsum = c0 + c1;                                                                     // d_h = c_h + c_{h+2}
                dsum += __shfl_xor_sync(0xffffffffu, dsum, 16);                                            // d_0 + d_1
                const int m = col / kKS, kk = col - m * kKS;
                if (h == 0 && ks + kk < kc) wsum[warp][m][ks + kk] = dsum;
            }
            __syncwarp();
        }
        __syncthreads();
        if (hit != nullptr) {
            for (int idx = t; idx < kMU * kc; idx += kMU) {
                const int r = idx / kc, k = idx - r * kc;
                if (u0 + r < U) hit[(u0 + r) * K + k0 + k] = (uint8_t)tile[r][k];
            }
        }
        for (int idx = t; idx < 4 * kc; idx += kMU) {
            const int m = idx / kc, k = idx - m * kc;
            double s = wsum[0][m][k];
#pragma unroll
            for (int w = 1; w < kMU / 32; ++w) s += wsum[w][m][k];
            block_sums[((int64_t)m * K + k0 + k) * gridDim.x + blockIdx.x] = s;
        }
    }
}

// one CTA per (metric, position): strided fixed-order partial sums, then a shared-memory tree
__global__ void __launch_bounds__(256) metrics_final_kernel(const double* __restrict__ block_sums, int32_t n_blocks,
                                                            double* __restrict__ sums)
{
    __shared__ double red[256];
    const double* src = block_sums + (int64_t)blockIdx.x * n_blocks;
    double s = 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += 256) s += src[b];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = red[0];
}

// one flag per device: the __constant__ table exists once per device (a second GPU driven from the same process
// would otherwise run with an all-zero table); the mutex makes the lazy upload thread-safe
static bool g_discount_ready[64] = {false};
static std::mutex g_discount_mu;

}  // namespace gmr

extern "C" int64_t gmr_hits_metrics_workspace_bytes(int32_t U, int32_t K)
{
    if (U <= 0 || K <= 0) return 0;
    const int64_t nb = ((int64_t)U + gmr::kMU - 1) / gmr::kMU;
    return gmr::align_up(nb * 4 * K * (int64_t)sizeof(double), 256);
}

extern "C" int gmr_hits_metrics(const int32_t* topk, const int64_t* gt_rowptr, const int32_t* gt_items, int32_t U,
                                int32_t K, uint8_t* hit, double* sums, void* workspace, int64_t workspace_bytes,
                                void* stream)
{
    GMR_REQUIRE(K >= 1 && K <= GMR_MAX_TOPK, "gmr_hits_metrics: K=%d outside [1, %d]", K, GMR_MAX_TOPK);
    GMR_REQUIRE(U >= 0, "gmr_hits_metrics: negative user count");
    GMR_REQUIRE(sums != nullptr, "gmr_hits_metrics: null sums");
    cudaStream_t st = (cudaStream_t)stream;
    if (U == 0) {
        GMR_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 4 * K, st));
        return GMR_OK;
    }
    GMR_REQUIRE(topk && gt_rowptr && gt_items, "gmr_hits_metrics: null operand");
    const int64_t need = gmr_hits_metrics_workspace_bytes(U, K);
    if (workspace == nullptr || workspace_bytes < need) {
        gmr::set_error("gmr_hits_metrics: workspace of %lld bytes required, %lld given", (long long)need,
                       (long long)workspace_bytes);
        return GMR_ERR_WORKSPACE;
    }
    {
        int dev = 0;
        GMR_CHECK_CUDA(cudaGetDevice(&dev));
        GMR_REQUIRE(dev >= 0 && dev < 64, "gmr_hits_metrics: device ordinal %d outside [0, 64)", dev);
        std::lock_guard<std::mutex> lock(gmr::g_discount_mu);
        if (!gmr::g_discount_ready[dev]) {
            // same expression as GenMMRec/src/utils/metrics.py:53,59: 1.0 / np.log2(rank + 1), rank = k + 1.
            // Synchronous copy: the first call on a device must not happen inside a CUDA-graph capture
            // (Trainer.graphed runs the step eagerly first).
            double h[GMR_MAX_TOPK];
            for (int k = 0; k < GMR_MAX_TOPK; ++k) h[k] = 1.0 / std::log2((double)k + 2.0);
            GMR_CHECK_CUDA(cudaMemcpyToSymbol(gmr::c_discount, h, sizeof(h)));
            gmr::g_discount_ready[dev] = true;
        }
    }
    const int nb = (int)(((int64_t)U + gmr::kMU - 1) / gmr::kMU);
    gmr::hits_metrics_kernel<<<nb, gmr::kMU, 0, st>>>(topk, gt_rowptr, gt_items, U, K, hit, (double*)workspace);
    GMR_LAUNCH_CHECK();
    gmr::metrics_final_kernel<<<4 * K, 256, 0, st>>>((const double*)workspace, nb, sums);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}
