// score_topk_screen.cu -- K2 + K3, second generation: fused score + mask + top-K where the dense
// contraction is ONE fp16 tcgen05 pass used as a certified SCREEN over a NORM-ORDERED catalogue, followed by
// an exact fp32 re-score of the few candidates that can still reach the top K.  sm_100a only.
//
// Replaces the same reference call chain as score_topk_simt.cu / score_topk_tc.cu
//   torch.matmul(u_e[user], i_e.T) -> scores[mask] = -1e10 -> torch.topk   (models/diffmm.py:276-278,
//   common/trainer.py:381-386) and returns THE SAME ids/scores as the fp32 path.
//
//   1. operands: user rows are scaled by a per-row power of two, item rows by one global power of two
//      (exact, order-preserving per row) so that every element fits fp16 with 11 significant bits, then
//      rounded to fp16.  The approximate score s~ (fp32 accumulation in TMEM) obeys
//          |s~(u, i) - s(u, i)| <= eps(u, i) = kScreenErr * |u|_2 * |e_i|_2 + small absolute term   (scaled units)
//      (2 * 2^-11 operand rounding + accumulation slack; subnormal flushes are covered by the absolute term).
//      The bound is PER ITEM: propagated item embeddings have a heavy-tailed norm distribution (max/median
//      ~ 85 at the 1M-user shape), a max-norm bound would keep hundreds of useless candidates per row.
//   2. screen: every item has an interval [lb, ub] = s~ -+ eps around its exact score.  If L is the K-th
//      largest lb among ANY set of unmasked items of the row, the exact K-th best score is >= L, so an item
//      with ub < L cannot be among the exact top K.  The sweep keeps, per row, every item with ub >= the
//      running L; nothing else is ever stored.  Whole 32-column chunks are rejected with one compare.
//   3. norm order + early stop (exact maximum-inner-product pruning by Cauchy-Schwarz): items are swept in
//      DESCENDING NORM order (cub radix sort per call), so every item at or after sweep position p
//      scores at most |u| * nb[p].  Once |u| * nb[p] < L for all 256 rows of a CTA, no remaining item
//      can enter any of their top-K lists and the CTA stops sweeping.  With popularity-skewed
//      embeddings this ends after a few tiles; with flat norms it degenerates to the full sweep.  The
//      stop tile is agreed by all warps at doubling checkpoints (1, 2, 4, ... tiles), so a full sweep pays
//      ~12 pipeline drains.
//   4. one persistent CTA per SM owns 2 x 128 users: each TMA-loaded item tile (128 items) is
//      multiplied against BOTH resident user tiles (the item operand is read from L2 once per 256
//      users), accumulators double-buffered in all 512 TMEM columns:
//        warp 0      TMA producer
//        warp 1      MMA issuer (one elected thread)
//        warps 2-9   epilogue: thread = user row.  Fast path per 32 columns: tcgen05.ld, a 3-input max
//                    tree, one compare against the row threshold.  Rows that see a survivor dump the 32
//                    scores to a swizzled shared-memory strip and the WARP appends them cooperatively
//                    (lane = column), so a hit costs the warp ~12 instructions instead of a 32-way
//                    divergent scan.  The train-history mask is applied when a row's buffer is pruned
//                    (256 binary searches in flight), never per score.
//   5. end of a user tile: prune, exact fp32 fmaf re-score of the survivors, sort by (score desc, id asc),
//      write the top K.  Rows whose buffer overflowed (massive near-ties), rows with fewer than K
//      unmasked items and rows with non-finite scores are queued for the fp32 kernel.
//   0. exact HEAD (before any of the above, no bias, K <= 128): the n_hot = 128 / 256 highest-norm items are scored
//      EXACTLY for every row by a register-tiled CUDA-core product whose per-score operation order is the oracle's
//      sequential fmaf chain; the row's exact K-th best head score s_K is compared with the Cauchy-Schwarz bound of
//      everything outside the head, |u| * nb[n_hot]: if the bound is smaller the head's top K IS the answer and the
//      row is written and flagged done -- no screen, no intervals, no re-score.  With popularity-skewed embeddings
//      that settles ~98 % of the rows; the screen pipeline then only touches 256-user groups with a live row.  With
//      flat norms (nb[n_hot] >= nb[K-1]: no row can finish) the head switches itself off on the device.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "topk_select.cuh"

namespace gmr {

// fp32 kernel entry used for the queued rows (score_topk_simt.cu)
int score_topk_simt_launch(const float* Eu, int64_t lde_u, const int64_t* users, const int32_t* row_map,
                           int32_t n_rows, const float* Ei, int64_t lde_i, const float* bias, int32_t I, int32_t D,
                           const int64_t* mask_rowptr, const int32_t* mask_items, int32_t K, int32_t* out_ids,
                           float* out_scores, void* workspace, cudaStream_t st, int grid_override);
int64_t score_simt_workspace_bytes(int32_t B, int32_t K);
void score_simt_set_dynamic_rows(const int32_t* n_rows_dev);

constexpr int kSM = 128;                        // users per UMMA tile (M)
constexpr int kSN = 128;                        // items per tile (N)
constexpr int kUT = 2;                          // user tiles per CTA
constexpr int kRowsPerCta = kUT * kSM;          // 256
constexpr int kEpiWarps = kUT * 4;              // 8
constexpr int kScrThreads = 64 + kEpiWarps * 32;  // 320
constexpr int kScrMaxStages = 10;
constexpr uint32_t kAtomBytes = 128 * 128;      // one 128-row x 64-half K-atom
constexpr float kScreenErr = 1.05e-3f;          // 2^-10 (1 + 2^-11) operand rounding + accumulation slack
constexpr int kStatSlots = 8;

// ---- 1. operand preparation ---------------------------------------------------------------------
__device__ __forceinline__ float pow2_scale_for(float absmax)
{
    // power of two s with absmax * s in [2^13, 2^14); 1 for zero / non-finite input
    if (!(absmax > 0.f) || !(absmax < INFINITY)) return 1.f;
    const int e = (int)((__float_as_uint(absmax) >> 23) & 0xffu) - 127;
    int sh = 13 - e;
    sh = sh < -60 ? -60 : (sh > 60 ? 60 : sh);
    return __uint_as_float((uint32_t)(sh + 127) << 23);
}

__global__ void __launch_bounds__(256)
    absmax_kernel(const float* __restrict__ E, int64_t lde, int64_t n_rows, int32_t D, uint32_t* __restrict__ out_bits)
{
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    float m = 0.f;
    for (int64_t r = w0; r < n_rows; r += nw)
        for (int d = lane; d < D; d += 32) m = fmaxf(m, fabsf(E[r * lde + d]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));
}

// item norms (UNSCALED, rounded up: they feed BOUNDS) as sortable bits + the identity permutation, and the global
// absmax that fixes the power-of-two item scale -- one pass over the item table.  The sort order does not depend on the
// scale; prep_items_kernel multiplies the sorted norms by it (exact: a power of two).
__global__ void __launch_bounds__(256)
    item_norms_kernel(const float* __restrict__ E, int64_t lde, int32_t n_rows, int32_t D, uint32_t* __restrict__ gmax_bits,
                      uint32_t* __restrict__ norm_bits, int32_t* __restrict__ ident)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    float ss = 0.f, m = 0.f;
    if ((lde & 1) == 0 && ((uintptr_t)E & 7) == 0) {   // D % 64 == 0: 8-byte loads
        const float2* e2 = reinterpret_cast<const float2*>(E + r * lde);
        for (int d = lane; d < D / 2; d += 32) {
            const float2 x = e2[d];
            ss = fmaf(x.x, x.x, fmaf(x.y, x.y, ss));
            m = fmaxf(m, fmaxf(fabsf(x.x), fabsf(x.y)));
        }
    } else {
        for (int d = lane; d < D; d += 32) {
            const float x = E[r * lde + d];
            ss = fmaf(x, x, ss);
            m = fmaxf(m, fabsf(x));
        }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, k);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
    }
    if (lane == 0) {
        float nrm = sqrtf(ss) * 1.000001f;
        if (!(nrm < INFINITY)) nrm = 3.0e38f;  // NaN / inf rows sort first and never allow an early stop
        norm_bits[r] = __float_as_uint(nrm);
        ident[r] = (int32_t)r;
        // one address for every row: test first (a stale read only costs a redundant atomic), so that after the first
        // few rows almost no atomic is issued
        if (m > 0.f && m < INFINITY && __float_as_uint(m) > *reinterpret_cast<volatile uint32_t*>(gmax_bits))
            atomicMax(gmax_bits, __float_as_uint(m));
    }
}

// items in sweep order: out[p, :] = fp16(E[perm[p], :] * s_i); positions >= n_rows are zero rows with
// norm 0 and id -1
__global__ void __launch_bounds__(256)
    prep_items_kernel(const float* __restrict__ E, int64_t lde, int32_t n_rows, int32_t n_rows_pad, int32_t D,
                      const uint32_t* __restrict__ gmax_bits, int32_t* __restrict__ perm, float* __restrict__ nb,
                      __half* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n_rows_pad) return;
    const float s = pow2_scale_for(__uint_as_float(*gmax_bits));
    __half* o = out + p * (int64_t)D;
    const float* src = (p < n_rows) ? E + (int64_t)perm[p] * lde : nullptr;
    if (src != nullptr && (lde & 1) == 0 && ((uintptr_t)E & 7) == 0) {
        const float2* s2 = reinterpret_cast<const float2*>(src);
        __half2* o2 = reinterpret_cast<__half2*>(o);
        for (int d = lane; d < D / 2; d += 32) {
            const float2 x = s2[d];
            o2[d] = __floats2half2_rn(x.x * s, x.y * s);
        }
    } else {
        for (int d = lane; d < D; d += 32) o[d] = __float2half_rn(src ? src[d] * s : 0.f);
    }
    if (lane == 0) {
        if (p >= n_rows) {
            perm[p] = -1;
            nb[p] = 0.f;
        } else {
            nb[p] *= s;   // sorted unscaled norm -> scaled units
        }
    }
}

// D = 64, 16-byte aligned rows: a half-warp per row (16 lanes x 128 bit), four rows in flight per half-warp -- the
// one-row-per-warp kernels above keep a single 8-byte load per lane in flight and reach a fifth of the copy bandwidth
// on this 128 MB pass, which every rank of a multi-GPU run repeats.
__global__ void __launch_bounds__(256)
    item_norms_d64_kernel(const float* __restrict__ E, int64_t lde, int32_t n_rows, uint32_t* __restrict__ gmax_bits,
                          uint32_t* __restrict__ norm_bits, int32_t* __restrict__ ident)
{
    const int lane = threadIdx.x & 31, sub = lane & 15, half = lane >> 4;
    const int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + half;
    float4 x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t r = r0 + 2 * q;
        x[q] = r < n_rows ? *reinterpret_cast<const float4*>(E + r * lde + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float mm = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t r = r0 + 2 * q;
        float ss = fmaf(x[q].x, x[q].x, fmaf(x[q].y, x[q].y, fmaf(x[q].z, x[q].z, x[q].w * x[q].w)));
        float m = fmaxf(fmaxf(fabsf(x[q].x), fabsf(x[q].y)), fmaxf(fabsf(x[q].z), fabsf(x[q].w)));
#pragma unroll
        for (int k = 8; k > 0; k >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, k);
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
        }
        if (sub == 0 && r < n_rows) {
            float nrm = sqrtf(ss) * 1.000001f;
            if (!(nrm < INFINITY)) nrm = 3.0e38f;
            norm_bits[r] = __float_as_uint(nrm);
            ident[r] = (int32_t)r;
            if (m > 0.f && m < INFINITY) mm = fmaxf(mm, m);
        }
    }
    mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, 16));
    if (lane == 0 && mm > 0.f && __float_as_uint(mm) > *reinterpret_cast<volatile uint32_t*>(gmax_bits))
        atomicMax(gmax_bits, __float_as_uint(mm));
}

__global__ void __launch_bounds__(256)
    prep_items_d64_kernel(const float* __restrict__ E, int64_t lde, int32_t n_rows, int32_t n_rows_pad,
                          const uint32_t* __restrict__ gmax_bits, int32_t* __restrict__ perm, float* __restrict__ nb,
                          __half* __restrict__ out)
{
    const int lane = threadIdx.x & 31, sub = lane & 15, half = lane >> 4;
    const int64_t p0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + half;
    const float s = pow2_scale_for(__uint_as_float(*gmax_bits));
    int src[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t p = p0 + 2 * q;
        src[q] = p < n_rows ? perm[p] : -1;
    }
    float4 x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        x[q] = src[q] >= 0 ? *reinterpret_cast<const float4*>(E + (int64_t)src[q] * lde + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t p = p0 + 2 * q;
        if (p >= n_rows_pad) continue;
        const __half2 lo = __floats2half2_rn(x[q].x * s, x[q].y * s), hi = __floats2half2_rn(x[q].z * s, x[q].w * s);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(out + p * 64 + 4 * sub) = o;
        if (sub == 0) {
            if (p >= n_rows) {
                perm[p] = -1;
                nb[p] = 0.f;
            } else {
                nb[p] *= s;
            }
        }
    }
}

// users: per-row power-of-two scale; writes the row's error coefficients (eps(u, i) = ce * |e_i| + ab, scaled
// units), its norm and the bias scale
struct RowConst {
    float ce, ab;   // eps(u, i) = ce * nb[i] + ab
    float na;       // scaled row norm (rounded up)
    float sc;       // s_u * s_i: scale applied to the bias
};

__global__ void __launch_bounds__(256)
    prep_users_kernel(const float* __restrict__ E, int64_t lde, const int64_t* __restrict__ rows, int32_t n_rows,
                      int32_t n_rows_pad, int32_t D, const uint32_t* __restrict__ gmax_bits, const float* __restrict__ nb,
                      const uint32_t* __restrict__ biasmax_bits, const uint8_t* __restrict__ row_done,
                      __half* __restrict__ out, RowConst* __restrict__ row_const)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows_pad) return;
    if (r < n_rows && row_done[r]) return;   // settled by the exact head: nothing reads its operand or constants again
    __half* o = out + r * (int64_t)D;
    const float* src = (r < n_rows) ? E + (rows ? rows[r] : r) * lde : nullptr;
    // the row stays in registers between the absmax pass and the conversion (D <= 256: 4 x float2 per lane)
    const bool vec = src != nullptr && (lde & 1) == 0 && ((uintptr_t)E & 7) == 0 && D <= 256;
    float2 xr[4];
    float m = 0.f;
    if (vec) {
        const float2* s2 = reinterpret_cast<const float2*>(src);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = lane + 32 * j;
            xr[j] = (d < D / 2) ? s2[d] : make_float2(0.f, 0.f);
            m = fmaxf(m, fmaxf(fabsf(xr[j].x), fabsf(xr[j].y)));
        }
    } else if (src) {
        for (int d = lane; d < D; d += 32) m = fmaxf(m, fabsf(src[d]));
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
    const float su = pow2_scale_for(m);
    float ss = 0.f;
    if (vec) {
        __half2* o2 = reinterpret_cast<__half2*>(o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = lane + 32 * j;
            if (d < D / 2) {
                const float x0 = xr[j].x * su, x1 = xr[j].y * su;
                o2[d] = __floats2half2_rn(x0, x1);
                ss = fmaf(x0, x0, fmaf(x1, x1, ss));
            }
        }
    } else {
        for (int d = lane; d < D; d += 32) {
            const float x = src ? src[d] * su : 0.f;
            o[d] = __float2half_rn(x);
            ss = fmaf(x, x, ss);
        }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, k);
    if (lane == 0) {
        float na = sqrtf(ss) * 1.000001f;
        if (!(na < INFINITY)) na = INFINITY;
        const float nbmax = nb[0];  // sweep order is norm-descending
        const float si = pow2_scale_for(__uint_as_float(*gmax_bits));
        RowConst rc;
        rc.sc = su * si;
        rc.ce = kScreenErr * na;
        rc.ab = 1e-6f * (na + nbmax);
        if (biasmax_bits != nullptr) {
            rc.ce += 2.4e-7f * na;
            rc.ab += 2.4e-7f * rc.sc * __uint_as_float(*biasmax_bits);
        }
        rc.na = na;
        row_const[r] = rc;
    }
}

// D = 64 form of prep_users_kernel (half-warp per row, four rows in flight per half-warp; see item_norms_d64_kernel)
__global__ void __launch_bounds__(256)
    prep_users_d64_kernel(const float* __restrict__ E, int64_t lde, const int64_t* __restrict__ rows, int32_t n_rows,
                          int32_t n_rows_pad, const uint32_t* __restrict__ gmax_bits, const float* __restrict__ nb,
                          const uint32_t* __restrict__ biasmax_bits, const uint8_t* __restrict__ row_done,
                          __half* __restrict__ out, RowConst* __restrict__ row_const)
{
    const int lane = threadIdx.x & 31, sub = lane & 15, half = lane >> 4;
    const int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + half;
    int64_t src[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t r = r0 + 2 * q;
        // rows settled by the exact head are not read: their fp16 row only feeds their own (ignored) accumulators
        src[q] = (r < n_rows && !row_done[r]) ? (rows ? rows[r] : r) : -1;
    }
    float4 x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        x[q] = src[q] >= 0 ? *reinterpret_cast<const float4*>(E + src[q] * lde + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float nbmax = nb[0];  // sweep order is norm-descending
    const float si = pow2_scale_for(__uint_as_float(*gmax_bits));
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t r = r0 + 2 * q;
        float m = fmaxf(fmaxf(fabsf(x[q].x), fabsf(x[q].y)), fmaxf(fabsf(x[q].z), fabsf(x[q].w)));
#pragma unroll
        for (int k = 8; k > 0; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
        const float su = pow2_scale_for(m);
        const float x0 = x[q].x * su, x1 = x[q].y * su, x2 = x[q].z * su, x3 = x[q].w * su;
        float ss = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, x3 * x3)));
#pragma unroll
        for (int k = 8; k > 0; k >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, k);
        if (r >= n_rows_pad) continue;
        // settled by the exact head: nothing reads the row's operand or constants again (its stale fp16 row only
        // feeds its own, ignored, accumulators)
        if (r < n_rows && row_done[r]) continue;
        const __half2 lo = __floats2half2_rn(x0, x1), hi = __floats2half2_rn(x2, x3);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(out + r * 64 + 4 * sub) = o;
        if (sub == 0) {
            float na = sqrtf(ss) * 1.000001f;
            if (!(na < INFINITY)) na = INFINITY;
            RowConst rc;
            rc.sc = su * si;
            rc.ce = kScreenErr * na;
            rc.ab = 1e-6f * (na + nbmax);
            if (biasmax_bits != nullptr) {
                rc.ce += 2.4e-7f * na;
                rc.ab += 2.4e-7f * rc.sc * __uint_as_float(*biasmax_bits);
            }
            rc.na = na;
            row_const[r] = rc;
        }
    }
}

// ---- 2. selection helpers (warp-cooperative, one row at a time) -----------------------------------
struct PruneResult {
    int cnt;      // keys kept in slots [0, cnt); -1: the row overflowed (near-ties defeat the screen)
    float lbK;    // new lower bound L of the row's exact K-th best score (approximate-score units)
};

// Keep only the keys that can still matter.  Keys at positions >= n_checked are first tested against the
// row's train-history list (ascending item ids, binary search) and dropped when masked.  Every remaining
// key (s~, sweep position p) carries the interval [lb, ub] = s~ -+ (ce * nb[p] + ab).  L = K-th
// largest lb of a POOL of keys (any subset of the row's unmasked keys gives a valid lower bound of the
// exact K-th score): the pool is each lane's M = NPL / 2 largest lb (32 * M >= 2 K keys; with slots
// filled in arrival order the pool's K-th is within a few ranks of the true K-th), sorted with a 32-bit
// bitonic network.  Keys with ub < L are dropped, the rest compacted in place (unsorted).  If that frees
// too little, L is recomputed from ALL keys (exact K-th lb).
//
// Key layout here: (ordered s~) << 32 | sweep position p  (NOT the item id: nb[] and perm[] are indexed by p).
template <int NPL, int NACT>
__device__ __noinline__ PruneResult screen_prune_impl(uint64_t* __restrict__ s_row, int cnt, int n_checked, float ce, float ab,
                                                 const float* __restrict__ nb, const int32_t* __restrict__ perm,
                                                 const int32_t* __restrict__ mask_items, int64_t mlo, int64_t mhi, int K,
                                                 int lane, unsigned long long* stats)
{
    // NACT <= NPL key registers per lane are live (cnt <= 32 * NACT): a half-full buffer -- the state at every
    // checkpoint after phase 0 -- costs half the loads, searches and compare-exchanges
    constexpr int CAP = 32 * NPL;
    constexpr int M = NPL / 2;
    static_assert(NACT == NPL || NACT == M, "NACT is NPL or NPL / 2");
    uint64_t k[NACT];
    uint32_t lb[NACT];  // order-preserving bits; 0 = empty
    float ub[NACT];
#pragma unroll
    for (int r = 0; r < NACT; ++r) {
        const int i = r * 32 + lane;
        k[r] = (i < cnt) ? s_row[i] : 0ull;
    }
    if (mlo < mhi) {
        // NPL binary searches in lock-step over the row's ascending train-history list: every step issues NPL
        // independent loads (branch-free lower bound), so the chain costs log2(len) latencies, not NPL * log2(len)
        int32_t id[NACT];
        int32_t base[NACT];   // offsets from mlo
#pragma unroll
        for (int r = 0; r < NACT; ++r) {
            const int i = r * 32 + lane;
            id[r] = (k[r] != 0ull && i >= n_checked) ? perm[(uint32_t)k[r]] : -1;  // -1: nothing to look up
            base[r] = 0;
        }
        const int32_t* __restrict__ ml = mask_items + mlo;
        int32_t len = (int32_t)(mhi - mlo);
        while (len > 1) {
            const int32_t half = len >> 1;
#pragma unroll
            for (int r = 0; r < NACT; ++r) base[r] += (ml[base[r] + half - 1] < id[r]) ? half : 0;
            len -= half;
        }
#pragma unroll
        for (int r = 0; r < NACT; ++r)
            if (id[r] >= 0 && ml[base[r]] == id[r]) k[r] = 0ull;
    }
#pragma unroll
    for (int r = 0; r < NACT; ++r) {
        lb[r] = 0u;
        ub[r] = -INFINITY;
        if (k[r] != 0ull) {
            const float e = fmaf(ce, nb[(uint32_t)k[r]], ab);
            const float sc = ordered_to_f32((uint32_t)(k[r] >> 32));
            lb[r] = f32_to_ordered(sc - e);
            ub[r] = sc + e;
        }
    }
    // pool: per-lane M largest lb (insertion into a tiny sorted list)
    uint32_t pool[M];
#pragma unroll
    for (int m = 0; m < M; ++m) pool[m] = 0u;
#pragma unroll
    for (int r = 0; r < NACT; ++r) {
        uint32_t x = lb[r];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint32_t hi = pool[m] > x ? pool[m] : x;
            x = pool[m] > x ? x : pool[m];
            pool[m] = hi;
        }
    }
    warp_bitonic_sort_desc<M, uint32_t>(pool, lane);
    const int kth = K - 1;  // K <= 32 * M
    uint32_t lk = 0u;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const uint32_t cand = __shfl_sync(0xffffffffu, pool[m], kth & 31);
        if (m == (kth >> 5)) lk = cand;
    }
    float L = (lk != 0u) ? ordered_to_f32(lk) : -INFINITY;
    int total = 0;
#pragma unroll
    for (int r = 0; r < NACT; ++r) total += __popc(__ballot_sync(0xffffffffu, lb[r] != 0u && ub[r] >= L));
    bool exact = false;
    if (total > CAP * 5 / 8) {
        // exact K-th largest lb over all keys
        exact = true;
        uint32_t all[NACT];
#pragma unroll
        for (int r = 0; r < NACT; ++r) all[r] = lb[r];
        warp_bitonic_sort_desc<NACT, uint32_t>(all, lane);
        lk = 0u;
#pragma unroll
        for (int r = 0; r < NACT; ++r) {
            const uint32_t cand = __shfl_sync(0xffffffffu, all[r], kth & 31);
            if (r == (kth >> 5)) lk = cand;
        }
        L = (lk != 0u) ? ordered_to_f32(lk) : -INFINITY;
        total = 0;
#pragma unroll
        for (int r = 0; r < NACT; ++r) total += __popc(__ballot_sync(0xffffffffu, lb[r] != 0u && ub[r] >= L));
    }
    // compact survivors in place, keeping their slot order (= sweep order for keys appended by the sweep): neighbouring
    // slots then hold neighbouring sweep positions, and the finalisation's 32 lanes read 32 rows of the staged item
    // block whose shared-memory bank groups (position mod 8) differ.  (A lane-major compaction put positions 32 apart
    // into neighbouring slots: 3x the shared-memory wavefronts, the limiter of the finalize kernel.)
    const uint32_t lt = (1u << lane) - 1u;
    int base = 0;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NACT; ++r) {
        const bool keep = lb[r] != 0u && ub[r] >= L;
        const unsigned kb = __ballot_sync(0xffffffffu, keep);
        if (keep) s_row[base + __popc(kb & lt)] = k[r];
        base += __popc(kb);
    }
    __syncwarp();
    if (stats && lane == 0) atomicAdd(&stats[exact ? 3 : 2], 1ull);
    PruneResult res;
    res.cnt = (total > CAP - 64) ? -1 : total;
    res.lbK = (total > CAP - 64) ? INFINITY : L;
    return res;
}

template <int NPL>
__device__ __forceinline__ PruneResult screen_prune(uint64_t* __restrict__ s_row, int cnt, int n_checked, float ce, float ab,
                                                    const float* __restrict__ nb, const int32_t* __restrict__ perm,
                                                    const int32_t* __restrict__ mask_items, int64_t mlo, int64_t mhi, int K,
                                                    int lane, unsigned long long* stats)
{
    if (cnt <= 16 * NPL && K <= 16 * NPL)   // warp-uniform
        return screen_prune_impl<NPL, NPL / 2>(s_row, cnt, n_checked, ce, ab, nb, perm, mask_items, mlo, mhi, K, lane, stats);
    return screen_prune_impl<NPL, NPL>(s_row, cnt, n_checked, ce, ab, nb, perm, mask_items, mlo, mhi, K, lane, stats);
}

// ---- 3. the fused kernel ----------------------------------------------------------------------------
struct ScrArgs {
    const float* Eu;
    int64_t lde_u;
    const int64_t* users;
    int32_t B;
    const float* Ei;
    int64_t lde_i;
    const float* bias;
    int32_t I, D;
    const int64_t* mask_rowptr;
    const int32_t* mask_items;
    int32_t K;
    int32_t* out_ids;
    float* out_scores;
    uint64_t* slots;            // [B_pad][CAP] candidate keys of every row: (ordered s~) << 32 | sweep position
    const RowConst* row_const;  // [B_pad]
    const float* nb;            // [I_pad] scaled item norms in sweep order (descending, rounded up)
    const int32_t* perm;        // [I_pad] sweep position -> item id (-1: padding)
    const uint32_t* biasmax_bits;
    // per-row sweep state carried between the kernels of one call
    int32_t* row_cnt;           // [B_pad] keys in the row's slots; -1: overflowed (fp32 path)
    int32_t* row_chk;           // [B_pad] leading keys already tested against the train-history mask
    float* row_L;               // [B_pad] lower bound of the row's exact K-th score; +inf: nothing left to collect
    int32_t* group_need;        // [n_groups] item tiles the 256-user group still needs (max over its rows)
    int32_t* work_counter;      // [1] dynamic group scheduler of the continuation sweep
    int32_t* fallback_rows;     // [B]
    int32_t* fallback_count;    // [1]
    int32_t n_stages;
    int32_t vec4;               // Eu / Ei rows are 16-byte aligned
    int32_t first_check;        // tiles swept before the first stop check (phase 0)
    uint8_t* row_done;          // [B_pad] 1: the exact head already wrote the row's top K
    int32_t* live_rows;         // [B] rows the head left to the screen, in arrival order (valid when *n_live >= 0)
    int32_t* n_live;            // [1] -1: the head did not run, every row is live
    unsigned long long* stats;  // [kStatSlots] or null
    int32_t debug;              // GMR_TC_DEBUG timing experiments: 1 = drain TMEM only, 2 = filter without appends,
                                // 3 = no early stop (full sweep; results stay exact)
};

__device__ __forceinline__ bool row_force_exact(const ScrArgs& a, int64_t b, int64_t& mlo, int64_t& mhi)
{
    mlo = mhi = 0;
    if (a.mask_rowptr != nullptr) {
        mlo = a.mask_rowptr[b];
        mhi = a.mask_rowptr[b + 1];
    }
    return (int64_t)a.I - (mhi - mlo) < (int64_t)a.K;  // masked items would surface: the fp32 kernel owns that case
}

// smallest tile index j in [lo, hi] with  |u| * nb[j * 128] (+ bias bound) < L : tiles [lo, j) can still matter
__device__ __forceinline__ int tiles_needed(const ScrArgs& a, const RowConst& rc, float L, float bias_abs_max, int lo, int hi)
{
    if (L == INFINITY) return lo;   // finished / dead / padding row
    if (!(L > -INFINITY)) return hi;  // nothing known yet
    const float bt = rc.sc * bias_abs_max;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (fmaf(rc.na, a.nb[(int64_t)mid * kSN], bt) * 1.000001f < L)
            hi = mid;
        else
            lo = mid + 1;
    }
    return hi;
}

template <int NSORT>
__device__ __forceinline__ void final_sort_write(uint64_t (&ek)[NSORT], int lane, const ScrArgs& a, int64_t rb)
{
    warp_bitonic_sort_desc<NSORT>(ek, lane);
#pragma unroll
    for (int c = 0; c < NSORT; ++c) {
        const int j = c * 32 + lane;
        if (j < a.K) {
            const bool ok = ek[c] != 0ull;
            a.out_ids[rb * a.K + j] = ok ? key_id(ek[c]) : -1;
            if (a.out_scores) a.out_scores[rb * a.K + j] = ok ? key_score(ek[c]) : -INFINITY;
        }
    }
}

// exact fp32 score of the item at sweep position p: the same sequential fmaf chain as the fp32 kernel and
// the oracle (bit-exact).  The user row is read from shared memory; the item row comes from the CTA's
// shared-memory copy of the `n_hot` highest-norm rows when p < n_hot (where nearly every candidate of a
// popularity-skewed catalogue lives: 32 lanes gathering 32 different 256-byte rows from global memory cost 32 L1
// wavefronts per load instruction, from padded shared memory ~6), otherwise from global memory with the 32 floats of
// each block loaded up front (one memory latency per block instead of one per element).
__device__ __forceinline__ uint64_t exact_key(const ScrArgs& a, const float* __restrict__ u_sm, uint32_t p,
                                              const float* __restrict__ e_hot, int n_hot, int hot_pitch)
{
    const int item = a.perm[p];
    float s = a.bias ? a.bias[item] : 0.f;
    const float4* u4 = reinterpret_cast<const float4*>(u_sm);
    if ((int)p < n_hot) {
        const float4* e4 = reinterpret_cast<const float4*>(e_hot + (size_t)p * hot_pitch);
        for (int d = 0; d < a.D / 4; ++d) {
            const float4 ev = e4[d];
            const float4 uv = u4[d];
            s = fmaf(uv.x, ev.x, s);
            s = fmaf(uv.y, ev.y, s);
            s = fmaf(uv.z, ev.z, s);
            s = fmaf(uv.w, ev.w, s);
        }
        return make_key(s, item);
    }
    const float* e = a.Ei + (int64_t)item * a.lde_i;
    if (a.vec4) {
        const float4* e4 = reinterpret_cast<const float4*>(e);
        for (int d0 = 0; d0 < a.D / 4; d0 += 8) {
            float4 ev[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) ev[j] = __ldg(e4 + d0 + j);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 uv = u4[d0 + j];
                s = fmaf(uv.x, ev[j].x, s);
                s = fmaf(uv.y, ev[j].y, s);
                s = fmaf(uv.z, ev[j].z, s);
                s = fmaf(uv.w, ev[j].w, s);
            }
        }
    } else {
        for (int d = 0; d < a.D; ++d) s = fmaf(u_sm[d], e[d], s);
    }
    return make_key(s, item);
}

template <int NSORT>
__device__ __forceinline__ void rescore_sort_write(const ScrArgs& a, const float* __restrict__ u_sm,
                                                   const uint64_t* __restrict__ s_row, int cnt, int lane, int64_t rb,
                                                   const float* __restrict__ e_hot, int n_hot, int hot_pitch)
{
    uint64_t ek[NSORT];
#pragma unroll
    for (int c = 0; c < NSORT; ++c) {
        const int j = c * 32 + lane;
        ek[c] = (j < cnt) ? exact_key(a, u_sm, (uint32_t)s_row[j], e_hot, n_hot, hot_pitch) : 0ull;
    }
    final_sort_write<NSORT>(ek, lane, a, rb);
}

// ---- 3.0 the exact head ---------------------------------------------------------------------------------
struct HeadArgs {
    const float* nb_sorted;     // [I] UNSCALED item norms in sweep order (rounded up), as the sort left them
    const int32_t* perm_sorted; // [I] sweep position -> item id
    uint16_t* hot_pos;          // [I] item id -> head position, 0xFFFF outside the head (memset before head_setup)
    int32_t* hot_id;            // [n_hot] head position -> item id (-1: past the catalogue)
    uint32_t* hot_bits;         // [ceil(I / 32)] bit i: item i is in the head (an L1-resident filter in front of hot_pos)
    uint32_t* row_bits;         // [B_pad][n_hot / 32] bit p of row b: head position p is in b's train history
    float* bound;               // [1] unscaled norm of the first item OUTSIDE the head (0: the head is the catalogue)
    int32_t* on;                // [1] 0: no row can finish in the head (flat norms) -- the head kernel returns at once
    int32_t* n_live;            // [1] ScrArgs::n_live: 0 when the head runs (it appends the rows it leaves), else -1
    int32_t n_hot;
    float margin;               // (1 + rounding slack of the norms and of the fp32 score chains)
};

__global__ void __launch_bounds__(256) score_head_setup_kernel(HeadArgs h, int32_t I, int32_t K)
{
    for (int p = threadIdx.x; p < h.n_hot; p += blockDim.x) {
        const int item = p < I ? h.perm_sorted[p] : -1;
        h.hot_id[p] = item;
        if (item >= 0) {
            h.hot_pos[item] = (uint16_t)p;
            atomicOr(&h.hot_bits[item >> 5], 1u << (item & 31));
        }
    }
    if (threadIdx.x == 0) {
        const float out = h.n_hot < I ? h.nb_sorted[h.n_hot] : 0.f;
        *h.bound = out;
        // s_K <= |u| * (K-th largest norm): a row can only finish if the outside bound is below that
        // (a computed norm below ~1e-15 may have lost squares to underflow and is no longer a bound: head off)
        const bool can = (I <= h.n_hot) || (K <= I && out >= 1e-15f && out * h.margin < h.nb_sorted[K - 1]);
        *h.on = can ? 1 : 0;
        *h.n_live = can ? 0 : -1;
    }
}

// row_bits of every row from the train-history CSR.  ENTRY-parallel: a warp owns kMaskChunk consecutive CSR entries
// whatever rows they belong to (a user with 10^5 interactions costs what 2,000 users with 50 cost; walking the lists row
// by row inside the head kernel made its longest row -- n_items / 4 entries in the synthetic workloads -- the kernel's
// duration).  The rows that bracket the chunk come from a binary search of the row pointers and their pointers are
// staged in shared memory; entries are then loaded 128 at a time (coalesced, independent), filtered by the L1-resident
// hot_bits (about one entry in seven passes), and a hit finds its row by a binary search of the staged pointers.
// (Walking the bracket row by row instead was latency-bound: two dependent loads per 32 entries, 0.33 ms.)
constexpr int kMaskChunk = 1024;
constexpr int kMaskBracket = 160;   // staged row pointers per warp; wider brackets (runs of empty rows) search global memory
__global__ void __launch_bounds__(256, 4)
    score_head_maskbits_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ items, int64_t B, HeadArgs h)
{
    if (*h.on == 0) return;
    __shared__ int32_t rp_all[8][kMaskBracket];   // row pointers of the bracket as offsets from the chunk start (clamped)
    const int lane = threadIdx.x & 31;
    int32_t* rp = rp_all[threadIdx.x >> 5];
    const int64_t base = rowptr[0], end = rowptr[B];
    const int words = h.n_hot >> 5;
    // the host does not know the number of entries: a fixed grid strides over the chunks
    for (int64_t j0 = base + ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * kMaskChunk; j0 < end;
         j0 += (int64_t)gridDim.x * 8 * kMaskChunk) {
        const int64_t j1 = (j0 + kMaskChunk < end) ? j0 + kMaskChunk : end;
        // lane 0: row of entry j0, lane 1: row of entry j1 - 1 (last r with rowptr[r] <= target; empty rows are skipped)
        int64_t lo = 0;
        {
            const int64_t target = lane == 0 ? j0 : j1 - 1;
            int64_t hi = B;
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (rowptr[mid] <= target) lo = mid; else hi = mid;
            }
        }
        const int64_t r0 = __shfl_sync(0xffffffffu, lo, 0), r1 = __shfl_sync(0xffffffffu, lo, 1);
        const int n_br = (int)((r1 - r0 + 2 <= kMaskBracket) ? r1 - r0 + 2 : 0);   // pointers r0 .. r1 + 1, 0: not staged
        __syncwarp();
        for (int q = lane; q < n_br; q += 32) {
            const int64_t off = rowptr[r0 + q] - j0;   // <= 0 for rows that begin before the chunk, > chunk for rows after it
            rp[q] = (int32_t)(off < -1 ? -1 : (off > 2 * kMaskChunk ? 2 * kMaskChunk : off));
        }
        __syncwarp();
        for (int64_t jb = j0; jb < j1; jb += 128) {
            int32_t it[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t j = jb + q * 32 + lane;
                it[q] = j < j1 ? items[j] : -1;
            }
            bool hit[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) hit[q] = it[q] >= 0 && ((h.hot_bits[it[q] >> 5] >> (it[q] & 31)) & 1u);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (!hit[q]) continue;
                const int64_t j = jb + q * 32 + lane;
                const uint32_t p = h.hot_pos[it[q]];
                int64_t row;
                if (n_br > 0) {
                    const int jj = (int)(j - j0);   // 0 .. kMaskChunk - 1
                    int a = 0, b = n_br - 1;        // rp[a] <= jj < rp[b]
                    while (b - a > 1) {
                        const int mid = (a + b) >> 1;
                        if (rp[mid] <= jj) a = mid; else b = mid;
                    }
                    row = r0 + a;
                } else {
                    int64_t a = r0, b = r1 + 1;     // rowptr[a] <= j < rowptr[b]
                    while (b - a > 1) {
                        const int64_t mid = (a + b) >> 1;
                        if (rowptr[mid] <= j) a = mid; else b = mid;
                    }
                    row = a;
                }
                atomicOr(&h.row_bits[row * words + (p >> 5)], 1u << (p & 31));
            }
        }
    }
}

// Bitonic sort of 16 * P keys owned by a HALF-warp, descending: lane `sub` (0 .. 15 inside its half) holds the elements
// sub * P .. sub * P + P - 1, so the log2(P) smallest strides of every merge are register-local compare-exchanges (no
// shuffle, no redundant compare) and only strides >= P cross lanes (xor masks < 16 stay inside the half).
// (Tried first: sorting the 32-bit score words alone and recovering every element's rank by a binary search of the sorted
// words in shared memory: 1.46 -> 1.67 ms, the dependent shared-memory probes cost more than the 64-bit compares save.
// The selection below carries the head position INSIDE a 32-bit key instead.)
template <int P, typename T>
__device__ __forceinline__ void halfwarp_bitonic_sort_desc(T (&k)[P], int sub)
{
    constexpr int N = 16 * P;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= P) {
                const int ls = stride / P;
                const bool take_max = (((sub * P) & size) == 0) == ((sub & ls) == 0);   // size > stride >= P: no r bits
#pragma unroll
                for (int r = 0; r < P; ++r) {
                    const T other = __shfl_xor_sync(0xffffffffu, k[r], ls);
                    k[r] = ((k[r] > other) == take_max) ? k[r] : other;
                }
            } else {
#pragma unroll
                for (int r = 0; r < P; ++r) {
                    if ((r & stride) == 0) {
                        const int r2 = r | stride;
                        const bool desc = size < P ? ((r & size) == 0) : ((((sub * P) | r) & size) == 0);
                        const T x = k[r], y = k[r2];
                        const bool sw = (x > y) != desc;   // descending block: larger key first
                        k[r] = sw ? y : x;
                        k[r2] = sw ? x : y;
                    }
                }
            }
        }
    }
}

// One warp owns R = 32 / NKEY rows at a time.  Phase A (row loads, norms, masked-position bitmaps: the loads of the R
// rows are issued together), phase B (register-tiled exact scores: lane l holds head items l, l + 32, ...), phase C
// (selection: scores go through shared memory to a half-warp-per-row layout; the loop over row pairs is NOT unrolled --
// the first version unrolled the sort network R times, 30,800 SASS instructions, and ran at 27 % issue-active on
// instruction-cache misses).
template <int NKEY>
__global__ void __launch_bounds__(256, 2) score_head_kernel(ScrArgs a, HeadArgs h)
{
    constexpr int R = 32 / NKEY;
    constexpr int NH = 32 * NKEY;
    constexpr int P = NH / 16;                                      // keys per lane in the selection
    if (*h.on == 0) return;
    extern __shared__ __align__(16) float head_smem[];
    const int D = a.D;
    const int pitch = D + 4;                                        // conflict-free 128-bit reads of 8 different rows
    float* e_hot = head_smem;                                       // [NH][D + 4]
    float* u_all = e_hot + (size_t)NH * pitch;                      // [8 warps][R][D]
    float* sc_all = u_all + 8 * R * D;                              // [8 warps][R][NH] exact scores
    uint32_t* bm_all = reinterpret_cast<uint32_t*>(sc_all + 8 * R * NH);  // [8 warps][R][NKEY] masked-position bitmaps
    int32_t* ids = reinterpret_cast<int32_t*>(bm_all + 8 * R * NKEY);     // [NH]
    float* na_all = reinterpret_cast<float*>(ids + NH);             // [8 warps][R] row norms (rounded up)
    uint32_t* so_all = reinterpret_cast<uint32_t*>(na_all + 8 * R); // [8 warps][2][NH] sorted 32-bit keys of the two rows in flight
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int p = threadIdx.x; p < NH; p += blockDim.x) ids[p] = h.hot_id[p];
    for (int idx = threadIdx.x; idx < NH * D; idx += blockDim.x) {
        const int p = idx / D, d = idx - p * D;
        const int item = h.hot_id[p];
        e_hot[(size_t)p * pitch + d] = item >= 0 ? a.Ei[(int64_t)item * a.lde_i + d] : 0.f;
    }
    __syncthreads();
    const float bound = *h.bound;
    const bool whole = a.I <= NH;   // nothing outside the head
    float* u_w = u_all + (size_t)w * R * D;
    float* sc_w = sc_all + (size_t)w * R * NH;
    uint32_t* bm = bm_all + w * R * NKEY;
    float* na_w = na_all + w * R;
    unsigned long long n_done = 0;
    for (int64_t b0 = ((int64_t)blockIdx.x * 8 + w) * R; b0 < a.B; b0 += (int64_t)gridDim.x * 8 * R) {
        // ---- A. rows -> shared memory, norms, masked head positions ----
        // lane q < R fetches row q's source index and mask range (one round trip for all R rows)
        int64_t my_src = 0, my_mlo = 0, my_mhi = 0;   // (the mask range only decides row_force_exact here)
        bool my_cand = false;
        {
            const int64_t b = b0 + (lane < R ? lane : 0);
            const int64_t bc = b < a.B ? b : a.B - 1;               // rows past the end read the last row, never a candidate
            my_src = a.users ? a.users[bc] : bc;
            if (a.mask_rowptr != nullptr) {
                my_mlo = a.mask_rowptr[bc];
                my_mhi = a.mask_rowptr[bc + 1];
            }
            // masked items would surface: the fp32 kernel owns that case (row_force_exact)
            my_cand = lane < R && b < a.B && !((int64_t)a.I - (my_mhi - my_mlo) < (int64_t)a.K);
        }
        unsigned cand_bits = __ballot_sync(0xffffffffu, my_cand);
        // masked head positions of the R rows: R * NKEY = 32 consecutive words (score_head_maskbits_kernel; rows past the
        // end lie inside the padded allocation)
        const uint32_t my_bits = (a.mask_rowptr != nullptr) ? h.row_bits[b0 * NKEY + lane] : 0u;
        float ss[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) ss[rr] = 0.f;
        {
            const float* src[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) src[rr] = a.Eu + __shfl_sync(0xffffffffu, my_src, rr) * a.lde_u;
            for (int d = lane; d < D; d += 32) {
                float v[R];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) v[rr] = src[rr][d];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    u_w[rr * D + d] = v[rr];
                    ss[rr] = fmaf(v[rr], v[rr], ss[rr]);
                }
            }
        }
        bm[lane] = my_bits;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            float t = ss[rr];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            // below 1e-30 the squares underflow and sqrt(ss) is no longer an upper bound of the norm
            if (!(t >= 1e-30f)) cand_bits &= ~(1u << rr);
            if (lane == 0) na_w[rr] = sqrtf(t) * 1.000001f;
        }
        __syncwarp();
        // ---- B. exact scores: acc[rr][r] = the oracle's fmaf chain over d = 0 .. D-1 starting from 0 ----
        {
            float acc[R][NKEY];
#pragma unroll
            for (int rr = 0; rr < R; ++rr)
#pragma unroll
                for (int r = 0; r < NKEY; ++r) acc[rr][r] = 0.f;
#pragma unroll 2
            for (int d4 = 0; d4 < D / 4; ++d4) {
                float4 e[NKEY];
#pragma unroll
                for (int r = 0; r < NKEY; ++r)
                    e[r] = *reinterpret_cast<const float4*>(e_hot + (size_t)(r * 32 + lane) * pitch + 4 * d4);
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const float4 u = *reinterpret_cast<const float4*>(u_w + rr * D + 4 * d4);
#pragma unroll
                    for (int r = 0; r < NKEY; ++r) {
                        float s = acc[rr][r];
                        s = fmaf(u.x, e[r].x, s);
                        s = fmaf(u.y, e[r].y, s);
                        s = fmaf(u.z, e[r].z, s);
                        s = fmaf(u.w, e[r].w, s);
                        acc[rr][r] = s;
                    }
                }
            }
#pragma unroll
            for (int rr = 0; rr < R; ++rr)
#pragma unroll
                for (int r = 0; r < NKEY; ++r) sc_w[rr * NH + r * 32 + lane] = acc[rr][r];
        }
        __syncwarp();
        // ---- C. per row (a half-warp each, two rows at a time): keys, sort, stop test ----
        const int sub = lane & 15, hw = lane >> 4;
#pragma unroll 1
        for (int pair = 0; pair < R / 2; ++pair) {
            const int rr = 2 * pair + hw;
            const int64_t b = b0 + rr;
            // Selection on 32-bit keys: (ordered score word with its PB low bits replaced by PMASK - head position).  A
            // compare-exchange is then two ALU instructions instead of five (64-bit compare + two selects), and the sorted
            // key itself says where the element's exact score and item id are.  The truncation can only misorder scores that
            // agree in their upper 32 - PB bits: if any two NEIGHBOURS among the first K + 1 sorted keys do (a few percent of
            // the rows; exact ties always do), both rows of the warp take the 64-bit (score, id) network below instead.
            constexpr uint32_t PB = (NH == 128) ? 7u : 8u, PMASK = (1u << PB) - 1u;
            const int kth = a.K - 1;   // K <= NH / 2
            const float na = na_w[rr];
            const uint32_t mw = bm[rr * NKEY + ((sub * P) >> 5)] >> ((sub * P) & 31);   // P <= 16 bits of one word
            uint32_t k32[P];
            bool finite = true;
#pragma unroll
            for (int r4 = 0; r4 < P / 4; ++r4) {
                const float4 s4 = *reinterpret_cast<const float4*>(sc_w + rr * NH + sub * P + 4 * r4);
                const int4 i4 = *reinterpret_cast<const int4*>(ids + sub * P + 4 * r4);
                const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
                const int iv[4] = {i4.x, i4.y, i4.z, i4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool ok = iv[q] >= 0 && !((mw >> (4 * r4 + q)) & 1u);
                    finite = finite && (!ok || fabsf(sv[q]) < INFINITY);
                    // (a finite score's ordered word is >= 2^23: a valid key is never 0)
                    k32[4 * r4 + q] = ok ? ((f32_to_ordered(sv[q]) & ~PMASK) | (PMASK - (uint32_t)(sub * P + 4 * r4 + q))) : 0u;
                }
            }
            const unsigned fin_bits = __ballot_sync(0xffffffffu, finite);
            const bool all_finite = ((fin_bits >> (16 * hw)) & 0xFFFFu) == 0xFFFFu;
            halfwarp_bitonic_sort_desc<P, uint32_t>(k32, sub);
            uint32_t* so32 = so_all + (size_t)(w * 2 + hw) * NH;
#pragma unroll
            for (int r = 0; r < P; ++r) so32[sub * P + r] = k32[r];
            bool tie = false;
            {
                const uint32_t next_first = __shfl_down_sync(0xffffffffu, k32[0], 1);   // rank (sub + 1) * P
#pragma unroll
                for (int r = 0; r < P; ++r) {
                    const uint32_t nxt = (r + 1 < P) ? k32[r + 1] : next_first;
                    tie = tie || (sub * P + r <= kth && k32[r] != 0u && ((k32[r] ^ nxt) >> PB) == 0u);   // ranks (j, j + 1), j < K
                }
            }
            const bool any_tie = __ballot_sync(0xffffffffu, tie) != 0u;   // warp-uniform: both rows switch together
            __syncwarp();
            bool done;
            if (!any_tie) {
                const uint32_t kk = so32[kth];   // 0: fewer than K unmasked head items
                const float s_k = sc_w[rr * NH + (PMASK - (kk & PMASK))];   // the K-th best EXACT score
                // every item outside the head scores (as computed in fp32) at most |u| * bound * margin
                done = b < a.B && ((cand_bits >> rr) & 1u) && all_finite && kk != 0u && na < INFINITY &&
                       (whole || na * bound * h.margin < s_k);
                if (done) {
#pragma unroll
                    for (int r = 0; r < P; ++r) {
                        const int j = sub * P + r;
                        if (j < a.K) {
                            const uint32_t pos = PMASK - (k32[r] & PMASK);
                            a.out_ids[b * a.K + j] = ids[pos];
                            if (a.out_scores) a.out_scores[b * a.K + j] = sc_w[rr * NH + pos] + 0.0f;   // (-0 -> +0 like key_score)
                        }
                    }
                }
            } else {
                uint64_t k[P];
#pragma unroll
                for (int r4 = 0; r4 < P / 4; ++r4) {
                    const float4 s4 = *reinterpret_cast<const float4*>(sc_w + rr * NH + sub * P + 4 * r4);
                    const int4 i4 = *reinterpret_cast<const int4*>(ids + sub * P + 4 * r4);
                    const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
                    const int iv[4] = {i4.x, i4.y, i4.z, i4.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const bool ok = iv[q] >= 0 && !((mw >> (4 * r4 + q)) & 1u);
                        k[4 * r4 + q] = ok ? make_key(sv[q], iv[q]) : 0ull;
                    }
                }
                halfwarp_bitonic_sort_desc<P, uint64_t>(k, sub);
                // the K-th key's score word through shared memory (a register array indexed by kth % P would live in local
                // memory); 0 = fewer than K unmasked head items (no finite score orders to 0)
                __syncwarp();
#pragma unroll
                for (int r = 0; r < P; ++r) so32[sub * P + r] = (uint32_t)(k[r] >> 32);
                __syncwarp();
                const uint32_t kk = so32[kth];
                done = b < a.B && ((cand_bits >> rr) & 1u) && all_finite && kk != 0u && na < INFINITY &&
                       (whole || na * bound * h.margin < ordered_to_f32(kk));
                if (done) {
#pragma unroll
                    for (int r = 0; r < P; ++r) {
                        const int j = sub * P + r;
                        if (j < a.K) {
                            a.out_ids[b * a.K + j] = key_id(k[r]);
                            if (a.out_scores) a.out_scores[b * a.K + j] = key_score(k[r]);
                        }
                    }
                }
                if (a.stats != nullptr && lane == 0) atomicAdd(&a.stats[4], 1ull);   // pairs that needed the 64-bit network
            }
            if (done && sub == 0) n_done += 1;
            if (sub == 0 && b < a.B) {
                a.row_done[b] = done ? 1 : 0;
                if (!done) a.live_rows[atomicAdd(a.n_live, 1)] = (int32_t)b;
            }
        }
        __syncwarp();
    }
    if (a.stats != nullptr && n_done) atomicAdd(&a.stats[7], n_done);
}

static size_t head_smem_bytes(int nkey, int32_t D)
{
    const int nh = 32 * nkey, r = 32 / nkey;
    return (size_t)nh * (D + 4) * 4 + (size_t)8 * r * D * 4 + (size_t)8 * r * nh * 4 + (size_t)8 * r * nkey * 4 + (size_t)nh * 4 +
           (size_t)8 * r * 4 + (size_t)8 * 2 * nh * 4;
}

// ---- 3a. the sweep kernel (TMA + tcgen05 + screening epilogue) ----------------------------------------
// phase 0: every 256-user group sweeps tiles [0, first_check) and stores its row state.
// phase 1: groups whose rows still need tiles (group_need, written by the checkpoint kernel) continue from
//          first_check with in-kernel stop checks at doubling tile counts; groups are handed out dynamically.
template <int NPL, bool HAS_BIAS>
__global__ void __launch_bounds__(kScrThreads, 1)
    score_screen_sweep_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, ScrArgs a,
                              int phase)
{
    constexpr int CAP = 32 * NPL;
    constexpr int TRIG = CAP - 32;

    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem_raw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int n_atoms = a.D / 64;
    const int n_stages = a.n_stages;
    uint8_t* sm_a = smem_raw;                                         // [kUT][n_atoms][16 KB]
    uint8_t* sm_b = sm_a + (size_t)kUT * n_atoms * kAtomBytes;        // [n_stages][16 KB]
    float* sm_strip = reinterpret_cast<float*>(sm_b + (size_t)n_stages * kAtomBytes);  // [kEpiWarps][32][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_strip + kEpiWarps * 1024);
    uint64_t* full_b = bars;                                  // [kScrMaxStages]
    uint64_t* empty_b = bars + kScrMaxStages;                 // [kScrMaxStages]
    uint64_t* tmem_full = bars + 2 * kScrMaxStages;           // [kUT][2]
    uint64_t* tmem_empty = bars + 2 * kScrMaxStages + 4;      // [kUT][2]
    uint64_t* a_full = bars + 2 * kScrMaxStages + 8;
    uint64_t* a_empty = bars + 2 * kScrMaxStages + 9;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kScrMaxStages + 10);
    int* sm_need = reinterpret_cast<int*>(bars + 2 * kScrMaxStages + 11);  // [0] tiles still needed, [1] next group

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_groups = (a.B + kRowsPerCta - 1) / kRowsPerCta;
    const int n_itiles = (a.I + kSN - 1) / kSN;
    const int T0 = a.first_check < n_itiles ? a.first_check : n_itiles;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&empty_b[s], 1);
        }
        for (int s = 0; s < 2 * kUT; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 128);
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        sm_need[0] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    // per-role persistent state (counters run across segments and user groups)
    uint32_t g = 0, t = 0, a_phase = 0;            // B'-atom counter, item-tile counter, A' barrier phase
    const int ew = warp - 2;
    const int u = (ew >> 2) & 1;
    const int q = warp & 3;
    const int row_cta = u * kSM + q * 32 + lane;  // epilogue: row inside the CTA's 256 users
    const int warp_row0 = u * kSM + q * 32;
    float* strip = sm_strip + (ew < 0 ? 0 : ew) * 1024;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const float bias_abs_max = (HAS_BIAS && a.biasmax_bits) ? __uint_as_float(*a.biasmax_bits) : 0.f;
    unsigned long long st_slow = 0, st_app = 0, st_tiles = 0;

    int ug_static = blockIdx.x;
    while (true) {
        // ---- next group ----
        int ug;
        if (phase == 0) {
            ug = ug_static;
            ug_static += gridDim.x;
        } else {
            if (threadIdx.x == 0) {
                int gi;
                do {
                    gi = atomicAdd(a.work_counter, 1);
                } while (gi < n_groups && a.group_need[gi] <= T0 && a.debug != 3);
                sm_need[1] = gi;
            }
            __syncthreads();
            ug = sm_need[1];
            __syncthreads();
        }
        if (ug >= n_groups) break;
        if (phase == 0) {
            // groups whose 256 rows were all settled by the exact head are not swept at all
            const int64_t bb = (int64_t)ug * kRowsPerCta + row_cta;
            const int live = (warp >= 2 && bb < a.B && !a.row_done[bb]) ? 1 : 0;
            if (__syncthreads_or(live) == 0) continue;
        }
        int it = 0, it_end = T0, next_check = 0x7fffffff;
        if (phase == 1) {
            it = T0;
            it_end = a.debug == 3 ? n_itiles : min(a.group_need[ug], n_itiles);
            next_check = 2 * T0;
            if (it >= it_end) continue;  // (debug == 3 with a one-segment catalogue)
        }

        // ---- group prologue ----
        int cnt = 0, chk = 0;
        float L = INFINITY;
        RowConst rc = {0.f, 0.f, 0.f, 0.f};
        int64_t mlo = 0, mhi = 0;
        const int64_t b = (int64_t)ug * kRowsPerCta + row_cta;
        uint64_t* cta_slots = a.slots + (int64_t)ug * kRowsPerCta * CAP;
        if (warp == 0) {
            if (lane == 0) {
                mbar_wait(a_empty, a_phase ^ 1);
                mbar_expect_tx(a_full, (uint32_t)(kUT * n_atoms) * kAtomBytes);
                for (int uu = 0; uu < kUT; ++uu)
                    for (int k = 0; k < n_atoms; ++k)
                        tma_load_2d(sm_a + (size_t)(uu * n_atoms + k) * kAtomBytes, &map_a, a_full, k * 64,
                                    (ug * kUT + uu) * kSM);
                a_phase ^= 1;
            }
        } else if (warp == 1) {
            if (lane == 0) {
                mbar_wait(a_full, a_phase);
                a_phase ^= 1;
            }
        } else if (b < a.B) {
            rc = a.row_const[b];
            const bool fe = row_force_exact(a, b, mlo, mhi);
            if (phase == 0) {
                // +inf: collect nothing (settled by the head, or redone on the fp32 path)
                L = (fe || !(rc.na < INFINITY) || a.row_done[b]) ? INFINITY : -INFINITY;
            } else {
                cnt = a.row_cnt[b];
                chk = a.row_chk[b];
                L = a.row_L[b];
            }
        }

        // ---- sweep in segments [it, seg_end); the stop tile is re-agreed at doubling checkpoints ----
        while (it < it_end) {
            const int seg_end = it_end < next_check ? it_end : next_check;
            if (warp == 0) {
                // ===== TMA producer =====
                if (lane == 0) {
                    for (int i2 = it; i2 < seg_end; ++i2) {
                        for (int k = 0; k < n_atoms; ++k, ++g) {
                            const uint32_t stage = g % (uint32_t)n_stages, ph = (g / (uint32_t)n_stages) & 1u;
                            mbar_wait(&empty_b[stage], ph ^ 1);
                            mbar_expect_tx(&full_b[stage], kAtomBytes);
                            tma_load_2d(sm_b + (size_t)stage * kAtomBytes, &map_b, &full_b[stage], k * 64, i2 * kSN);
                        }
                    }
                }
            } else if (warp == 1) {
                // ===== MMA issuer (one elected thread) =====
                if (lane == 0) {
                    const uint32_t idesc = umma_idesc_f16(kSM, kSN);
                    for (int i2 = it; i2 < seg_end; ++i2, ++t) {
                        const uint32_t acc = t & 1u, acc_phase = (t >> 1) & 1u;
                        for (int k = 0; k < n_atoms; ++k) {
                            const uint32_t gi = g + (uint32_t)k;
                            mbar_wait(&full_b[gi % (uint32_t)n_stages], (gi / (uint32_t)n_stages) & 1u);
                        }
                        tc_fence_after();
                        for (int uu = 0; uu < kUT; ++uu) {
                            mbar_wait(&tmem_empty[uu * 2 + acc], acc_phase ^ 1);
                            tc_fence_after();
                            const uint32_t d_tmem = tmem_base + (uint32_t)(uu * 2 + acc) * kSN;
                            for (int k = 0; k < n_atoms; ++k) {
                                const uint32_t a_base = smem_u32(sm_a + (size_t)(uu * n_atoms + k) * kAtomBytes);
                                const uint32_t b_base =
                                    smem_u32(sm_b + (size_t)((g + (uint32_t)k) % (uint32_t)n_stages) * kAtomBytes);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    tc_mma_bf16(d_tmem, umma_desc_sw128(a_base + kk * 32),
                                                umma_desc_sw128(b_base + kk * 32), idesc, (k | kk) ? 1u : 0u);
                            }
                            tc_commit(&tmem_full[uu * 2 + acc]);
                        }
                        for (int k = 0; k < n_atoms; ++k) tc_commit(&empty_b[(g + (uint32_t)k) % (uint32_t)n_stages]);
                        g += (uint32_t)n_atoms;
                    }
                }
            } else {
                // ===== epilogue: thread = user row; warp % 4 = TMEM lane quarter =====
                for (int i2 = it; i2 < seg_end; ++i2, ++t) {
                    const uint32_t acc = t & 1u, acc_phase = (t >> 1) & 1u;
                    const int i0 = i2 * kSN;
                    st_tiles += 1;
                    mbar_wait(&tmem_full[u * 2 + acc], acc_phase);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * 2 + acc) * kSN;
#pragma unroll 1
                    for (int c0 = 0; c0 < kSN; c0 += 32) {
                        uint32_t v[32];
                        tc_ld_32x32(taddr + c0, v);
                        tc_ld_wait();
                        if (a.debug == 1) continue;
                        const int cbase = i0 + c0;
                        const float nbi = a.nb[cbase + lane];  // this lane's column (nb is padded to the tile grid)
                        if (HAS_BIAS) {
                            const int item = a.perm[cbase + lane];
                            const float bl = (item >= 0) ? a.bias[item] : 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                v[j] = __float_as_uint(fmaf(__shfl_sync(0xffffffffu, bl, j), rc.sc, __uint_as_float(v[j])));
                        }
                        float m0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
                        float m1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
                        for (int j = 4; j < 32; j += 4) {
                            m0 = fmaxf(fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), m0);
                            m1 = fmaxf(fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), m1);
                        }
                        // ub of the chunk's best column >= L ?   (norms descend: the chunk's largest norm is column 0's)
                        const float nbc = __shfl_sync(0xffffffffu, nbi, 0);
                        const bool flag = fmaxf(m0, m1) + fmaf(rc.ce, nbc, rc.ab) >= L;
                        const unsigned f = __ballot_sync(0xffffffffu, flag);
                        if (f == 0u || a.debug == 2) continue;
                        // ---- slow path (warp-cooperative append; lane = column) ----
                        st_slow += 1;
                        if (__popc(f) >= 8) {
                            // ---- dense form (cold thresholds: the first tiles of a sweep, small catalogues) ----
                            // Most rows of the warp have something to append: the cooperative loop below would run once per
                            // flagged row (~60 instructions each).  Here every flagged thread appends its own row's columns
                            // straight from its registers, in column (= sweep) order.  The column's own norm is replaced by
                            // the chunk's largest (nbc): a looser test only appends a few more candidates.
                            if (flag) {
                                const float e = fmaf(rc.ce, nbc, rc.ab);
                                uint64_t* dst = cta_slots + (int64_t)(warp_row0 + lane) * CAP + cnt;
                                const int ncol = a.I - cbase;   // columns >= ncol are padding
                                int n = 0;
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const float sj = __uint_as_float(v[j]);
                                    if (j < ncol && sj + e >= L) {
                                        dst[n] = ((uint64_t)f32_to_ordered(sj) << 32) | (uint32_t)(cbase + j);
                                        ++n;
                                    }
                                }
                                cnt += n;
                                st_app += 32ull * (unsigned)n;   // (the cooperative form counts every append on all 32 lanes)
                            }
                            __syncwarp();
                        } else {
                        if (flag) {
#pragma unroll
                            for (int j8 = 0; j8 < 8; ++j8) {
                                float4 w;
                                w.x = __uint_as_float(v[4 * j8 + 0]);
                                w.y = __uint_as_float(v[4 * j8 + 1]);
                                w.z = __uint_as_float(v[4 * j8 + 2]);
                                w.w = __uint_as_float(v[4 * j8 + 3]);
                                *reinterpret_cast<float4*>(strip + lane * 32 + ((j8 ^ (lane & 7)) << 2)) = w;
                            }
                        }
                        const bool col_ok = cbase + lane < a.I;
                        __syncwarp();
                        unsigned ff = f;
                        while (ff) {
                            const int l = __ffs(ff) - 1;
                            ff &= ff - 1;
                            const float s = strip[l * 32 + (((lane >> 2) ^ (l & 7)) << 2) + (lane & 3)];
                            const float L_l = __shfl_sync(0xffffffffu, L, l);
                            const float ce_l = __shfl_sync(0xffffffffu, rc.ce, l);
                            const float ab_l = __shfl_sync(0xffffffffu, rc.ab, l);
                            const int cnt_l = __shfl_sync(0xffffffffu, cnt, l);
                            const bool take = (s + fmaf(ce_l, nbi, ab_l) >= L_l) && col_ok;
                            const unsigned tb = __ballot_sync(0xffffffffu, take);
                            if (take)
                                cta_slots[(int64_t)(warp_row0 + l) * CAP + cnt_l + __popc(tb & lt_mask)] =
                                    ((uint64_t)f32_to_ordered(s) << 32) | (uint32_t)(cbase + lane);
                            const int n = __popc(tb);
                            if (lane == l) cnt += n;
                            st_app += n;
                        }
                        __syncwarp();
                        }
                        unsigned need = __ballot_sync(0xffffffffu, cnt >= TRIG);
                        while (need) {
                            const int l = __ffs(need) - 1;
                            need &= need - 1;
                            const PruneResult pr = screen_prune<NPL>(
                                cta_slots + (int64_t)(warp_row0 + l) * CAP, __shfl_sync(0xffffffffu, cnt, l),
                                __shfl_sync(0xffffffffu, chk, l), __shfl_sync(0xffffffffu, rc.ce, l),
                                __shfl_sync(0xffffffffu, rc.ab, l), a.nb, a.perm, a.mask_items,
                                __shfl_sync(0xffffffffu, mlo, l), __shfl_sync(0xffffffffu, mhi, l), a.K, lane, a.stats);
                            if (lane == l) {
                                cnt = pr.cnt;
                                chk = pr.cnt;
                                L = fmaxf(L, pr.lbK);
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&tmem_empty[u * 2 + acc]);
                }
            }
            it = seg_end;
            if (it < it_end) {
                // ---- in-kernel checkpoint (phase 1 only): how many tiles do this CTA's rows still need? ----
                if (warp >= 2 && a.debug == 0) {
                    // tighten every live row's L first (a prune needs >= K keys to say anything)
                    for (int l = 0; l < 32; ++l) {
                        const int cnt_l = __shfl_sync(0xffffffffu, cnt, l);
                        if (cnt_l < a.K || cnt_l == __shfl_sync(0xffffffffu, chk, l)) continue;  // warp-uniform
                        const PruneResult pr = screen_prune<NPL>(
                            cta_slots + (int64_t)(warp_row0 + l) * CAP, cnt_l, __shfl_sync(0xffffffffu, chk, l),
                            __shfl_sync(0xffffffffu, rc.ce, l), __shfl_sync(0xffffffffu, rc.ab, l), a.nb, a.perm,
                            a.mask_items, __shfl_sync(0xffffffffu, mlo, l), __shfl_sync(0xffffffffu, mhi, l), a.K, lane,
                            a.stats);
                        if (lane == l) {
                            cnt = pr.cnt;
                            chk = pr.cnt;
                            L = fmaxf(L, pr.lbK);
                        }
                    }
                    int need_tiles = tiles_needed(a, rc, L, bias_abs_max, it, it_end);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) need_tiles = max(need_tiles, __shfl_xor_sync(0xffffffffu, need_tiles, o));
                    if (lane == 0) atomicMax(&sm_need[0], need_tiles);
                } else if (warp >= 2 && lane == 0) {
                    atomicMax(&sm_need[0], it_end);
                }
                __syncthreads();
                const int agreed = sm_need[0];
                __syncthreads();
                if (threadIdx.x == 0) sm_need[0] = 0;
                if (agreed < it_end) it_end = agreed > it ? agreed : it;
                if (a.debug == 3) it_end = n_itiles;
                next_check = next_check * 2;
            }
        }

        // ---- group epilogue: hand the row state to the next kernel ----
        if (warp == 1) {
            if (lane == 0) tc_commit(a_empty);  // A' tiles free once every MMA of this group retired
        } else if (warp >= 2 && b < a.B) {
            a.row_cnt[b] = cnt;
            a.row_chk[b] = chk;
            a.row_L[b] = L;
        }
    }
    if (warp >= 2 && a.stats != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) st_app += __shfl_xor_sync(0xffffffffu, st_app, o);
        if (lane == 0) {
            atomicAdd(&a.stats[0], st_slow);
            atomicAdd(&a.stats[1], st_app / 32);  // every lane counted each append
            if (ew == 0) atomicAdd(&a.stats[6], st_tiles);  // item tiles swept, summed over user groups
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- 3b. checkpoint after phase 0: one warp per row at full occupancy --------------------------------
// Applies the train-history mask to the row's first candidates, computes L and the number of item tiles the
// row still needs; the maximum over each 256-user group decides whether phase 1 touches the group at all.
template <int NPL>
__global__ void __launch_bounds__(256, (NPL == 8) ? 4 : ((NPL == 16) ? 3 : 2))
    score_screen_checkpoint_kernel(ScrArgs a)
{
    constexpr int CAP = 32 * NPL;
    const int lane = threadIdx.x & 31;
    const int nl = *a.n_live;                      // >= 0: only the rows the exact head left over
    const int64_t n_work = nl >= 0 ? (int64_t)nl : (int64_t)a.B;
    const int n_itiles = (a.I + kSN - 1) / kSN;
    const int T0 = a.first_check < n_itiles ? a.first_check : n_itiles;
    for (int64_t idx = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); idx < n_work;
         idx += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t b = nl >= 0 ? (int64_t)a.live_rows[idx] : idx;
    const RowConst rc = a.row_const[b];
    int64_t mlo, mhi;
    row_force_exact(a, b, mlo, mhi);
    int cnt = a.row_cnt[b];
    float L = a.row_L[b];
    if (cnt >= a.K) {
        const PruneResult pr = screen_prune<NPL>(a.slots + b * CAP, cnt, a.row_chk[b], rc.ce, rc.ab, a.nb, a.perm,
                                                 a.mask_items, mlo, mhi, a.K, lane, a.stats);
        cnt = pr.cnt;
        L = fmaxf(L, pr.lbK);
        if (lane == 0) {
            a.row_cnt[b] = cnt;
            a.row_chk[b] = cnt;
            a.row_L[b] = L;
        }
    }
    if (lane == 0) {
        const float bias_abs_max = a.biasmax_bits ? __uint_as_float(*a.biasmax_bits) : 0.f;
        const int need = tiles_needed(a, rc, L, bias_abs_max, T0, n_itiles);
        if (need > T0) atomicMax(&a.group_need[b / kRowsPerCta], need);
    }
    }
}

// ---- 3c. finalisation: one warp per row, persistent CTAs ---------------------------------------------
// Last prune if anything was appended since the previous one, exact fp32 re-score of the survivors, sort by
// (score desc, id asc), write the top K; undecidable rows are queued for the fp32 kernel.  Each CTA first copies the
// `n_hot` highest-norm item rows (sweep positions 0 .. n_hot-1) into padded shared memory.
template <int NPL>
__global__ void __launch_bounds__(256, (NPL == 8) ? 3 : 2)
    score_screen_finalize_kernel(ScrArgs a, int n_hot)
{
    constexpr int CAP = 32 * NPL;
    extern __shared__ __align__(16) float fin_smem[];
    const int hot_pitch = a.D + 4;                      // +16 B: conflict-free 128-bit reads of 8 different rows
    float* u_all = fin_smem;                            // [8][D]
    float* e_hot = fin_smem + 8 * a.D;                  // [n_hot][D + 4]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nl = *a.n_live;                      // >= 0: only the rows the exact head left over
    const int64_t n_work = nl >= 0 ? (int64_t)nl : (int64_t)a.B;
    if ((int64_t)blockIdx.x * 8 >= n_work) return;   // no row for this CTA: skip the staging as well
    for (int idx = threadIdx.x; idx < n_hot * (a.D / 4); idx += blockDim.x) {
        const int p = idx / (a.D / 4), j = idx - p * (a.D / 4);
        const int item = a.perm[p];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item >= 0) v = __ldg(reinterpret_cast<const float4*>(a.Ei + (int64_t)item * a.lde_i) + j);
        *reinterpret_cast<float4*>(e_hot + (size_t)p * hot_pitch + 4 * j) = v;
    }
    __syncthreads();
    float* u_sm = u_all + w * a.D;
    for (int64_t idx = (int64_t)blockIdx.x * 8 + w; idx < n_work; idx += (int64_t)gridDim.x * 8) {
        const int64_t b = nl >= 0 ? (int64_t)a.live_rows[idx] : idx;
        const RowConst rc = a.row_const[b];
        int64_t mlo, mhi;
        const bool fe = row_force_exact(a, b, mlo, mhi);
        int cnt = a.row_cnt[b];
        const int chk = a.row_chk[b];
        uint64_t* s_row = a.slots + b * CAP;
        bool fallback = fe || cnt < a.K;
        if (!fallback && cnt != chk) {  // appended since the last prune
            const PruneResult pr = screen_prune<NPL>(s_row, cnt, chk, rc.ce, rc.ab, a.nb, a.perm, a.mask_items, mlo, mhi,
                                                     a.K, lane, a.stats);
            cnt = pr.cnt;
            fallback = cnt < a.K;  // overflow (-1) or too few unmasked candidates
        }
        if (fallback) {
            if (lane == 0) a.fallback_rows[atomicAdd(a.fallback_count, 1)] = (int32_t)b;
            continue;
        }
        const float* urow = a.Eu + (a.users ? a.users[b] : b) * a.lde_u;
        __syncwarp();
        for (int d = lane; d < a.D; d += 32) u_sm[d] = urow[d];
        __syncwarp();
        if (a.stats && lane == 0) atomicAdd(&a.stats[5], (unsigned long long)cnt);
        if (cnt <= 64)
            rescore_sort_write<2>(a, u_sm, s_row, cnt, lane, b, e_hot, n_hot, hot_pitch);
        else if (cnt <= 128)
            rescore_sort_write<4>(a, u_sm, s_row, cnt, lane, b, e_hot, n_hot, hot_pitch);
        else
            rescore_sort_write<NPL>(a, u_sm, s_row, cnt, lane, b, e_hot, n_hot, hot_pitch);
    }
}

// ---- host side ------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn scr_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [rows, kd] fp16 row-major, boxes of 64 (K) x 128 rows, 128-byte swizzle
static bool scr_make_map(CUtensorMap* m, void* base, int64_t rows, int64_t kd)
{
    EncodeTiledFn fn = scr_encode_fn();
    if (fn == nullptr) return false;
    cuuint64_t dims[2] = {(cuuint64_t)kd, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kd * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int scr_npl(int32_t K) { return K <= 64 ? 8 : (K <= 128 ? 16 : 32); }

static size_t scr_smem_fixed(int32_t D)
{
    return (size_t)kUT * (D / 64) * kAtomBytes + (size_t)kEpiWarps * 4096 + (2 * kScrMaxStages + 12) * sizeof(uint64_t) + 1024;
}
static int scr_stages(int32_t D)
{
    const int64_t room = (int64_t)227 * 1024 - (int64_t)scr_smem_fixed(D);
    const int s = (int)(room / kAtomBytes);
    return s > kScrMaxStages ? kScrMaxStages : s;
}

bool score_screen_supported(int32_t D, int32_t K)
{
    return D % 64 == 0 && D >= 64 && D <= 256 && K >= 1 && K <= 256 && scr_stages(D) >= D / 64;
}

static int scr_grid(int32_t B)
{
    const int groups = (B + kRowsPerCta - 1) / kRowsPerCta;
    return groups < sm_count() ? groups : sm_count();
}

static size_t scr_sort_temp_bytes(int32_t I)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                              (const int32_t*)nullptr, (int32_t*)nullptr, I);
    return bytes;
}

struct ScrLayout {
    int64_t a_h, b_h, row_const, nb, perm, nb_raw, ident, sort_tmp, misc, fallback, row_cnt, row_chk, row_L, group_need,
        slots, simt, hot_pos, hot_id, row_done, live_rows, hot_bits, row_bits, total;
    int64_t sort_tmp_bytes;
    int32_t b_pad, i_pad;
};

static ScrLayout scr_layout(int32_t B, int32_t I, int32_t D, int32_t K)
{
    ScrLayout L;
    L.b_pad = (B + kRowsPerCta - 1) / kRowsPerCta * kRowsPerCta;
    L.i_pad = (I + kSN - 1) / kSN * kSN;
    const int cap = 32 * scr_npl(K);
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        const int64_t o = off;
        off += align_up(bytes, 256);
        return o;
    };
    L.a_h = take((int64_t)L.b_pad * D * 2);
    L.b_h = take((int64_t)L.i_pad * D * 2);
    L.row_const = take((int64_t)L.b_pad * sizeof(RowConst));
    L.nb = take((int64_t)L.i_pad * 4);
    L.perm = take((int64_t)L.i_pad * 4);
    L.nb_raw = take((int64_t)L.i_pad * 4);
    L.ident = take((int64_t)L.i_pad * 4);
    L.sort_tmp_bytes = (int64_t)scr_sort_temp_bytes(I);
    L.sort_tmp = take(L.sort_tmp_bytes);
    L.misc = take(256);  // u32 [0] item absmax bits, [1] fallback count, [3] bias absmax, [4] group scheduler; u64 stats at +64;
                         // f32 [32] head bound, i32 [33] head on, i32 [34] live-row count (-1: all rows)
    L.fallback = take((int64_t)B * 4);
    L.row_cnt = take((int64_t)L.b_pad * 4);
    L.row_chk = take((int64_t)L.b_pad * 4);
    L.row_L = take((int64_t)L.b_pad * 4);
    L.group_need = take((int64_t)(L.b_pad / kRowsPerCta) * 4);
    L.slots = take((int64_t)L.b_pad * cap * 8);
    L.simt = take(score_simt_workspace_bytes(B, K));
    L.hot_pos = take((int64_t)I * 2);
    L.hot_id = take(256 * 4);
    L.row_done = take((int64_t)L.b_pad);
    L.live_rows = take((int64_t)B * 4);
    L.hot_bits = take(((int64_t)I + 31) / 32 * 4);
    L.row_bits = take((int64_t)L.b_pad * (K <= 64 ? 4 : 8) * 4);
    L.total = off;
    return L;
}

int64_t score_screen_workspace_bytes(int32_t B, int32_t I, int32_t D, int32_t K) { return scr_layout(B, I, D, K).total; }

int score_screen_fallback_count(const void* workspace, int32_t B, int32_t I, int32_t D, int32_t K, int32_t* count_host,
                                cudaStream_t st)
{
    const ScrLayout L = scr_layout(B, I, D, K);
    GMR_CHECK_CUDA(cudaMemcpyAsync(count_host, (const uint8_t*)workspace + L.misc + 4, sizeof(int32_t),
                                   cudaMemcpyDeviceToHost, st));
    GMR_CHECK_CUDA(cudaStreamSynchronize(st));
    return GMR_OK;
}

// diagnostics of the last call with GMR_SCREEN_STATS=1: [0] slow-path chunks, [1] appends, [2] pool prunes,
// [3] exact prunes, [4] row pairs of the exact head that needed the 64-bit network, [5] exactly re-scored candidates,
// [6] item tiles swept (summed over 256-user groups),
// [7] rows settled by the exact head
int score_screen_stats(const void* workspace, int32_t B, int32_t I, int32_t D, int32_t K, uint64_t* stats_host,
                       cudaStream_t st)
{
    const ScrLayout L = scr_layout(B, I, D, K);
    GMR_CHECK_CUDA(cudaMemcpyAsync(stats_host, (const uint8_t*)workspace + L.misc + 64, kStatSlots * sizeof(uint64_t),
                                   cudaMemcpyDeviceToHost, st));
    GMR_CHECK_CUDA(cudaStreamSynchronize(st));
    return GMR_OK;
}

int score_topk_screen_launch(const float* Eu, int64_t lde_u, const int64_t* users, int32_t B, const float* Ei,
                             int64_t lde_i, const float* bias, int32_t I, int32_t D, const int64_t* mask_rowptr,
                             const int32_t* mask_items, int32_t K, int32_t* out_ids, float* out_scores, void* workspace,
                             int64_t workspace_bytes, cudaStream_t st)
{
    const ScrLayout L = scr_layout(B, I, D, K);
    if (workspace_bytes < L.total) {
        set_error("score_topk_screen: workspace of %lld bytes required, %lld given", (long long)L.total,
                  (long long)workspace_bytes);
        return GMR_ERR_WORKSPACE;
    }
    if ((uintptr_t)workspace % 256 != 0) {
        set_error("score_topk_screen: workspace must be 256-byte aligned");
        return GMR_ERR_INVALID;
    }
    const int n_stages = scr_stages(D);
    const size_t smem = scr_smem_fixed(D) + (size_t)n_stages * kAtomBytes;
    uint8_t* ws = (uint8_t*)workspace;
    __half* a_h = (__half*)(ws + L.a_h);
    __half* b_h = (__half*)(ws + L.b_h);
    RowConst* row_const = (RowConst*)(ws + L.row_const);
    float* nb = (float*)(ws + L.nb);
    int32_t* perm = (int32_t*)(ws + L.perm);
    uint32_t* nb_raw = (uint32_t*)(ws + L.nb_raw);
    int32_t* ident = (int32_t*)(ws + L.ident);
    uint32_t* misc = (uint32_t*)(ws + L.misc);
    int32_t* fallback = (int32_t*)(ws + L.fallback);

    GMR_CHECK_CUDA(cudaMemsetAsync(misc, 0, 256, st));
    const int wpb = 8;
    if (bias != nullptr) {
        const int64_t warps = ((int64_t)I + 3) / 4;  // ~4 rows per warp
        int grid = (int)((warps + wpb - 1) / wpb);
        if (grid > 16 * sm_count()) grid = 16 * sm_count();
        if (grid < 1) grid = 1;
        absmax_kernel<<<grid, wpb * 32, 0, st>>>(bias, 1, I, 1, misc + 3);
        GMR_LAUNCH_CHECK();
    }
    const bool items_d64 = D == 64 && lde_i % 4 == 0 && (uintptr_t)Ei % 16 == 0;
    if (items_d64)
        item_norms_d64_kernel<<<(I + 8 * wpb - 1) / (8 * wpb), wpb * 32, 0, st>>>(Ei, lde_i, I, misc, nb_raw, ident);
    else
        item_norms_kernel<<<(I + wpb - 1) / wpb, wpb * 32, 0, st>>>(Ei, lde_i, I, D, misc, nb_raw, ident);
    GMR_LAUNCH_CHECK();
    {
        size_t tmp = (size_t)L.sort_tmp_bytes;
        GMR_CHECK_CUDA(cub::DeviceRadixSort::SortPairsDescending(ws + L.sort_tmp, tmp, nb_raw, (uint32_t*)nb, ident, perm, I,
                                                                 0, 32, st));
    }
    ScrArgs a;
    a.Eu = Eu; a.lde_u = lde_u; a.users = users; a.B = B; a.Ei = Ei; a.lde_i = lde_i; a.bias = bias; a.I = I; a.D = D;
    a.mask_rowptr = mask_rowptr; a.mask_items = mask_items; a.K = K; a.out_ids = out_ids; a.out_scores = out_scores;
    a.slots = (uint64_t*)(ws + L.slots); a.row_const = row_const; a.nb = nb; a.perm = perm;
    a.biasmax_bits = bias ? misc + 3 : nullptr;
    a.row_cnt = (int32_t*)(ws + L.row_cnt); a.row_chk = (int32_t*)(ws + L.row_chk); a.row_L = (float*)(ws + L.row_L);
    a.group_need = (int32_t*)(ws + L.group_need); a.work_counter = (int32_t*)(misc + 4);
    a.fallback_rows = fallback; a.fallback_count = (int32_t*)(misc + 1); a.n_stages = n_stages;
    a.vec4 = ((lde_u % 4 == 0) && (lde_i % 4 == 0) && ((uintptr_t)Eu % 16 == 0) && ((uintptr_t)Ei % 16 == 0)) ? 1 : 0;
    a.first_check = (2 * K + kSN - 1) / kSN;  // enough tiles for ~2K candidates before the first stop check
    a.stats = getenv("GMR_SCREEN_STATS") ? (unsigned long long*)(ws + L.misc + 64) : nullptr;
    a.debug = getenv("GMR_TC_DEBUG") ? atoi(getenv("GMR_TC_DEBUG")) : 0;
    a.row_done = ws + L.row_done;
    a.live_rows = (int32_t*)(ws + L.live_rows); a.n_live = (int32_t*)(misc + 34);
    GMR_CHECK_CUDA(cudaMemsetAsync(a.row_done, 0, (size_t)L.b_pad, st));
    GMR_CHECK_CUDA(cudaMemsetAsync(a.n_live, 0xFF, 4, st));   // -1 unless the head runs
    // exact head over the n_hot highest-norm items (sorted norms are still unscaled here); see the file header
    {
        const int nkey = K <= 64 ? 4 : 8;
        const size_t hsmem = head_smem_bytes(nkey, D);
        const char* env = getenv("GMR_SCREEN_HEAD");
        const bool head = bias == nullptr && K <= 128 && a.debug == 0 && hsmem <= 200 * 1024 && !(env && atoi(env) == 0);
        if (head) {
            HeadArgs h;
            h.nb_sorted = nb; h.perm_sorted = perm; h.hot_pos = (uint16_t*)(ws + L.hot_pos);
            h.hot_id = (int32_t*)(ws + L.hot_id); h.bound = (float*)(misc + 32); h.on = (int32_t*)(misc + 33);
            h.n_live = a.n_live;
            h.n_hot = 32 * nkey;
            h.margin = 1.00002f + 2.5e-7f * (float)D;
            h.hot_bits = (uint32_t*)(ws + L.hot_bits); h.row_bits = (uint32_t*)(ws + L.row_bits);
            GMR_CHECK_CUDA(cudaMemsetAsync(h.hot_pos, 0xFF, (size_t)I * 2, st));
            GMR_CHECK_CUDA(cudaMemsetAsync(h.hot_bits, 0, (size_t)(((int64_t)I + 31) / 32 * 4), st));
            score_head_setup_kernel<<<1, 256, 0, st>>>(h, I, K);
            GMR_LAUNCH_CHECK();
            if (mask_rowptr != nullptr) {
                GMR_CHECK_CUDA(cudaMemsetAsync(h.row_bits, 0, (size_t)L.b_pad * nkey * 4, st));
                score_head_maskbits_kernel<<<8 * sm_count(), 256, 0, st>>>(mask_rowptr, mask_items, (int64_t)B, h);
                GMR_LAUNCH_CHECK();
            }
            const int per_sm = hsmem <= 100 * 1024 ? 2 : 1;
            const int64_t want = ((int64_t)B + 8 * (32 / nkey) - 1) / (8 * (32 / nkey));
            const int hgrid = (int)(want < (int64_t)per_sm * sm_count() ? want : (int64_t)per_sm * sm_count());
            if (nkey == 4) {
                GMR_CHECK_CUDA(cudaFuncSetAttribute(score_head_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
                score_head_kernel<4><<<hgrid, 256, hsmem, st>>>(a, h);
            } else {
                GMR_CHECK_CUDA(cudaFuncSetAttribute(score_head_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
                score_head_kernel<8><<<hgrid, 256, hsmem, st>>>(a, h);
            }
            GMR_LAUNCH_CHECK();
        }
    }
    if (items_d64)
        prep_items_d64_kernel<<<(L.i_pad + 8 * wpb - 1) / (8 * wpb), wpb * 32, 0, st>>>(Ei, lde_i, I, L.i_pad, misc, perm, nb, b_h);
    else
        prep_items_kernel<<<(L.i_pad + wpb - 1) / wpb, wpb * 32, 0, st>>>(Ei, lde_i, I, L.i_pad, D, misc, perm, nb, b_h);
    GMR_LAUNCH_CHECK();
    if (D == 64 && lde_u % 4 == 0 && (uintptr_t)Eu % 16 == 0)
        prep_users_d64_kernel<<<(L.b_pad + 8 * wpb - 1) / (8 * wpb), wpb * 32, 0, st>>>(
            Eu, lde_u, users, B, L.b_pad, misc, nb, bias ? misc + 3 : nullptr, a.row_done, a_h, row_const);
    else
        prep_users_kernel<<<(L.b_pad + wpb - 1) / wpb, wpb * 32, 0, st>>>(Eu, lde_u, users, B, L.b_pad, D, misc, nb,
                                                                         bias ? misc + 3 : nullptr, a.row_done, a_h, row_const);
    GMR_LAUNCH_CHECK();
    CUtensorMap map_a, map_b;
    if (!scr_make_map(&map_a, a_h, L.b_pad, D) || !scr_make_map(&map_b, b_h, L.i_pad, D)) {
        set_error("score_topk_screen: cuTensorMapEncodeTiled unavailable or failed");
        return GMR_ERR_CUDA;
    }
    const int n_groups = L.b_pad / kRowsPerCta;
    GMR_CHECK_CUDA(cudaMemsetAsync(a.group_need, 0, (size_t)n_groups * 4, st));
    const int grid = scr_grid(B);
    const int npl = scr_npl(K);
    const int row_blocks = (B + 7) / 8;
    // finalisation: persistent CTAs (2 per SM) that stage the n_hot highest-norm item rows in shared memory
    int n_hot = 0;
    if (a.vec4) {
        for (n_hot = 256; n_hot > 0; n_hot >>= 1)
            if ((size_t)n_hot * (D + 4) * 4 + 8 * (size_t)D * 4 <= 100 * 1024) break;
        if (n_hot > L.i_pad) n_hot = L.i_pad;
    }
    const size_t fin_smem = (size_t)n_hot * (D + 4) * 4 + 8 * (size_t)D * 4;
    int fin_ctas_per_sm = (int)((220 * 1024) / (fin_smem > 0 ? fin_smem : 1));
    fin_ctas_per_sm = fin_ctas_per_sm < 1 ? 1 : (fin_ctas_per_sm > 3 ? 3 : fin_ctas_per_sm);
    const int fin_grid = row_blocks < fin_ctas_per_sm * sm_count() ? row_blocks : fin_ctas_per_sm * sm_count();
    const int chk_grid = row_blocks < 32 * sm_count() ? row_blocks : 32 * sm_count();   // strides over the live rows
    // phase 0 sweep -> checkpoint (mask + first L, per row at full occupancy) -> phase 1 sweep for the groups that
    // still need tiles -> finalisation (exact re-score + sort, per row at full occupancy)
#define GMR_SCR_LAUNCH(NPL, BIAS)                                                                                  \
    do {                                                                                                           \
        GMR_CHECK_CUDA(cudaFuncSetAttribute(score_screen_sweep_kernel<NPL, BIAS>,                                  \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        score_screen_sweep_kernel<NPL, BIAS><<<grid, kScrThreads, smem, st>>>(map_a, map_b, a, 0);                 \
        GMR_LAUNCH_CHECK();                                                                                        \
        if (a.debug == 0 || a.debug == 3) {                                                                        \
            score_screen_checkpoint_kernel<NPL><<<chk_grid, 256, 0, st>>>(a);                                      \
            GMR_LAUNCH_CHECK();                                                                                    \
        }                                                                                                          \
        score_screen_sweep_kernel<NPL, BIAS><<<grid, kScrThreads, smem, st>>>(map_a, map_b, a, 1);                 \
        GMR_LAUNCH_CHECK();                                                                                        \
        if (a.debug == 0 || a.debug == 3) {                                                                        \
            GMR_CHECK_CUDA(cudaFuncSetAttribute(score_screen_finalize_kernel<NPL>,                                 \
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));      \
            score_screen_finalize_kernel<NPL><<<fin_grid, 256, fin_smem, st>>>(a, n_hot);                          \
            GMR_LAUNCH_CHECK();                                                                                    \
        }                                                                                                          \
    } while (0)
    if (npl == 8) {
        if (bias) GMR_SCR_LAUNCH(8, true); else GMR_SCR_LAUNCH(8, false);
    } else if (npl == 16) {
        if (bias) GMR_SCR_LAUNCH(16, true); else GMR_SCR_LAUNCH(16, false);
    } else {
        if (bias) GMR_SCR_LAUNCH(32, true); else GMR_SCR_LAUNCH(32, false);
    }
#undef GMR_SCR_LAUNCH
    // queued rows: exact fp32 kernel, row count read on the device (no host synchronisation)
    score_simt_set_dynamic_rows((const int32_t*)(misc + 1));
    const int fb_grid = 2 * sm_count() < (B + 127) / 128 ? 2 * sm_count() : (B + 127) / 128;
    const int rc = score_topk_simt_launch(Eu, lde_u, users, fallback, B, Ei, lde_i, bias, I, D, mask_rowptr, mask_items, K,
                                          out_ids, out_scores, ws + L.simt, st, fb_grid);
    score_simt_set_dynamic_rows(nullptr);
    return rc;
}

}  // namespace gmr
