// topk_select.cuh -- per-row running top-K used by the fused scoring kernels (K2).
//
// Every candidate is one 64-bit key: (order-preserving fp32 score) << 32 | (0xFFFFFFFF - item id),
// so "larger key" == "higher score, then lower item id" -- the total order of the oracle
// (oracle/gmr_oracle.c: better()).  A row owns CAP = 32 * NPL key slots in the workspace:
// slots [0, KP) hold the current best KP = CAP / 4 keys (sorted, after a compaction), the rest is
// an append-only pending region.  A score is appended only if it beats the row's threshold (the
// key at position `kth` after the last compaction), so once the threshold is warm almost nothing
// is appended.  A compaction is a warp-level bitonic sort of the CAP keys held NPL per lane.
#pragma once

#include <stdint.h>

namespace gmr {

__device__ __forceinline__ uint32_t f32_to_ordered(float f)
{
    const uint32_t u = __float_as_uint(f + 0.0f);  // -0.0f -> +0.0f so that equal floats give equal keys
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(uint32_t o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ uint64_t make_key(float score, int32_t id)
{
    return ((uint64_t)f32_to_ordered(score) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)id);
}
__device__ __forceinline__ int32_t key_id(uint64_t k) { return (int32_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFu)); }
__device__ __forceinline__ float key_score(uint64_t k) { return ordered_to_f32((uint32_t)(k >> 32)); }

// Sort 32 * NPL keys held as k[r] = element (r * 32 + lane), descending (element 0 = largest).
template <int NPL, typename T = uint64_t>
__device__ __forceinline__ void warp_bitonic_sort_desc(T (&k)[NPL], int lane)
{
    constexpr int N = 32 * NPL;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride < 32) {
#pragma unroll
                for (int r = 0; r < NPL; ++r) {
                    const int i = r * 32 + lane;
                    const T other = __shfl_xor_sync(0xffffffffu, k[r], stride);
                    const bool desc = (i & size) == 0;
                    const bool lower = (lane & stride) == 0;
                    const bool take_max = (lower == desc);
                    const T mx = k[r] > other ? k[r] : other;
                    const T mn = k[r] > other ? other : k[r];
                    k[r] = take_max ? mx : mn;
                }
            } else {
                constexpr int dummy = 0;
                (void)dummy;
                const int rs = stride >> 5;
#pragma unroll
                for (int r = 0; r < NPL; ++r) {
                    if ((r & rs) == 0) {
                        const int r2 = r | rs;
                        const int i = r * 32 + lane;
                        const bool desc = (i & size) == 0;
                        const T a = k[r], b = k[r2];
                        const T mx = a > b ? a : b;
                        const T mn = a > b ? b : a;
                        k[r] = desc ? mx : mn;
                        k[r2] = desc ? mn : mx;
                    }
                }
            }
        }
    }
}

// Per-row selection state kept in shared memory by the calling CTA.
struct RowState {
    int cnt;           // valid keys in the row's slots (may run past CAP while appends are failing)
    float thr_score;   // score part of the threshold (fast reject)
    uint64_t thr_key;  // full threshold key: only keys > thr_key can still enter the top-K
};

// Binary search of `item` in the ascending list items[lo, hi).
__device__ __forceinline__ bool sorted_contains(const int32_t* __restrict__ items, int64_t lo, int64_t hi, int32_t item)
{
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int32_t v = items[mid];
        if (v == item) return true;
        if (v < item)
            lo = mid + 1;
        else
            hi = mid;
    }
    return false;
}

// Warp-cooperative compaction of one row: keep the best KP keys (sorted) in slots [0, KP),
// refresh the threshold from position `kth` (0-based).  All 32 lanes must call.
// When a mask list is given (mask_lo < mask_hi), keys were appended WITHOUT looking at the
// train-history mask; it is applied here, 32 * NPL binary searches in flight at once instead of one
// dependent global-load chain per append: a masked item keeps its id and scores exactly -1e10f
// (GenMMRec/src/common/trainer.py:384).
template <int NPL>
__device__ __noinline__ void compact_row(uint64_t* __restrict__ slots, RowState* st, int kth, int lane,
                                         const int32_t* __restrict__ mask_items = nullptr, int64_t mask_lo = 0,
                                         int64_t mask_hi = 0)
{
    constexpr int CAP = 32 * NPL;
    constexpr int KP = CAP / 4;
    static_assert(NPL % 4 == 0, "KP must be a multiple of 32 keys");
    const int cnt = min(st->cnt, CAP);
    uint64_t k[NPL];
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
        const int i = r * 32 + lane;
        k[r] = (i < cnt) ? slots[i] : 0ull;
    }
    if (mask_lo < mask_hi) {
        const uint32_t masked_hi = f32_to_ordered(-1e10f);
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            if (k[r] != 0ull && (uint32_t)(k[r] >> 32) != masked_hi &&
                sorted_contains(mask_items, mask_lo, mask_hi, key_id(k[r])))
                k[r] = ((uint64_t)masked_hi << 32) | (k[r] & 0xFFFFFFFFull);
        }
    }
    warp_bitonic_sort_desc<NPL>(k, lane);
#pragma unroll
    for (int r = 0; r < NPL / 4; ++r) slots[r * 32 + lane] = k[r];
    // threshold = key at sorted position kth (kth < KP)
    uint64_t t = 0ull;
#pragma unroll
    for (int r = 0; r < NPL / 4; ++r) {
        const uint64_t cand = __shfl_sync(0xffffffffu, k[r], kth & 31);
        if (r == (kth >> 5)) t = cand;
    }
    __syncwarp();
    if (lane == 0) {
        st->cnt = min(cnt, KP);
        const bool full = cnt > kth;
        st->thr_key = full ? t : 0ull;
        st->thr_score = full ? key_score(t) : -INFINITY;
    }
    __syncwarp();
}

// Cheap, conservative pruning of one row (warp-cooperative, O(CAP) work instead of a sort):
// if every lane holds at least two keys >= x, then at least 64 keys are >= x, so
// x = min over lanes of the lane's 2nd-largest key is a valid LOWER bound of the KP-th largest key for
// KP <= 64 * (NPL / 8)... in general the m-th largest per lane with 32 * m >= KP.  Keys below x can
// never reach the top KP and are dropped; survivors are compacted in place (unsorted) and x becomes
// the new row threshold.  Returns false when pruning could not free a quarter of the slots (massive
// ties): the caller then runs the exact compaction.
template <int NPL>
__device__ __noinline__ bool prune_row(uint64_t* __restrict__ slots, RowState* st, int lane)
{
    constexpr int CAP = 32 * NPL;
    constexpr int KP = CAP / 4;
    constexpr int M = KP / 32;  // per-lane order statistic: 32 * M == KP
    const int cnt = min(st->cnt, CAP);
    uint64_t k[NPL];
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
        const int i = r * 32 + lane;
        k[r] = (i < cnt) ? slots[i] : 0ull;
    }
    // lane-local M largest keys (M = 2 for NPL = 8): insertion into a tiny sorted list
    uint64_t top[M];
#pragma unroll
    for (int m = 0; m < M; ++m) top[m] = 0ull;
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
        uint64_t x = k[r];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint64_t hi = top[m] > x ? top[m] : x;
            x = top[m] > x ? x : top[m];
            top[m] = hi;
        }
    }
    uint64_t x = top[M - 1];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint64_t y = __shfl_xor_sync(0xffffffffu, x, o);
        x = y < x ? y : x;
    }
    // keep keys >= x, compact in place
    int mine = 0;
#pragma unroll
    for (int r = 0; r < NPL; ++r) mine += (k[r] != 0ull && k[r] >= x) ? 1 : 0;
    int pre = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += y;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    int pos = pre - mine;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NPL; ++r)
        if (k[r] != 0ull && k[r] >= x) slots[pos++] = k[r];
    __syncwarp();
    if (lane == 0) {
        st->cnt = total;
        if (x != 0ull && x > st->thr_key) {
            st->thr_key = x;
            st->thr_score = key_score(x);
        }
    }
    __syncwarp();
    return total <= CAP - CAP / 4;
}

}  // namespace gmr
