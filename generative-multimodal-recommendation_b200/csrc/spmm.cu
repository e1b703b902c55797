// spmm.cu -- K1: CSR SpMM  Y = alpha * A * X + beta * Y  (fp32, sm_100a)
//
// Replaces torch.sparse.mm / torch.spmm(A_coo, X) at the reference call sites listed in
// include/gmr.h (GenMMRec/src/models/diffmm.py:136-152,284-285 and siblings).
//
// Schedule
//   * the plan cuts every row into "virtual rows" of at most `chunk` nonzeros; one WARP owns one
//     virtual row, so no warp ever walks a power-law row alone (item rows of the 1M-user shape
//     reach 10^5..10^6 nonzeros).  Virtual rows are contiguous in nnz space, so they are described
//     by one CSR-like pointer array (vptr) plus a destination per virtual row.
//   * whole rows are written straight to Y; chunks of split rows write fp32 partial sums to the
//     workspace and a second kernel adds them up in slot order (fixed order => deterministic).
//   * inside a warp the (col, val) pairs of up to 32 nonzeros are loaded coalesced, staged in
//     shared memory, and consumed LANES lanes per nonzero: D = 64 uses 16 lanes x 128-bit per
//     gathered embedding row, i.e. two nonzeros per warp step; loads are issued UNROLL steps ahead
//     of the FMAs to keep >= 8 independent 128-bit gathers in flight per warp.
//   * the CSR stream is read with L1::no_allocate + L2::evict_first, gathered rows with
//     L2::evict_last: the only operand with reuse keeps the 126 MB L2.
//
// HBM-bound: algorithmic bytes = nnz*8 + (n_rows+1)*4 + n_cols*D*4 + n_rows*D*4.
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "common.cuh"

struct gmr_spmm_plan {
    int64_t n_rows = 0, n_cols = 0, nnz = 0;
    int32_t chunk = 0;
    int64_t n_vrows = 0;       // virtual rows (chunks)
    int64_t n_split_rows = 0;  // rows cut into more than one chunk
    int64_t n_slots = 0;       // partial-sum slots = chunks of split rows
    int32_t max_slots_per_row = 0;
    // device arrays (null when the plan is the identity: no row longer than `chunk`)
    int32_t* d_vptr = nullptr;         // [n_vrows + 1]
    int32_t* d_vrow = nullptr;         // [n_vrows]  row id, or -1 - slot for a chunk of a split row
    int32_t* d_split_row = nullptr;    // [n_split_rows]
    int32_t* d_split_first = nullptr;  // [n_split_rows + 1] first slot of each split row
};

namespace gmr {

constexpr int kWarps = 8;  // warps per CTA (256 threads)

template <typename Vec>
struct VecOps;
template <>
struct VecOps<float4> {
    static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ float4 gather(const float4* p, uint64_t pol) { return ld_gather_f4(p, pol); }
    static __device__ __forceinline__ void fma(float4& a, float w, const float4& x)
    {
        a.x = fmaf(w, x.x, a.x);
        a.y = fmaf(w, x.y, a.y);
        a.z = fmaf(w, x.z, a.z);
        a.w = fmaf(w, x.w, a.w);
    }
    static __device__ __forceinline__ float4 xor_add(float4 a, int m)
    {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, m);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, m);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, m);
        a.w += __shfl_xor_sync(0xffffffffu, a.w, m);
        return a;
    }
    static __device__ __forceinline__ float4 axpby(float alpha, const float4& a, float beta, const float4& y)
    {
        return make_float4(fmaf(alpha, a.x, beta * y.x), fmaf(alpha, a.y, beta * y.y), fmaf(alpha, a.z, beta * y.z),
                           fmaf(alpha, a.w, beta * y.w));
    }
    static __device__ __forceinline__ float4 scale(float alpha, const float4& a)
    {
        return make_float4(alpha * a.x, alpha * a.y, alpha * a.z, alpha * a.w);
    }
    static __device__ __forceinline__ float4 add(const float4& a, const float4& b)
    {
        return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
};
template <>
struct VecOps<float> {
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float gather(const float* p, uint64_t pol) { return ld_gather_f1(p, pol); }
    static __device__ __forceinline__ void fma(float& a, float w, const float& x) { a = fmaf(w, x, a); }
    static __device__ __forceinline__ float xor_add(float a, int m) { return a + __shfl_xor_sync(0xffffffffu, a, m); }
    static __device__ __forceinline__ float axpby(float alpha, const float& a, float beta, const float& y)
    {
        return fmaf(alpha, a, beta * y);
    }
    static __device__ __forceinline__ float scale(float alpha, const float& a) { return alpha * a; }
    static __device__ __forceinline__ float add(const float& a, const float& b) { return a + b; }
};

struct PushArgs {
    float* const* y_peers;
    int32_t n_peers;
    int64_t row_offset;
};

// One warp per virtual row.  Vec = float4 (D % 4 == 0, aligned) or float.  DV = row length in Vec.
// OFF32: every gathered address fits a 32-bit offset in Vec units (n_cols * ldx / VEC < 2^32), so the
// staged entry is the row offset itself and the gather address is one IMAD.WIDE away.
// EXACT: DV == LANES * ITER, no column predicate.
template <typename Vec, int LANES, int ITER, bool IDENT, bool PUSH, bool OFF32, bool EXACT>
__global__ void __launch_bounds__(kWarps * 32, (ITER == 1) ? 4 : 2)
    spmm_vrow_kernel(const int32_t* __restrict__ vptr, const int32_t* __restrict__ vrow,
                     const int32_t* __restrict__ col, const float* __restrict__ val, const float* __restrict__ X,
                     int64_t ldx, float* __restrict__ Y, int64_t ldy, float* __restrict__ partial, int32_t DV,
                     int64_t n_vrows, float alpha, float beta, PushArgs push)
{
    using Ops = VecOps<Vec>;
    constexpr int VEC = sizeof(Vec) / 4;
    constexpr int NPS = 32 / LANES;              // nonzeros consumed per warp step
    constexpr int UNROLL = (ITER >= 2) ? 2 : 4;  // steps whose gathers are issued back to back
    constexpr int GROUP = NPS * UNROLL;          // nonzeros per unrolled batch
    constexpr int STAGE = 256;  // (col, val) pairs staged per pass: a whole chunk at the default size
    // double-buffered staging of the CSR stream: the NEXT virtual row's (col, val) pairs are copied global -> shared
    // with cp.async while the current row's gathers are in flight, and the row after that has its pointers loaded,
    // so a row's critical path is its gathers only (it used to be pointer load -> CSR load -> gathers -> store:
    // two of the four memory latencies of a 50-nonzero row)
    __shared__ int scol[kWarps][2][STAGE];
    __shared__ float sval[kWarps][2][STAGE];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane % LANES, grp = lane / LANES;
    const int vcol0 = blockIdx.y * (LANES * ITER) + sub;  // this lane's first Vec column
    const uint32_t ldv = (uint32_t)(ldx / VEC);           // row pitch in Vec units (OFF32 only)
    const Vec* __restrict__ Xl = reinterpret_cast<const Vec*>(X) + vcol0;
    bool colok[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) colok[it] = EXACT || (vcol0 + it * LANES < DV);

    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    auto issue_stage = [&](int buf, int base, int n) {  // n <= STAGE entries, one cp.async group per call
        for (int k = lane; k < n; k += 32) {
            asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(&scol[warp][buf][k])),
                         "l"(col + base + k), "l"(pol_stream)
                         : "memory");
            asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(&sval[warp][buf][k])),
                         "l"(val + base + k), "l"(pol_stream)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // Persistent warps: every warp walks virtual rows v, v + W, v + 2W, ... (W = warps in the grid).  Row lengths are
    // power-law distributed; with one virtual row per warp a CTA stays resident until its longest row is done and the
    // SM runs at a third of its warp slots (ncu: 32 % warps active).  Striding keeps every slot busy until the tail.
    const int64_t n_warps_grid = (int64_t)gridDim.x * kWarps;
    int64_t v = (int64_t)blockIdx.x * kWarps + warp;
    if (v >= n_vrows) return;  // whole warp leaves together; only __syncwarp below
    int begin = vptr[v], end = vptr[v + 1];
    int dst = IDENT ? (int)v : vrow[v];
    issue_stage(0, begin, min(STAGE, end - begin));
    int64_t vn = v + n_warps_grid;
    int begin1 = 0, end1 = 0, dst1 = 0;
    if (vn < n_vrows) {
        begin1 = vptr[vn];
        end1 = vptr[vn + 1];
        dst1 = IDENT ? (int)vn : vrow[vn];
    }
    int buf = 0;
    while (true) {
    const bool has_next = vn < n_vrows;
    if (has_next)
        issue_stage(buf ^ 1, begin1, min(STAGE, end1 - begin1));
    else
        asm volatile("cp.async.commit_group;" ::: "memory");  // keeps the group count uniform
    const int64_t vnn = vn + n_warps_grid;
    int begin2 = 0, end2 = 0, dst2 = 0;
    if (vnn < n_vrows) {
        begin2 = vptr[vnn];
        end2 = vptr[vnn + 1];
        dst2 = IDENT ? (int)vnn : vrow[vnn];
    }
    Vec acc[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) acc[it] = Ops::zero();

    auto row_ptr = [&](int c) -> const Vec* {
        if (OFF32) return Xl + (uint32_t)c * ldv;
        return reinterpret_cast<const Vec*>(X + (int64_t)c * ldx) + vcol0;
    };
    const int* my_col = scol[warp][buf];
    const float* my_val = sval[warp][buf];
    // one batch = GROUP nonzeros: UNROLL gathers per lane
    auto load_group = [&](int g, float (&cw)[UNROLL], Vec (&xv)[UNROLL][ITER]) {
        int cc[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            cc[q] = my_col[g * GROUP + q * NPS + grp];
            cw[q] = my_val[g * GROUP + q * NPS + grp];
        }
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            const Vec* xr = row_ptr(cc[q]);
#pragma unroll
            for (int it = 0; it < ITER; ++it)
                xv[q][it] = colok[it] ? Ops::gather(xr + it * LANES, pol_keep) : Ops::zero();
        }
    };
    auto fma_group = [&](const float (&cw)[UNROLL], const Vec (&xv)[UNROLL][ITER]) {
#pragma unroll
        for (int q = 0; q < UNROLL; ++q)
#pragma unroll
            for (int it = 0; it < ITER; ++it) Ops::fma(acc[it], cw[q], xv[q][it]);
    };

    asm volatile("cp.async.wait_group 1;" ::: "memory");  // this row's first STAGE entries have landed (own copies)
    __syncwarp();                                         // ... and every other lane's
    for (int base = begin; base < end; base += STAGE) {
        const int n = min(STAGE, end - base);
        if (base != begin) {  // rows longer than one stage (custom chunk sizes only): plain synchronous staging
            __syncwarp();
            for (int k = lane; k < n; k += 32) {
                scol[warp][buf][k] = ld_stream_s32(col + base + k, pol_stream);
                sval[warp][buf][k] = ld_stream_f32(val + base + k, pol_stream);
            }
            __syncwarp();
        }
        // software pipeline over full batches: the gathers of batch g+1 are issued before the FMAs of
        // batch g, so every warp keeps UNROLL..2*UNROLL independent 128-bit gathers in flight
        const int ng = n / GROUP;
        if (ng > 0) {
            float cwA[UNROLL], cwB[UNROLL];
            Vec xA[UNROLL][ITER], xB[UNROLL][ITER];
            load_group(0, cwA, xA);
            int g = 0;
            while (true) {
                if (g + 1 < ng) load_group(g + 1, cwB, xB);
                fma_group(cwA, xA);
                if (++g >= ng) break;
                if (g + 1 < ng) load_group(g + 1, cwA, xA);
                fma_group(cwB, xB);
                if (++g >= ng) break;
            }
        }
        for (int e = ng * GROUP + grp; e < n; e += NPS) {  // ragged tail of the row
            const Vec* xr = row_ptr(my_col[e]);
            const float w = my_val[e];
#pragma unroll
            for (int it = 0; it < ITER; ++it)
                if (colok[it]) Ops::fma(acc[it], w, Ops::gather(xr + it * LANES, pol_keep));
        }
    }
    __syncwarp();  // every lane is done with this stage buffer before the row after next is copied into it

    if (NPS == 2) {
#pragma unroll
        for (int it = 0; it < ITER; ++it) acc[it] = Ops::xor_add(acc[it], 16);
    }
    if (grp == 0) {
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int vc = vcol0 + it * LANES;
        if (!colok[it]) continue;
        if (!IDENT && dst < 0) {
            // chunk of a split row: raw partial sum, reduced later in slot order
            Vec* p = reinterpret_cast<Vec*>(partial + (int64_t)(-1 - dst) * ((int64_t)DV * VEC)) + vc;
            *p = acc[it];
        } else if (PUSH) {
            const Vec out = Ops::scale(alpha, acc[it]);
            for (int p = 0; p < push.n_peers; ++p) {
                Vec* y = reinterpret_cast<Vec*>(push.y_peers[p] + (push.row_offset + dst) * ldy) + vc;
                *y = out;
            }
        } else {
            Vec* y = reinterpret_cast<Vec*>(Y + (int64_t)dst * ldy) + vc;
            *y = (beta == 0.f) ? Ops::scale(alpha, acc[it]) : Ops::axpby(alpha, acc[it], beta, *y);
        }
    }
    }
    if (!has_next) break;
    v = vn; begin = begin1; end = end1; dst = dst1;
    vn = vnn; begin1 = begin2; end1 = end2; dst1 = dst2;
    buf ^= 1;
    }  // persistent loop over virtual rows
}

// One CTA per split row: warp w sums the slots of its fixed sub-range in order, the eight warp
// sums are then added in warp order.  Fixed order => deterministic.
template <typename Vec, bool PUSH>
__global__ void __launch_bounds__(kWarps * 32)
    spmm_reduce_kernel(const int32_t* __restrict__ split_row, const int32_t* __restrict__ split_first,
                       const float* __restrict__ partial, float* __restrict__ Y, int64_t ldy, int32_t DV,
                       float alpha, float beta, PushArgs push)
{
    using Ops = VecOps<Vec>;
    constexpr int VEC = sizeof(Vec) / 4;
    extern __shared__ float4 red_smem[];  // [kWarps][DV] Vec
    Vec* red = reinterpret_cast<Vec*>(red_smem);

    const int s = blockIdx.x;
    const int row = split_row[s];
    const int first = split_first[s], n = split_first[s + 1] - first;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (n + kWarps - 1) / kWarps;
    const int lo = min(n, warp * per), hi = min(n, lo + per);
    const Vec* base = reinterpret_cast<const Vec*>(partial + (int64_t)first * ((int64_t)DV * VEC));

    for (int vc = lane; vc < DV; vc += 32) {
        Vec a = Ops::zero();
        for (int k = lo; k < hi; ++k) a = Ops::add(a, base[(int64_t)k * DV + vc]);
        red[warp * DV + vc] = a;
    }
    __syncthreads();
    for (int vc = threadIdx.x; vc < DV; vc += kWarps * 32) {
        Vec a = red[vc];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) a = Ops::add(a, red[w * DV + vc]);
        if (PUSH) {
            const Vec out = Ops::scale(alpha, a);
            for (int p = 0; p < push.n_peers; ++p)
                reinterpret_cast<Vec*>(push.y_peers[p] + (push.row_offset + row) * ldy)[vc] = out;
        } else {
            Vec* y = reinterpret_cast<Vec*>(Y + (int64_t)row * ldy) + vc;
            *y = (beta == 0.f) ? Ops::scale(alpha, a) : Ops::axpby(alpha, a, beta, *y);
        }
    }
}

// Short-row variant (D = 64 floats, float4 lanes): one HALF-warp per virtual row, so a warp retires two rows per trip.
// kNN modality graphs (10 neighbours per item) and the merged modal-mix graph (~5 nonzeros per row) spend the general
// kernel's time on per-row overhead, not on gathers: a full warp per row stages 256 slots for 5 entries and leaves half
// its gather slots empty.  Here the (col, val) pairs live in registers (one per lane, broadcast by shuffle), the next
// row's pairs and the row after that's pointers are loaded while this row's gathers are in flight, and nothing touches
// shared memory.  Summation order is the general kernel's (even nonzeros into one fmaf chain, odd ones into another,
// chains added at the end), so the two kernels return bit-identical rows and the choice is a pure scheduling matter.
template <bool IDENT>
__global__ void __launch_bounds__(kWarps * 32, 3)
    spmm_short_d64_kernel(const int32_t* __restrict__ vptr, const int32_t* __restrict__ vrow,
                          const int32_t* __restrict__ col, const float* __restrict__ val, const float* __restrict__ X,
                          int64_t ldx, float* __restrict__ Y, int64_t ldy, float* __restrict__ partial, int64_t n_vrows,
                          float alpha, float beta)
{
    using Ops = VecOps<float4>;
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31, sub = lane & 15, half = lane >> 4;
    const int64_t n_hw = (int64_t)gridDim.x * (kWarps * 2);  // half-warps in the grid (even: the two halves stay paired)
    int64_t v = ((int64_t)blockIdx.x * (kWarps * 32) + threadIdx.x) >> 4;
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const float4* __restrict__ Xl = reinterpret_cast<const float4*>(X) + sub;
    const int64_t ldv = ldx >> 2;

    auto load_ptr = [&](int64_t row, int& b, int& e, int& d) {
        if (row < n_vrows) {
            b = vptr[row];
            e = vptr[row + 1];
            d = IDENT ? (int)row : vrow[row];
        } else {
            b = e = d = 0;
        }
    };
    auto load_cv = [&](int b, int e, int& c, float& w) {
        c = 0;
        w = 0.f;
        if (b + sub < e) {
            c = ld_stream_s32(col + b + sub, pol_stream);
            w = ld_stream_f32(val + b + sub, pol_stream);
        }
    };
    int b, e, dst, b1, e1, dst1, c;
    float w;
    load_ptr(v, b, e, dst);
    load_cv(b, e, c, w);
    load_ptr(v + n_hw, b1, e1, dst1);
    while (v - half < n_vrows) {  // warp-uniform: half 1 of the last pair may idle on an empty row
        int c1, b2, e2, dst2;
        float w1;
        load_cv(b1, e1, c1, w1);
        load_ptr(v + 2 * n_hw, b2, e2, dst2);
        const bool live = v < n_vrows;
        const bool to_y = live && (IDENT || dst >= 0);
        float4* yp = reinterpret_cast<float4*>(Y + (int64_t)(to_y ? dst : 0) * ldy) + sub;
        float4 yold = Ops::zero();
        if (to_y && beta != 0.f) yold = *yp;
        const int n = e - b;
        const int nmax = max(n, __shfl_xor_sync(kFull, n, 16));
        float4 acc_e = Ops::zero(), acc_o = Ops::zero();
        for (int base = 0; base < nmax; base += 16) {
            if (base) load_cv(b + base, e, c, w);  // rows longer than 16 nonzeros: plain reload
            const int m = n - base;
            const int mm = min(16, nmax - base);
            for (int j = 0; j < mm; j += 4) {
                float4 xv[4];
                float ww[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cj = __shfl_sync(kFull, c, j + q, 16);
                    ww[q] = __shfl_sync(kFull, w, j + q, 16);
                    xv[q] = (j + q < m) ? Ops::gather(Xl + (int64_t)cj * ldv, pol_keep) : Ops::zero();
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (j + q < m) Ops::fma((q & 1) ? acc_o : acc_e, ww[q], xv[q]);
            }
        }
        const float4 acc = Ops::add(acc_e, acc_o);
        if (to_y)
            *yp = (beta == 0.f) ? Ops::scale(alpha, acc) : Ops::axpby(alpha, acc, beta, yold);
        else if (live)
            reinterpret_cast<float4*>(partial + (int64_t)(-1 - dst) * 64)[sub] = acc;
        v += n_hw;
        b = b1; e = e1; dst = dst1; c = c1; w = w1;
        b1 = b2; e1 = e2; dst1 = dst2;
    }
}

// GMR_SPMM_SHORT: unset = pick by mean virtual-row length, 0 = never, 1 = whenever the shape allows.  Read on every call
// (a getenv is noise next to a launch) so the tests can run both kernels in one process.
static int spmm_short_mode()
{
    const char* s = getenv("GMR_SPMM_SHORT");
    return (s && *s) ? atoi(s) : -1;
}

template <typename Vec, int LANES, int ITER, bool PUSH, bool OFF32, bool EXACT>
static int launch_vrow2(const gmr_spmm_plan* plan, const int32_t* rowptr, const int32_t* col, const float* val,
                        const float* X, int64_t ldx, float* Y, int64_t ldy, float* partial, int32_t DV, float alpha,
                        float beta, PushArgs push, cudaStream_t st)
{
    const bool ident = plan->d_vptr == nullptr;
    const int64_t nv = plan->n_vrows;
    if (nv == 0) return GMR_OK;
    // persistent grid: enough CTAs to fill every SM's warp slots (8 CTAs x 8 warps), never more than the work
    const int64_t want = (nv + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)sm_count() * 8;
    dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)((DV + LANES * ITER - 1) / (LANES * ITER)));
    dim3 block(kWarps * 32);
    if (ident)
        spmm_vrow_kernel<Vec, LANES, ITER, true, PUSH, OFF32, EXACT><<<grid, block, 0, st>>>(
            rowptr, nullptr, col, val, X, ldx, Y, ldy, partial, DV, nv, alpha, beta, push);
    else
        spmm_vrow_kernel<Vec, LANES, ITER, false, PUSH, OFF32, EXACT><<<grid, block, 0, st>>>(
            plan->d_vptr, plan->d_vrow, col, val, X, ldx, Y, ldy, partial, DV, nv, alpha, beta, push);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

template <typename Vec, int LANES, int ITER, bool PUSH>
static int launch_vrow(const gmr_spmm_plan* plan, const int32_t* rowptr, const int32_t* col, const float* val,
                       const float* X, int64_t ldx, float* Y, int64_t ldy, float* partial, int32_t DV, float alpha,
                       float beta, PushArgs push, cudaStream_t st)
{
    constexpr int VEC = sizeof(Vec) / 4;
    // 32-bit row offsets (in Vec units) whenever the whole X operand is addressable that way
    const bool off32 = (ldx % VEC == 0) && ((plan->n_cols + 1) * (ldx / VEC) < (int64_t)0xffffffffll);
    const bool exact = (DV == LANES * ITER);
#define GMR_GO(O, E) \
    return launch_vrow2<Vec, LANES, ITER, PUSH, O, E>(plan, rowptr, col, val, X, ldx, Y, ldy, partial, DV, alpha, beta, push, st)
    if (off32 && exact) GMR_GO(true, true);
    if (off32) GMR_GO(true, false);
    GMR_GO(false, false);
#undef GMR_GO
}

template <typename Vec, bool PUSH>
static int dispatch(const gmr_spmm_plan* plan, const int32_t* rowptr, const int32_t* col, const float* val,
                    const float* X, int64_t ldx, float* Y, int64_t ldy, float* partial, int32_t D, float alpha,
                    float beta, PushArgs push, cudaStream_t st)
{
    constexpr int VEC = sizeof(Vec) / 4;
    const int32_t DV = D / VEC;
    int rc;
    const int short_mode = spmm_short_mode();
    if (VEC == 4 && DV == 16 && !PUSH && plan->n_vrows > 0 && short_mode != 0 &&
        (short_mode > 0 || plan->nnz <= 12 * plan->n_vrows)) {
        const int64_t want = (plan->n_vrows + 2 * kWarps - 1) / (2 * kWarps);
        const int64_t cap = (int64_t)sm_count() * 3;
        const unsigned grid = (unsigned)(want < cap ? want : cap);
        if (plan->d_vptr == nullptr)
            spmm_short_d64_kernel<true><<<grid, kWarps * 32, 0, st>>>(rowptr, nullptr, col, val, X, ldx, Y, ldy, partial,
                                                                      plan->n_vrows, alpha, beta);
        else
            spmm_short_d64_kernel<false><<<grid, kWarps * 32, 0, st>>>(plan->d_vptr, plan->d_vrow, col, val, X, ldx, Y,
                                                                       ldy, partial, plan->n_vrows, alpha, beta);
        GMR_LAUNCH_CHECK();
        rc = GMR_OK;
    } else if (DV <= 16)
        rc = launch_vrow<Vec, 16, 1, PUSH>(plan, rowptr, col, val, X, ldx, Y, ldy, partial, DV, alpha, beta, push, st);
    else if (DV <= 32)
        rc = launch_vrow<Vec, 32, 1, PUSH>(plan, rowptr, col, val, X, ldx, Y, ldy, partial, DV, alpha, beta, push, st);
    else if (DV <= 64)
        rc = launch_vrow<Vec, 32, 2, PUSH>(plan, rowptr, col, val, X, ldx, Y, ldy, partial, DV, alpha, beta, push, st);
    else
        rc = launch_vrow<Vec, 32, 4, PUSH>(plan, rowptr, col, val, X, ldx, Y, ldy, partial, DV, alpha, beta, push, st);
    if (rc != GMR_OK) return rc;
    if (plan->n_split_rows > 0) {
        const size_t smem = (size_t)kWarps * DV * sizeof(Vec);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(spmm_reduce_kernel<Vec, PUSH>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) {
                set_error("row width D=%d too large for the split-row reduction (%zu B shared)", D, smem);
                return GMR_ERR_UNSUPPORTED;
            }
        }
        spmm_reduce_kernel<Vec, PUSH><<<(unsigned)plan->n_split_rows, kWarps * 32, smem, st>>>(
            plan->d_split_row, plan->d_split_first, partial, Y, ldy, DV, alpha, beta, push);
        GMR_LAUNCH_CHECK();
    }
    return GMR_OK;
}

static int spmm_common(const gmr_spmm_plan* plan, const int32_t* rowptr, const int32_t* col, const float* val,
                       const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t D, float alpha, float beta,
                       PushArgs push, bool is_push, void* workspace, int64_t workspace_bytes, void* stream)
{
    GMR_REQUIRE(plan != nullptr, "gmr_spmm: plan is null");
    GMR_REQUIRE(D >= 1, "gmr_spmm: D must be >= 1 (got %d)", D);
    GMR_REQUIRE(ldx >= D && ldy >= D, "gmr_spmm: leading dimensions (%lld, %lld) smaller than D=%d", (long long)ldx,
                (long long)ldy, D);
    if (plan->n_rows == 0) return GMR_OK;
    GMR_REQUIRE(rowptr && X && (Y || is_push), "gmr_spmm: null operand");
    GMR_REQUIRE(plan->nnz == 0 || (col && val), "gmr_spmm: null col/val with nnz > 0");
    const int64_t need = gmr_spmm_workspace_bytes(plan, D);
    if (need > 0 && (workspace == nullptr || workspace_bytes < need)) {
        set_error("gmr_spmm: workspace of %lld bytes required, %lld given", (long long)need, (long long)workspace_bytes);
        return GMR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = (float*)workspace;
    bool vec4 = (D % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && ((uintptr_t)X % 16 == 0);
    if (!is_push) vec4 = vec4 && ((uintptr_t)Y % 16 == 0);
    if (is_push) {
        if (vec4) return dispatch<float4, true>(plan, rowptr, col, val, X, ldx, Y, ldy, partial, D, alpha, beta, push, st);
        return dispatch<float, true>(plan, rowptr, col, val, X, ldx, Y, ldy, partial, D, alpha, beta, push, st);
    }
    if (vec4) return dispatch<float4, false>(plan, rowptr, col, val, X, ldx, Y, ldy, partial, D, alpha, beta, push, st);
    return dispatch<float, false>(plan, rowptr, col, val, X, ldx, Y, ldy, partial, D, alpha, beta, push, st);
}

// ---- row-wise glue ---------------------------------------------------------------------------
// out = a*x + b*y + c * z / max(||z||_2, eps); one warp per row.
__global__ void __launch_bounds__(256)
    rows_axpby_norm_kernel(const float* x, int64_t ldx, const float* y, int64_t ldy,   // no __restrict__: out may alias
                           const float* z, int64_t ldz, float* out, int64_t ldo,       // x, y or z (in-place accumulate)
                           int64_t n_rows, int32_t D, float a, float b, float c, float eps)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    float inv = 0.f;
    if (z != nullptr) {
        float ss = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float t = z[r * ldz + d];
            ss = fmaf(t, t, ss);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
        inv = c / fmaxf(sqrtf(ss), eps);
    }
    for (int d = lane; d < D; d += 32) {
        float o = a * x[r * ldx + d];
        if (y != nullptr) o = fmaf(b, y[r * ldy + d], o);
        if (z != nullptr) o = fmaf(inv, z[r * ldz + d], o);
        out[r * ldo + d] = o;
    }
}

// D = 64, 16-byte aligned rows: 16 lanes x float4 per row, two rows per warp, one 128-bit load per operand and lane
// (the generic kernel's 4-byte loads reach a third of the copy bandwidth on this 1.2 GB pass).
__global__ void __launch_bounds__(256)
    rows_axpby_norm_d64_kernel(const float* x, int64_t ldx, const float* y, int64_t ldy,   // no __restrict__: out may alias
                               const float* z, int64_t ldz, float* out, int64_t ldo,       // x, y or z
                               int64_t n_rows, float a, float b, float c, float eps)
{
    const int lane = threadIdx.x & 31, sub = lane & 15;
    const int64_t r = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + (lane >> 4);
    const bool ok = r < n_rows;
    float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
    float inv = 0.f;
    if (z != nullptr) {
        if (ok) zv = *reinterpret_cast<const float4*>(z + r * ldz + 4 * sub);
        float ss = fmaf(zv.x, zv.x, fmaf(zv.y, zv.y, fmaf(zv.z, zv.z, zv.w * zv.w)));
#pragma unroll
        for (int m = 8; m > 0; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);   // within the 16-lane half
        inv = c / fmaxf(sqrtf(ss), eps);
    }
    if (!ok) return;
    const float4 xv = (z == x && ldz == ldx) ? zv : *reinterpret_cast<const float4*>(x + r * ldx + 4 * sub);
    float4 o = make_float4(a * xv.x, a * xv.y, a * xv.z, a * xv.w);
    if (y != nullptr) {
        const float4 yv = *reinterpret_cast<const float4*>(y + r * ldy + 4 * sub);
        o.x = fmaf(b, yv.x, o.x); o.y = fmaf(b, yv.y, o.y); o.z = fmaf(b, yv.z, o.z); o.w = fmaf(b, yv.w, o.w);
    }
    if (z != nullptr) {
        o.x = fmaf(inv, zv.x, o.x); o.y = fmaf(inv, zv.y, o.y); o.z = fmaf(inv, zv.z, o.z); o.w = fmaf(inv, zv.w, o.w);
    }
    *reinterpret_cast<float4*>(out + r * ldo + 4 * sub) = o;
}

// out[:, 0:D] = w1 * n(act(x1)) + w2 * n(act(x2)),  out[:, D:2D] = out[:, 0:D] + y   (y optional)
// with act = leaky ReLU (slope) and n(v) = v / max(||v||_2, eps); one warp per row, values kept in
// registers between the norm pass and the write (D <= 256).
__global__ void __launch_bounds__(256)
    rows_normalize_mix_kernel(const float* __restrict__ x1, int64_t ld1, const float* __restrict__ x2, int64_t ld2,
                              const float* __restrict__ y, int64_t ldy, float* __restrict__ out, int64_t ldo,
                              int64_t n_rows, int32_t D, float w1, float w2, float slope, float eps)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    float a[8], b[8];
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d = lane + 32 * j;
        a[j] = b[j] = 0.f;
        if (d < D) {
            float v = x1[r * ld1 + d];
            v = v > 0.f ? v : v * slope;
            a[j] = v;
            sa = fmaf(v, v, sa);
            v = x2[r * ld2 + d];
            v = v > 0.f ? v : v * slope;
            b[j] = v;
            sb = fmaf(v, v, sb);
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, m);
        sb += __shfl_xor_sync(0xffffffffu, sb, m);
    }
    const float ia = w1 / fmaxf(sqrtf(sa), eps), ib = w2 / fmaxf(sqrtf(sb), eps);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d = lane + 32 * j;
        if (d < D) {
            const float z = fmaf(ia, a[j], ib * b[j]);
            out[r * ldo + d] = z;
            if (y != nullptr) out[r * ldo + D + d] = z + y[r * ldy + d];
        }
    }
}

// Copy a block of rows into every peer's replica (dense all-gather by peer stores over NVLink): 128-bit loads,
// one store per peer; grid-stride over 16-byte pieces.
__global__ void __launch_bounds__(256)
    rows_push_kernel(const float4* __restrict__ src, int64_t ld4, int64_t n_rows, int32_t d4, float* const* __restrict__ y_peers,
                     int32_t n_peers, int64_t row_offset, int64_t ldy4)
{
    const int64_t total = n_rows * d4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d4;
        const int c = (int)(i - r * d4);
        const float4 v = src[r * ld4 + c];
        for (int p = 0; p < n_peers; ++p) reinterpret_cast<float4*>(y_peers[p])[(row_offset + r) * ldy4 + c] = v;
    }
}

}  // namespace gmr

// ---- C ABI -------------------------------------------------------------------------------------

extern "C" int gmr_spmm_plan_create(gmr_spmm_plan_t** out, const int32_t* rowptr, int64_t n_rows, int64_t n_cols,
                                    int32_t chunk_nnz, void* stream)
{
    GMR_REQUIRE(out != nullptr, "gmr_spmm_plan_create: null output handle");
    *out = nullptr;
    GMR_REQUIRE(n_rows >= 0 && n_cols >= 0, "gmr_spmm_plan_create: negative shape");
    GMR_REQUIRE(n_rows < (int64_t)0x7fffffff && n_cols < (int64_t)0x7fffffff, "gmr_spmm_plan_create: shape exceeds int32");
    GMR_REQUIRE(n_rows == 0 || rowptr != nullptr, "gmr_spmm_plan_create: null rowptr");
    GMR_REQUIRE(chunk_nnz >= 0, "gmr_spmm_plan_create: negative chunk");
    const int32_t chunk = chunk_nnz == 0 ? 256 : std::max(32, (chunk_nnz + 31) / 32 * 32);
    cudaStream_t st = (cudaStream_t)stream;

    std::vector<int32_t> h(n_rows + 1, 0);
    if (n_rows > 0) {
        GMR_CHECK_CUDA(cudaMemcpyAsync(h.data(), rowptr, sizeof(int32_t) * (n_rows + 1), cudaMemcpyDeviceToHost, st));
        GMR_CHECK_CUDA(cudaStreamSynchronize(st));
    }
    GMR_REQUIRE(h[0] == 0, "gmr_spmm_plan_create: rowptr[0] must be 0 (got %d)", h[0]);
    for (int64_t r = 0; r < n_rows; ++r)
        GMR_REQUIRE(h[r + 1] >= h[r], "gmr_spmm_plan_create: rowptr not monotone at row %lld", (long long)r);

    auto* p = new gmr_spmm_plan();
    p->n_rows = n_rows;
    p->n_cols = n_cols;
    p->nnz = n_rows > 0 ? h[n_rows] : 0;
    p->chunk = chunk;
    int64_t n_split = 0, n_extra = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int32_t len = h[r + 1] - h[r];
        if (len > chunk) {
            ++n_split;
            n_extra += (len + chunk - 1) / chunk - 1;
        }
    }
    p->n_split_rows = n_split;
    p->n_vrows = n_rows + n_extra;
    if (n_split > 0) {
        std::vector<int32_t> vptr, vrow, srow, sfirst;
        vptr.reserve(p->n_vrows + 1);
        vrow.reserve(p->n_vrows);
        vptr.push_back(0);
        int32_t slot = 0;
        for (int64_t r = 0; r < n_rows; ++r) {
            const int32_t b = h[r], e = h[r + 1], len = e - b;
            if (len <= chunk) {
                vptr.push_back(e);
                vrow.push_back((int32_t)r);
            } else {
                const int32_t k = (len + chunk - 1) / chunk;
                srow.push_back((int32_t)r);
                sfirst.push_back(slot);
                p->max_slots_per_row = std::max(p->max_slots_per_row, k);
                for (int32_t c = 0; c < k; ++c) {
                    vptr.push_back(std::min(b + (c + 1) * chunk, e));
                    vrow.push_back(-1 - slot);
                    ++slot;
                }
            }
        }
        sfirst.push_back(slot);
        p->n_slots = slot;
        auto upload = [&](int32_t** d, const std::vector<int32_t>& v) -> cudaError_t {
            cudaError_t e = cudaMalloc((void**)d, sizeof(int32_t) * std::max<size_t>(1, v.size()));
            if (e != cudaSuccess) return e;
            return cudaMemcpyAsync(*d, v.data(), sizeof(int32_t) * v.size(), cudaMemcpyHostToDevice, st);
        };
        cudaError_t e = upload(&p->d_vptr, vptr);
        if (e == cudaSuccess) e = upload(&p->d_vrow, vrow);
        if (e == cudaSuccess) e = upload(&p->d_split_row, srow);
        if (e == cudaSuccess) e = upload(&p->d_split_first, sfirst);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // host vectors die at scope exit
        if (e != cudaSuccess) {
            gmr::set_error("gmr_spmm_plan_create: %s", cudaGetErrorString(e));
            gmr_spmm_plan_destroy(p);
            return GMR_ERR_CUDA;
        }
    }
    *out = p;
    return GMR_OK;
}

extern "C" int gmr_spmm_plan_destroy(gmr_spmm_plan_t* p)
{
    if (p == nullptr) return GMR_OK;
    cudaFree(p->d_vptr);
    cudaFree(p->d_vrow);
    cudaFree(p->d_split_row);
    cudaFree(p->d_split_first);
    delete p;
    return GMR_OK;
}

extern "C" int gmr_spmm_plan_stats(const gmr_spmm_plan_t* p, int64_t* n_chunks, int64_t* n_split_rows, int64_t* nnz)
{
    GMR_REQUIRE(p != nullptr, "gmr_spmm_plan_stats: null plan");
    if (n_chunks) *n_chunks = p->n_vrows;
    if (n_split_rows) *n_split_rows = p->n_split_rows;
    if (nnz) *nnz = p->nnz;
    return GMR_OK;
}

extern "C" int64_t gmr_spmm_workspace_bytes(const gmr_spmm_plan_t* p, int32_t D)
{
    if (p == nullptr || D < 1) return 0;
    return gmr::align_up(p->n_slots * (int64_t)D * (int64_t)sizeof(float), 256);
}

extern "C" int gmr_spmm_csr_f32(const gmr_spmm_plan_t* plan, const int32_t* rowptr, const int32_t* col,
                                const float* val, const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t D,
                                float alpha, float beta, void* workspace, int64_t workspace_bytes, void* stream)
{
    gmr::PushArgs push{nullptr, 0, 0};
    return gmr::spmm_common(plan, rowptr, col, val, X, ldx, Y, ldy, D, alpha, beta, push, false, workspace,
                            workspace_bytes, stream);
}

extern "C" int gmr_spmm_csr_f32_push(const gmr_spmm_plan_t* plan, const int32_t* rowptr, const int32_t* col,
                                     const float* val, const float* X, int64_t ldx, float* const* y_peers,
                                     int32_t n_peers, int64_t row_offset, int64_t ldy, int32_t D, float alpha,
                                     void* workspace, int64_t workspace_bytes, void* stream)
{
    GMR_REQUIRE(y_peers != nullptr && n_peers >= 1, "gmr_spmm_csr_f32_push: need at least one destination");
    GMR_REQUIRE(row_offset >= 0, "gmr_spmm_csr_f32_push: negative row offset");
    gmr::PushArgs push{y_peers, n_peers, row_offset};
    return gmr::spmm_common(plan, rowptr, col, val, X, ldx, nullptr, ldy, D, alpha, 0.f, push, true, workspace,
                            workspace_bytes, stream);
}

extern "C" int gmr_rows_axpby_norm_f32(const float* x, int64_t ldx, const float* y, int64_t ldy, const float* z,
                                       int64_t ldz, float* out, int64_t ldo, int64_t n_rows, int32_t D, float a,
                                       float b, float c, float eps, void* stream)
{
    GMR_REQUIRE(x != nullptr && out != nullptr, "gmr_rows_axpby_norm_f32: null x/out");
    GMR_REQUIRE(D >= 1 && n_rows >= 0, "gmr_rows_axpby_norm_f32: bad shape");
    if (n_rows == 0) return GMR_OK;
    const int wpb = 8;
    auto al16 = [](const float* p, int64_t ld) { return p == nullptr || ((uintptr_t)p % 16 == 0 && ld % 4 == 0); };
    if (D == 64 && al16(x, ldx) && al16(y, ldy) && al16(z, ldz) && al16(out, ldo)) {
        const int64_t rows_per_block = 2 * wpb;
        gmr::rows_axpby_norm_d64_kernel<<<(unsigned)((n_rows + rows_per_block - 1) / rows_per_block), wpb * 32, 0,
                                         (cudaStream_t)stream>>>(x, ldx, y, ldy, z, ldz, out, ldo, n_rows, a, b, c, eps);
        GMR_LAUNCH_CHECK();
        return GMR_OK;
    }
    gmr::rows_axpby_norm_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        x, ldx, y, ldy, z, ldz, out, ldo, n_rows, D, a, b, c, eps);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

extern "C" int gmr_rows_normalize_mix_f32(const float* x1, int64_t ld1, const float* x2, int64_t ld2, const float* y,
                                          int64_t ldy, float* out, int64_t ldo, int64_t n_rows, int32_t D, float w1,
                                          float w2, float slope, float eps, void* stream)
{
    GMR_REQUIRE(x1 != nullptr && x2 != nullptr && out != nullptr, "gmr_rows_normalize_mix_f32: null x1/x2/out");
    GMR_REQUIRE(D >= 1 && D <= 256 && n_rows >= 0, "gmr_rows_normalize_mix_f32: D must be in [1, 256] (got %d)", D);
    GMR_REQUIRE(ldo >= (y ? 2 * D : D), "gmr_rows_normalize_mix_f32: ldo too small");
    if (n_rows == 0) return GMR_OK;
    const int wpb = 8;
    gmr::rows_normalize_mix_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        x1, ld1, x2, ld2, y, ldy, out, ldo, n_rows, D, w1, w2, slope, eps);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

extern "C" int gmr_rows_push_f32(const float* src, int64_t ld, int64_t n_rows, int32_t D, float* const* y_peers,
                                 int32_t n_peers, int64_t row_offset, int64_t ldy, void* stream)
{
    GMR_REQUIRE(y_peers != nullptr && n_peers >= 1, "gmr_rows_push_f32: need at least one destination");
    GMR_REQUIRE(n_rows >= 0 && D >= 4 && D % 4 == 0 && ld % 4 == 0 && ldy % 4 == 0 && row_offset >= 0,
                "gmr_rows_push_f32: D, ld and ldy must be multiples of 4 (got D=%d ld=%lld ldy=%lld)", D, (long long)ld,
                (long long)ldy);
    if (n_rows == 0) return GMR_OK;
    GMR_REQUIRE(src != nullptr && (uintptr_t)src % 16 == 0, "gmr_rows_push_f32: src must be 16-byte aligned");
    const int64_t total = n_rows * (D / 4);
    int64_t blocks = (total + 255) / 256;
    // NVLink-bound: a small grid saturates the links and leaves the SMs to the SpMMs running beside it on the
    // compute stream (a full-size grid slowed those by 2x while waiting on store back-pressure)
    const int64_t cap = (int64_t)gmr::sm_count() / 2;
    if (blocks > cap) blocks = cap;
    gmr::rows_push_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(src), ld / 4, n_rows, D / 4, y_peers, n_peers, row_offset, ldy / 4);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}
