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

#include <cub/cub.cuh>

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
    int32_t* d_order = nullptr;        // [n_vrows] virtual rows by descending length (narrow kernels; built on demand)
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

// Packed form for rows of at most 32 Vec (D <= 128 floats with float4): a CTA takes 256 / (8 DV) split rows at once, 8 DV
// threads per row (sub-range w, column vc).  One CTA per split row left all but 8 DV of its 256 threads idle -- with tens
// of thousands of split rows that cost 30-90 us per product (0.18 ms of the 13.7 ms single-GPU step, 15 % of a narrow
// pass).  Same sub-ranges, same order of additions: same bits as spmm_reduce_kernel.
template <typename Vec>
__global__ void __launch_bounds__(kWarps * 32)
    spmm_reduce_packed_kernel(const int32_t* __restrict__ split_row, const int32_t* __restrict__ split_first,
                              const float* __restrict__ partial, float* __restrict__ Y, int64_t ldy, int32_t DV, int32_t n_split,
                              int32_t rows_per_cta, float alpha, float beta)
{
    using Ops = VecOps<Vec>;
    constexpr int VEC = sizeof(Vec) / 4;
    extern __shared__ float4 red_smem[];  // [rows_per_cta][kWarps][DV] Vec
    Vec* red = reinterpret_cast<Vec*>(red_smem);
    const int T = kWarps * DV;
    const int lr = threadIdx.x / T, t = threadIdx.x - lr * T;
    const int w = t / DV, vc = t - w * DV;
    const int s = blockIdx.x * rows_per_cta + lr;
    const bool active = lr < rows_per_cta && s < n_split;
    int row = 0;
    if (active) {
        row = split_row[s];
        const int first = split_first[s], n = split_first[s + 1] - first;
        const int per = (n + kWarps - 1) / kWarps;
        const int lo = min(n, w * per), hi = min(n, lo + per);
        const Vec* base = reinterpret_cast<const Vec*>(partial + (int64_t)first * ((int64_t)DV * VEC));
        Vec a = Ops::zero();
        for (int k = lo; k < hi; ++k) a = Ops::add(a, base[(int64_t)k * DV + vc]);
        red[(lr * kWarps + w) * DV + vc] = a;
    }
    __syncthreads();
    if (active && w == 0) {
        Vec a = red[(lr * kWarps) * DV + vc];
#pragma unroll
        for (int w2 = 1; w2 < kWarps; ++w2) a = Ops::add(a, red[(lr * kWarps + w2) * DV + vc]);
        Vec* y = reinterpret_cast<Vec*>(Y + (int64_t)row * ldy) + vc;
        *y = (beta == 0.f) ? Ops::scale(alpha, a) : Ops::axpby(alpha, a, beta, *y);
    }
}

// the split-row reduction of a product into Y (not the push form): packed when a row fits 32 threads per sub-range
template <typename Vec>
static int launch_reduce(const gmr_spmm_plan* plan, const float* partial, float* Y, int64_t ldy, int32_t DV, float alpha,
                         float beta, cudaStream_t st)
{
    if (plan->n_split_rows == 0) return GMR_OK;
    if (kWarps * DV <= kWarps * 32) {
        const int rows_per_cta = (kWarps * 32) / (kWarps * DV);
        const size_t smem = (size_t)rows_per_cta * kWarps * DV * sizeof(Vec);
        const unsigned grid = (unsigned)((plan->n_split_rows + rows_per_cta - 1) / rows_per_cta);
        spmm_reduce_packed_kernel<Vec><<<grid, kWarps * 32, smem, st>>>(plan->d_split_row, plan->d_split_first, partial, Y, ldy, DV,
                                                                       (int32_t)plan->n_split_rows, rows_per_cta, alpha, beta);
        GMR_LAUNCH_CHECK();
        return GMR_OK;
    }
    const size_t smem = (size_t)kWarps * DV * sizeof(Vec);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(spmm_reduce_kernel<Vec, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("row width of %d vectors too large for the split-row reduction (%zu B shared)", DV, smem);
            return GMR_ERR_UNSUPPORTED;
        }
    }
    PushArgs push{nullptr, 0, 0};
    spmm_reduce_kernel<Vec, false><<<(unsigned)plan->n_split_rows, kWarps * 32, smem, st>>>(plan->d_split_row, plan->d_split_first,
                                                                                           partial, Y, ldy, DV, alpha, beta, push);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
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


// ---- narrow rows (column-sharded propagation: a rank owns D = 8 / 16 / 32 / 64 of the embedding columns) ------------
// A row of D floats needs only D / 4 lanes, so a warp works on RPW = 32 / (D/4 * CHAINS) virtual rows at once.  The
// per-column operation order is the wide kernels': CHAINS = 2 keeps one sequential fmaf chain over the even and one over
// the odd nonzeros of the virtual row and adds them at the end (what spmm_vrow_kernel does for D <= 64 and the short-row
// kernel for D = 64); CHAINS = 1 is the single chain of the 128-column pass.  A column slice of the product therefore
// has the SAME BITS as the corresponding columns of the wide product, which is what lets dist.ColShardedDiffMM shard the
// propagation by embedding column and still return the single-GPU result.  Virtual rows are taken in descending-length
// order (plan->d_order) so that the RPW rows a warp walks together have (almost) equal lengths; the LPR lanes of a row
// load LPR consecutive (col, val) pairs with one instruction and hand them round by shuffle.
template <int DC4, int CHAINS, bool IDENT>
__global__ void __launch_bounds__(kWarps * 32, 5)
    spmm_narrow_kernel(const int32_t* __restrict__ vptr, const int32_t* __restrict__ vrow, const int32_t* __restrict__ order,
                       const int32_t* __restrict__ col, const float* __restrict__ val, const float* __restrict__ X,
                       int64_t ldx, float* __restrict__ Y, int64_t ldy, float* __restrict__ partial, int64_t n_vrows,
                       float alpha, float beta)
{
    using Ops = VecOps<float4>;
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int LPR = DC4 * CHAINS;   // lanes per virtual row
    constexpr int RPW = 32 / LPR;       // virtual rows per warp
    constexpr int U = 4;                // gathers in flight per lane
    constexpr int PER = U * CHAINS;     // nonzeros of a row consumed per iteration
    constexpr int NLOAD = (PER + LPR - 1) / LPR;
    static_assert(LPR <= 32 && (PER % LPR == 0 || LPR % PER == 0), "lane layout");
    const int lane = threadIdx.x & 31;
    const int rs = lane / LPR, within = lane % LPR, chain = within / DC4, sub = within % DC4;
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const float4* __restrict__ Xl = reinterpret_cast<const float4*>(X) + sub;
    const int64_t ldv = ldx >> 2;
    const int64_t n_groups = (n_vrows + RPW - 1) / RPW;
    const int64_t n_warps_grid = (int64_t)gridDim.x * kWarps;
    for (int64_t grp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); grp < n_groups; grp += n_warps_grid) {
        const int64_t idx = grp * RPW + rs;
        const bool live = idx < n_vrows;
        int b = 0, e = 0, dst = 0;
        if (live) {
            const int v = order[idx];
            b = vptr[v];
            e = vptr[v + 1];
            dst = IDENT ? v : vrow[v];
        }
        const int n = e - b;
        int nmax = n;
#pragma unroll
        for (int m = LPR; m < 32; m <<= 1) nmax = max(nmax, __shfl_xor_sync(kFull, nmax, m));
        const bool to_y = live && (IDENT || dst >= 0);
        float4* yp = reinterpret_cast<float4*>(Y + (int64_t)(to_y ? dst : 0) * ldy) + sub;
        float4 yold = Ops::zero();
        if (to_y && chain == 0 && beta != 0.f) yold = *yp;
        float4 acc = Ops::zero();
        for (int base = 0; base < nmax; base += PER) {
            int c[NLOAD];
            float w[NLOAD];
#pragma unroll
            for (int l = 0; l < NLOAD; ++l) {
                const int p = base + l * LPR + within;
                c[l] = 0;
                w[l] = 0.f;
                if (p < n && (NLOAD > 1 || within < PER)) {
                    c[l] = ld_stream_s32(col + b + p, pol_stream);
                    w[l] = ld_stream_f32(val + b + p, pol_stream);
                }
            }
            float4 xv[U];
            float ww[U];
            bool ok[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int pos = q * CHAINS + chain;             // position inside this iteration
                const int l = (q * CHAINS) / LPR;               // which index load holds it (compile time)
                const int cj = __shfl_sync(kFull, c[l], pos % LPR, LPR);
                ww[q] = __shfl_sync(kFull, w[l], pos % LPR, LPR);
                ok[q] = base + pos < n;
                xv[q] = ok[q] ? Ops::gather(Xl + (int64_t)cj * ldv, pol_keep) : Ops::zero();
            }
#pragma unroll
            for (int q = 0; q < U; ++q)
                if (ok[q]) Ops::fma(acc, ww[q], xv[q]);
        }
        if (CHAINS == 2) acc = Ops::xor_add(acc, DC4);   // even chain + odd chain
        if (chain == 0) {
            if (to_y)
                *yp = (beta == 0.f) ? Ops::scale(alpha, acc) : Ops::axpby(alpha, acc, beta, yold);
            else if (live)
                reinterpret_cast<float4*>(partial + (int64_t)(-1 - dst) * (DC4 * 4))[sub] = acc;
        }
    }
}


// ---- sliced-ELL form of the narrow kernel ---------------------------------------------------------------------------
// The gathers of a column shard are bound by the L1 request rate (~1 cache line per clock and SM, measured:
// tools/gather_narrow_bench.cu -- a 32-byte row costs as much as a 64-byte one), so every request the CSR stream takes
// is a gather lost: eight rows x 16-byte pieces of `col` and of `val` per warp load are 32 requests per 64 nonzeros, half
// as many as the gathers themselves.  The SELL snapshot stores the (col, val) pairs of the RPW rows a warp walks together
// (same descending-length order as above) interleaved per iteration: block `it` of a group holds entries
// [it * 8, it * 8 + 8) of each of its rows, row after row (512 contiguous bytes per warp with eight rows), nothing on the
// path chases a row pointer, and the next entries are fetched while the current gathers are in flight.  Same per-column
// operation order, same bits.
constexpr int kSellPer = 8;   // entries of a row per iteration

struct SellView {
    const uint32_t* goff;   // [n_groups + 1] first block of each group
    const int2* meta;       // [n_groups * RPW] (dst, n) per row slot; n = 0 for padding slots
    const uint2* cv;        // [n_blocks + 1][RPW][8] (col, val bits)
    int64_t n_groups;
};

template <int DC4, int CHAINS, int MINB>
__global__ void __launch_bounds__(kWarps * 32, MINB)
    spmm_sell_kernel(SellView sv, const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy,
                     float* __restrict__ partial, float alpha, float beta)
{
    using Ops = VecOps<float4>;
    constexpr int LPR = DC4 * CHAINS;
    constexpr int RPW = 32 / LPR;
    constexpr int PER = kSellPer;
    constexpr int QPB = 2 / CHAINS;   // quads (4 steps of one chain) per 8-entry block: 1 with two chains, 2 with one
    const int lane = threadIdx.x & 31;
    const int rs = lane / LPR, within = lane % LPR, chain = within / DC4, sub = within % DC4;
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const float4* __restrict__ Xl = reinterpret_cast<const float4*>(X) + sub;
    const int64_t ldv = ldx >> 2;
    const int64_t n_warps_grid = (int64_t)gridDim.x * kWarps;
    int64_t grp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (grp >= sv.n_groups) return;
    // A lane reads the four entries of its chain's next quad itself (two 128-bit loads; the DC4 lanes of a chain read the
    // same addresses and the snapshot stores a block chain-major: [p0 p2 p4 p6 | p1 p3 p5 p7] with two chains), so no
    // index ever crosses lanes: on this kernel a shuffle costs an LSU wavefront exactly like a gathered row does, and the
    // data pipe's one wavefront per clock is the bound (ncu: l1tex data-pipe wavefronts 95 % of peak).
    auto load_quad = [&](const uint2* p, uint4& e01, uint4& e23) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=r"(e01.x), "=r"(e01.y), "=r"(e01.z), "=r"(e01.w) : "l"(p), "l"(pol_stream));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=r"(e23.x), "=r"(e23.y), "=r"(e23.z), "=r"(e23.w) : "l"(p + 2), "l"(pol_stream));
    };
    auto load_header = [&](int64_t g, int2& m, uint32_t& o, uint32_t& oe) {
        m = make_int2(0x7fffffff, 0);
        o = oe = 0;
        if (g < sv.n_groups) {
            m = sv.meta[g * RPW + rs];
            o = sv.goff[g];
            oe = sv.goff[g + 1];
        }
    };
    // quad j of a group starting at block `o`: block o + j / QPB, entries [chain * 4 + (j % QPB) * 4, +4) of row slot rs
    auto quad_ptr = [&](uint32_t o, int j) {
        return sv.cv + ((int64_t)o + j / QPB) * (RPW * PER) + rs * PER + (CHAINS == 2 ? chain * 4 : (j % QPB) * 4);
    };
    int2 meta;
    uint32_t off, off_end;
    load_header(grp, meta, off, off_end);
    uint4 c01, c23;
    load_quad(quad_ptr(off, 0), c01, c23);
    while (true) {
        const int64_t gnext = grp + n_warps_grid;
        const bool has_next = gnext < sv.n_groups;
        int2 meta1;
        uint32_t off1, off_end1;
        load_header(gnext, meta1, off1, off_end1);   // the next group's header travels under this group's gathers
        const int n = meta.y, dst = meta.x;
        const bool live = dst != 0x7fffffff;
        const bool to_y = live && dst >= 0;
        float4* yp = reinterpret_cast<float4*>(Y + (int64_t)(to_y ? dst : 0) * ldy) + sub;
        float4 yold = Ops::zero();
        if (to_y && chain == 0 && beta != 0.f) yold = *yp;
        float4 acc = Ops::zero();
        const int nquad = (int)(off_end - off) * QPB;
        for (int j = 0; j < nquad; ++j) {
            // the next quad's entries (or the next group's first quad) load under this quad's gathers
            uint4 n01, n23;
            load_quad((j + 1 < nquad) ? quad_ptr(off, j + 1) : quad_ptr(off1, 0), n01, n23);
            const int pos0 = (CHAINS == 2) ? (j * 8 + chain) : (j * 4);   // row position of step 0; steps advance by CHAINS
            const uint32_t cc[4] = {c01.x, c01.z, c23.x, c23.z};
            const uint32_t wb[4] = {c01.y, c01.w, c23.y, c23.w};
            float4 xv[4];
            bool ok[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ok[q] = pos0 + q * CHAINS < n;
                xv[q] = ok[q] ? Ops::gather(Xl + (int64_t)cc[q] * ldv, pol_keep) : Ops::zero();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (ok[q]) Ops::fma(acc, __uint_as_float(wb[q]), xv[q]);
            c01 = n01;
            c23 = n23;
        }
        if (nquad == 0) load_quad(quad_ptr(off1, 0), c01, c23);   // an all-empty group: its successor's first quad
        if (CHAINS == 2) acc = Ops::xor_add(acc, DC4);
        if (chain == 0) {
            if (to_y)
                *yp = (beta == 0.f) ? Ops::scale(alpha, acc) : Ops::axpby(alpha, acc, beta, yold);
            else if (live)
                reinterpret_cast<float4*>(partial + (int64_t)(-1 - dst) * (DC4 * 4))[sub] = acc;
        }
        if (!has_next) break;
        grp = gnext;
        meta = meta1; off = off1; off_end = off_end1;
    }
}

// builders of the SELL snapshot
__global__ void sell_group_iters_kernel(const int32_t* __restrict__ vptr, const int32_t* __restrict__ order, int64_t n_vrows,
                                        int rpw, int64_t n_groups, uint32_t* __restrict__ iters)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int v = order[g * rpw];   // longest row of the group (descending order)
    const int n = vptr[v + 1] - vptr[v];
    iters[g] = (uint32_t)((n + kSellPer - 1) / kSellPer);
}

// one warp per row slot: meta + the row's entries into its group's blocks
template <bool IDENT>
__global__ void __launch_bounds__(256)
    sell_fill_kernel(const int32_t* __restrict__ vptr, const int32_t* __restrict__ vrow, const int32_t* __restrict__ order,
                     const int32_t* __restrict__ col, const float* __restrict__ val, int64_t n_vrows, int rpw, int chains,
                     int64_t n_slots, const uint32_t* __restrict__ goff, int2* __restrict__ meta, uint2* __restrict__ cv)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_slots) return;
    if (i >= n_vrows) {
        if (lane == 0) meta[i] = make_int2(0x7fffffff, 0);   // padding slot of the last group
        return;
    }
    const int v = order[i];
    const int b = vptr[v], n = vptr[v + 1] - b;
    if (lane == 0) meta[i] = make_int2(IDENT ? v : vrow[v], n);
    const int64_t g = i / rpw;
    const int rs = (int)(i - g * rpw);
    uint2* base = cv + (int64_t)goff[g] * (rpw * kSellPer) + rs * kSellPer;
    for (int p = lane; p < n; p += 32) {
        const int k = p % kSellPer;
        const int slot = (chains == 2) ? ((k & 1) * 4 + (k >> 1)) : k;   // two chains: even entries first, then the odd ones
        base[(int64_t)(p / kSellPer) * (rpw * kSellPer) + slot] = make_uint2((uint32_t)col[b + p], __float_as_uint(val[b + p]));
    }
}

// iota for the order sort
__global__ void iota_len_kernel(const int32_t* __restrict__ vptr, int64_t n, uint32_t* __restrict__ len, int32_t* __restrict__ id)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        len[i] = (uint32_t)(vptr[i + 1] - vptr[i]);
        id[i] = (int32_t)i;
    }
}

template <int DC4, int CHAINS>
static int launch_narrow(const gmr_spmm_plan* plan, const int32_t* rowptr, const int32_t* col, const float* val,
                         const float* X, int64_t ldx, float* Y, int64_t ldy, float* partial, float alpha, float beta,
                         cudaStream_t st)
{
    constexpr int RPW = 32 / (DC4 * CHAINS);
    const int64_t nv = plan->n_vrows;
    if (nv == 0) return GMR_OK;
    const int64_t groups = (nv + RPW - 1) / RPW;
    const int64_t want = (groups + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)sm_count() * 5;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (plan->d_vptr == nullptr)
        spmm_narrow_kernel<DC4, CHAINS, true><<<grid, kWarps * 32, 0, st>>>(rowptr, nullptr, plan->d_order, col, val, X, ldx, Y,
                                                                           ldy, partial, nv, alpha, beta);
    else
        spmm_narrow_kernel<DC4, CHAINS, false><<<grid, kWarps * 32, 0, st>>>(plan->d_vptr, plan->d_vrow, plan->d_order, col,
                                                                            val, X, ldx, Y, ldy, partial, nv, alpha, beta);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
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
    if (!PUSH) return launch_reduce<Vec>(plan, partial, Y, ldy, DV, alpha, beta, st);
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
        // neighbours first (1, 2, 4, 8): every aligned block of 2 / 4 / 8 lanes is a subtree, so a rank that owns a contiguous
        // block of columns can form its part of the sum and the parts combine to the same bits (gmr_rows_sumsq_f32)
        for (int m = 1; m < 16; m <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);   // within the 16-lane half
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

// D = 64 form of rows_normalize_mix_kernel: a half-warp owns a row (16 lanes x 128 bit), four rows in flight per
// half-warp (12 independent 16-byte loads per lane before the first use): 0.21 -> 0.09 ms for 500k rows
__global__ void __launch_bounds__(256)
    rows_normalize_mix_d64_kernel(const float* __restrict__ x1, int64_t ld1, const float* __restrict__ x2, int64_t ld2,
                                  const float* __restrict__ y, int64_t ldy, float* __restrict__ out, int64_t ldo,
                                  int64_t n_rows, float w1, float w2, float slope, float eps)
{
    const int lane = threadIdx.x & 31, sub = lane & 15, half = lane >> 4;
    const int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + half;
    float4 a[4], b[4], yv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t r = r0 + 2 * q;
        const bool ok = r < n_rows;
        a[q] = ok ? *reinterpret_cast<const float4*>(x1 + r * ld1 + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        b[q] = ok ? *reinterpret_cast<const float4*>(x2 + r * ld2 + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        yv[q] = (ok && y != nullptr) ? *reinterpret_cast<const float4*>(y + r * ldy + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t r = r0 + 2 * q;
        float4 u = a[q], v = b[q];
        u.x = u.x > 0.f ? u.x : u.x * slope; u.y = u.y > 0.f ? u.y : u.y * slope;
        u.z = u.z > 0.f ? u.z : u.z * slope; u.w = u.w > 0.f ? u.w : u.w * slope;
        v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
        v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
        float sa = fmaf(u.x, u.x, fmaf(u.y, u.y, fmaf(u.z, u.z, u.w * u.w)));
        float sb = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
        for (int m = 8; m > 0; m >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, m);
            sb += __shfl_xor_sync(0xffffffffu, sb, m);
        }
        if (r >= n_rows) continue;
        const float ia = w1 / fmaxf(sqrtf(sa), eps), ib = w2 / fmaxf(sqrtf(sb), eps);
        float4 z;
        z.x = fmaf(ia, u.x, ib * v.x); z.y = fmaf(ia, u.y, ib * v.y);
        z.z = fmaf(ia, u.z, ib * v.z); z.w = fmaf(ia, u.w, ib * v.w);
        *reinterpret_cast<float4*>(out + r * ldo + 4 * sub) = z;
        if (y != nullptr) {
            float4 zy;
            zy.x = z.x + yv[q].x; zy.y = z.y + yv[q].y; zy.z = z.z + yv[q].z; zy.w = z.w + yv[q].w;
            *reinterpret_cast<float4*>(out + r * ldo + 64 + 4 * sub) = zy;
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
    cudaFree(p->d_order);
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


extern "C" int gmr_spmm_plan_enable_narrow(gmr_spmm_plan_t* p, const int32_t* rowptr, void* stream)
{
    GMR_REQUIRE(p != nullptr, "gmr_spmm_plan_enable_narrow: null plan");
    if (p->d_order != nullptr || p->n_vrows == 0) return GMR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = p->n_vrows;
    const int32_t* vptr = p->d_vptr ? p->d_vptr : rowptr;
    GMR_REQUIRE(vptr != nullptr, "gmr_spmm_plan_enable_narrow: null rowptr");
    uint32_t *len = nullptr, *len_out = nullptr;
    int32_t *id = nullptr, *order = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, len, len_out, id, order, (int)n, 0, 32, st);
    if (e == cudaSuccess) e = cudaMalloc((void**)&len, sizeof(uint32_t) * n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&len_out, sizeof(uint32_t) * n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&id, sizeof(int32_t) * n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&order, sizeof(int32_t) * n);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) {
        gmr::iota_len_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(vptr, n, len, id);
        e = cudaGetLastError();
    }
    // stable: virtual rows of equal length keep their row order
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, len, len_out, id, order, (int)n, 0, 32, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(len);
    cudaFree(len_out);
    cudaFree(id);
    cudaFree(tmp);
    if (e != cudaSuccess) {
        cudaFree(order);
        gmr::set_error("gmr_spmm_plan_enable_narrow: %s", cudaGetErrorString(e));
        return GMR_ERR_CUDA;
    }
    p->d_order = order;
    return GMR_OK;
}

extern "C" int gmr_spmm_narrow_f32(const gmr_spmm_plan_t* plan, const int32_t* rowptr, const int32_t* col, const float* val,
                                   const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t D, int32_t chains, float alpha,
                                   float beta, void* workspace, int64_t workspace_bytes, void* stream)
{
    GMR_REQUIRE(plan != nullptr, "gmr_spmm_narrow_f32: plan is null");
    GMR_REQUIRE((D == 8 || D == 16 || D == 32 || D == 64) && (chains == 1 || chains == 2),
                "gmr_spmm_narrow_f32: D must be 8, 16, 32 or 64 and chains 1 or 2 (got D=%d chains=%d)", D, chains);
    GMR_REQUIRE(!(D == 64 && chains == 2), "gmr_spmm_narrow_f32: D=64 with two chains is the wide kernel (gmr_spmm_csr_f32)");
    GMR_REQUIRE(ldx >= D && ldy >= D && ldx % 4 == 0 && ldy % 4 == 0, "gmr_spmm_narrow_f32: bad leading dimensions (%lld, %lld)",
                (long long)ldx, (long long)ldy);
    if (plan->n_rows == 0) return GMR_OK;
    GMR_REQUIRE(rowptr && X && Y && (uintptr_t)X % 16 == 0 && (uintptr_t)Y % 16 == 0, "gmr_spmm_narrow_f32: null or unaligned operand");
    GMR_REQUIRE(plan->nnz == 0 || (col && val), "gmr_spmm_narrow_f32: null col/val with nnz > 0");
    GMR_REQUIRE(plan->d_order != nullptr || plan->n_vrows == 0, "gmr_spmm_narrow_f32: call gmr_spmm_plan_enable_narrow first");
    const int64_t need = gmr_spmm_workspace_bytes(plan, D);
    if (need > 0 && (workspace == nullptr || workspace_bytes < need)) {
        gmr::set_error("gmr_spmm_narrow_f32: workspace of %lld bytes required, %lld given", (long long)need, (long long)workspace_bytes);
        return GMR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = (float*)workspace;
    int rc = GMR_ERR_INVALID;
#define GMR_NARROW(DC4, CH) \
    rc = gmr::launch_narrow<DC4, CH>(plan, rowptr, col, val, X, ldx, Y, ldy, partial, alpha, beta, st)
    if (D == 8 && chains == 2) GMR_NARROW(2, 2);
    else if (D == 8) GMR_NARROW(2, 1);
    else if (D == 16 && chains == 2) GMR_NARROW(4, 2);
    else if (D == 16) GMR_NARROW(4, 1);
    else if (D == 32 && chains == 2) GMR_NARROW(8, 2);
    else if (D == 32) GMR_NARROW(8, 1);
    else GMR_NARROW(16, 1);
#undef GMR_NARROW
    if (rc != GMR_OK) return rc;
    return gmr::launch_reduce<float4>(plan, partial, Y, ldy, D / 4, alpha, beta, st);
}

struct gmr_spmm_sell {
    const gmr_spmm_plan* base = nullptr;
    int32_t rpw = 0, chains = 0;
    int64_t n_groups = 0, n_blocks = 0;
    uint32_t* d_goff = nullptr;
    int2* d_meta = nullptr;
    uint2* d_cv = nullptr;
};

static int sell_fill(gmr_spmm_sell* s, const int32_t* rowptr, const int32_t* col, const float* val, cudaStream_t st)
{
    const gmr_spmm_plan* p = s->base;
    const int64_t n_slots = s->n_groups * s->rpw;
    GMR_CHECK_CUDA(cudaMemsetAsync(s->d_cv, 0, sizeof(uint2) * (size_t)(s->n_blocks + 1) * s->rpw * gmr::kSellPer, st));
    const unsigned grid = (unsigned)((n_slots * 32 + 255) / 256);
    if (p->d_vptr == nullptr)
        gmr::sell_fill_kernel<true><<<grid, 256, 0, st>>>(rowptr, nullptr, p->d_order, col, val, p->n_vrows, s->rpw, s->chains, n_slots,
                                                         s->d_goff, s->d_meta, s->d_cv);
    else
        gmr::sell_fill_kernel<false><<<grid, 256, 0, st>>>(p->d_vptr, p->d_vrow, p->d_order, col, val, p->n_vrows, s->rpw, s->chains,
                                                          n_slots, s->d_goff, s->d_meta, s->d_cv);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

extern "C" int gmr_spmm_sell_destroy(gmr_spmm_sell_t* s)
{
    if (s == nullptr) return GMR_OK;
    cudaFree(s->d_goff);
    cudaFree(s->d_meta);
    cudaFree(s->d_cv);
    delete s;
    return GMR_OK;
}

extern "C" int gmr_spmm_sell_create(gmr_spmm_sell_t** out, gmr_spmm_plan_t* plan, const int32_t* rowptr, const int32_t* col,
                                    const float* val, int32_t D, int32_t chains, void* stream)
{
    GMR_REQUIRE(out != nullptr && plan != nullptr, "gmr_spmm_sell_create: null argument");
    *out = nullptr;
    GMR_REQUIRE((D == 8 || D == 16 || D == 32 || D == 64) && (chains == 1 || chains == 2) && !(D == 64 && chains == 2),
                "gmr_spmm_sell_create: D must be 8, 16, 32 (chains 1 or 2) or 64 (chains 1); got D=%d chains=%d", D, chains);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = gmr_spmm_plan_enable_narrow(plan, rowptr, stream);
    if (rc != GMR_OK) return rc;
    auto* s = new gmr_spmm_sell();
    s->base = plan;
    s->rpw = 32 / ((D / 4) * chains);
    s->chains = chains;
    s->n_groups = (plan->n_vrows + s->rpw - 1) / s->rpw;
    if (s->n_groups == 0) {
        *out = s;
        return GMR_OK;
    }
    const int32_t* vptr = plan->d_vptr ? plan->d_vptr : rowptr;
    uint32_t* iters = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    const int ng = (int)s->n_groups;
    cudaError_t e = cudaMalloc((void**)&iters, sizeof(uint32_t) * (ng + 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_goff, sizeof(uint32_t) * (ng + 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_meta, sizeof(int2) * (size_t)ng * s->rpw);
    if (e == cudaSuccess) e = cudaMemsetAsync(iters, 0, sizeof(uint32_t) * (ng + 1), st);
    if (e == cudaSuccess) {
        gmr::sell_group_iters_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(vptr, plan->d_order, plan->n_vrows, s->rpw, ng, iters);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, iters, s->d_goff, ng + 1, st);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, iters, s->d_goff, ng + 1, st);
    uint32_t total = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, s->d_goff + ng, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(iters);
    cudaFree(tmp);
    s->n_blocks = total;
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_cv, sizeof(uint2) * (size_t)(s->n_blocks + 1) * s->rpw * gmr::kSellPer);
    if (e != cudaSuccess) {
        gmr::set_error("gmr_spmm_sell_create: %s", cudaGetErrorString(e));
        gmr_spmm_sell_destroy(s);
        return GMR_ERR_CUDA;
    }
    rc = sell_fill(s, rowptr, col, val, st);
    if (rc != GMR_OK) {
        gmr_spmm_sell_destroy(s);
        return rc;
    }
    *out = s;
    return GMR_OK;
}

extern "C" int gmr_spmm_sell_set_values(gmr_spmm_sell_t* s, const int32_t* rowptr, const int32_t* col, const float* val, void* stream)
{
    GMR_REQUIRE(s != nullptr, "gmr_spmm_sell_set_values: null snapshot");
    if (s->n_groups == 0) return GMR_OK;
    return sell_fill(s, rowptr, col, val, (cudaStream_t)stream);
}

extern "C" int64_t gmr_spmm_sell_bytes(const gmr_spmm_sell_t* s)
{
    if (s == nullptr) return 0;
    return (int64_t)sizeof(uint2) * (s->n_blocks + 1) * s->rpw * gmr::kSellPer + (int64_t)sizeof(int2) * s->n_groups * s->rpw +
           (int64_t)sizeof(uint32_t) * (s->n_groups + 1);
}

extern "C" int gmr_spmm_sell_f32(const gmr_spmm_sell_t* s, const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t D,
                                 int32_t chains, float alpha, float beta, void* workspace, int64_t workspace_bytes, void* stream)
{
    GMR_REQUIRE(s != nullptr, "gmr_spmm_sell_f32: null snapshot");
    GMR_REQUIRE((D == 8 || D == 16 || D == 32 || D == 64) && (chains == 1 || chains == 2) && !(D == 64 && chains == 2) &&
                    32 / ((D / 4) * chains) == s->rpw && chains == s->chains,
                "gmr_spmm_sell_f32: D=%d chains=%d does not match the snapshot (%d rows per warp, %d chains)", D, chains, s->rpw,
                s->chains);
    GMR_REQUIRE(ldx >= D && ldy >= D && ldx % 4 == 0 && ldy % 4 == 0, "gmr_spmm_sell_f32: bad leading dimensions (%lld, %lld)",
                (long long)ldx, (long long)ldy);
    const gmr_spmm_plan* plan = s->base;
    if (plan->n_rows == 0 || s->n_groups == 0) return GMR_OK;
    GMR_REQUIRE(X && Y && (uintptr_t)X % 16 == 0 && (uintptr_t)Y % 16 == 0, "gmr_spmm_sell_f32: null or unaligned operand");
    const int64_t need = gmr_spmm_workspace_bytes(plan, D);
    if (need > 0 && (workspace == nullptr || workspace_bytes < need)) {
        gmr::set_error("gmr_spmm_sell_f32: workspace of %lld bytes required, %lld given", (long long)need, (long long)workspace_bytes);
        return GMR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = (float*)workspace;
    gmr::SellView sv{s->d_goff, s->d_meta, s->d_cv, s->n_groups};
    const int64_t want = (s->n_groups + gmr::kWarps - 1) / gmr::kWarps;
    // resident CTAs per SM: 3 (no spills; measured 3-5 % faster) or 4 (64 registers, a few spilled words); GMR_SELL_MINB picks
    const char* mb = getenv("GMR_SELL_MINB");
    const int minb = (mb && atoi(mb) == 4) ? 4 : 3;
    const int64_t cap = (int64_t)gmr::sm_count() * minb;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
#define GMR_SELL(DC4, CH)                                                                                                   \
    do {                                                                                                                    \
        if (minb == 3)                                                                                                      \
            gmr::spmm_sell_kernel<DC4, CH, 3><<<grid, gmr::kWarps * 32, 0, st>>>(sv, X, ldx, Y, ldy, partial, alpha, beta);  \
        else                                                                                                                \
            gmr::spmm_sell_kernel<DC4, CH, 4><<<grid, gmr::kWarps * 32, 0, st>>>(sv, X, ldx, Y, ldy, partial, alpha, beta);  \
    } while (0)
    if (D == 8 && chains == 2) GMR_SELL(2, 2);
    else if (D == 8) GMR_SELL(2, 1);
    else if (D == 16 && chains == 2) GMR_SELL(4, 2);
    else if (D == 16) GMR_SELL(4, 1);
    else if (D == 32 && chains == 2) GMR_SELL(8, 2);
    else if (D == 32) GMR_SELL(8, 1);
    else GMR_SELL(16, 1);
#undef GMR_SELL
    GMR_LAUNCH_CHECK();
    return gmr::launch_reduce<float4>(plan, partial, Y, ldy, D / 4, alpha, beta, st);
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
    const bool v4 = D == 64 && ld1 % 4 == 0 && ld2 % 4 == 0 && ldo % 4 == 0 && (y == nullptr || ldy % 4 == 0) &&
                    (uintptr_t)x1 % 16 == 0 && (uintptr_t)x2 % 16 == 0 && (uintptr_t)out % 16 == 0 &&
                    (y == nullptr || (uintptr_t)y % 16 == 0);
    if (v4)
        gmr::rows_normalize_mix_d64_kernel<<<(unsigned)((n_rows + 8 * wpb - 1) / (8 * wpb)), wpb * 32, 0, (cudaStream_t)stream>>>(
            x1, ld1, x2, ld2, y, ldy, out, ldo, n_rows, w1, w2, slope, eps);
    else
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
