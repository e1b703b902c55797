// colshard.cu -- glue of the COLUMN-sharded propagation (dist.ColShardedDiffMM, sm_100a)
//
// With the embedding columns sharded over the ranks (rank g owns columns [g * dc, (g + 1) * dc) of every row) the
// SpMM layers need no collective at all: A X is column-separable.  What is left to exchange is small and lives here:
//   * gmr_cols_push_f32     the layout changes between row blocks and column blocks (the row-sharded projections going
//                           in, the user block / item table the scoring wants coming out): strided copies of column
//                           slices straight into the peers' buffers over NVLink peer memory
//   * gmr_rows_sumsq_f32    a rank's part of the squared row norms of F.normalize (GenMMRec/src/models/diffmm.py:166)
//   * gmr_rows_axpby_ss_f32 out = a x + b y + c z / max(sqrt(sum of the ranks' parts), eps): the parts are added as the
//                           balanced tree the single-GPU row kernel (rows_axpby_norm_d64_kernel) uses, so the bits match
//   * gmr_peer_barrier      a stream-ordered cross-rank barrier through flags in peer memory (replaces a 1-element NCCL
//                           all-reduce per exchange)
#include "common.cuh"

namespace gmr {

constexpr int kMaxPeers = 16;

// rows [0, n_rows) x dc columns -> every peer (routed == false) or the peer whose row range holds the row (routed).
template <bool ROUTED>
__global__ void __launch_bounds__(256)
    cols_push_kernel(const float4* __restrict__ src, int64_t ld4, int64_t n_rows, int32_t dc4, int64_t src_col0_4,
                     int64_t src_col_step4, float* const* __restrict__ y_peers, int32_t n_peers,
                     const int64_t* __restrict__ row_bounds, int64_t dst_row_offset, int64_t ldy4, int64_t dst_col0_4,
                     int32_t n_seg, int64_t src_seg_step4, int64_t dst_seg_step4)
{
    __shared__ int64_t rb[kMaxPeers + 1];
    if (ROUTED) {
        if (threadIdx.x <= n_peers) rb[threadIdx.x] = row_bounds[threadIdx.x];
        __syncthreads();
    }
    // piece index = (row, segment, 16-byte column): consecutive threads write consecutive destination addresses whenever
    // the destination row is exactly the segments side by side (contiguous remote stores: NVLink moves 32-byte granules
    // scattered at a 256-byte pitch at a third of the rate of full lines, measured 0.39 vs 0.15 ms for the item table)
    const int per_row = n_seg * dc4;
    const int64_t total = n_rows * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / per_row;
        const int rem = (int)(i - r * per_row);
        const int sg = rem / dc4, c = rem - sg * dc4;
        const int64_t so = r * ld4 + src_col0_4 + sg * src_seg_step4 + c;
        const int64_t dcol = dst_col0_4 + sg * dst_seg_step4 + c;
        if (ROUTED) {
            int p = 0;
            while (p + 1 < n_peers && r >= rb[p + 1]) ++p;
            if (r < rb[0] || r >= rb[n_peers]) continue;
            const float4 v = src[so + p * src_col_step4];
            reinterpret_cast<float4*>(y_peers[p])[(dst_row_offset + r - rb[p]) * ldy4 + dcol] = v;
        } else {
            for (int p = 0; p < n_peers; ++p) {
                const float4 v = src[so + p * src_col_step4];
                reinterpret_cast<float4*>(y_peers[p])[(dst_row_offset + r) * ldy4 + dcol] = v;
            }
        }
    }
}

// out[r, p * dc + c] = slabs[p][r, c]: the column slabs the peers stored contiguously, side by side as full rows
__global__ void __launch_bounds__(256)
    slabs_to_rows_kernel(const float4* __restrict__ slabs, int32_t n_slabs, int64_t slab_stride4, int64_t n_rows, int32_t dc4,
                         float4* __restrict__ out, int64_t ldo4)
{
    const int per_row = n_slabs * dc4;
    const int64_t total = n_rows * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / per_row;
        const int rem = (int)(i - r * per_row);
        const int p = rem / dc4, c = rem - p * dc4;
        out[r * ldo4 + rem] = slabs[p * slab_stride4 + r * dc4 + c];
    }
}

// out[r] = sum_d z[r, d]^2 for D = 4 * LPR columns, LPR lanes per row: per lane fmaf(x, x, fmaf(y, y, fmaf(z, z, w * w))),
// then neighbours first -- the subtree of rows_axpby_norm_d64_kernel's 16-lane butterfly that covers these columns.
template <int LPR>
__global__ void __launch_bounds__(256)
    rows_sumsq_kernel(const float* __restrict__ z, int64_t ldz, int64_t n_rows, float* __restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = t / LPR;
    const int sub = (int)(t % LPR);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < n_rows) v = *reinterpret_cast<const float4*>(z + r * ldz + 4 * sub);
    float ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
    for (int m = 1; m < LPR; m <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
    if (r < n_rows && sub == 0) out[r] = ss;
}

// one thread per float4 of a row
__global__ void __launch_bounds__(256)
    rows_axpby_ss_kernel(const float* x, int64_t ldx, const float* y, int64_t ldy, const float* z, int64_t ldz,   // out may alias
                         const float* __restrict__ ss_parts, int32_t n_parts, int64_t ss_stride, float* out, int64_t ldo,
                         int64_t n_rows, int32_t d4, float a, float b, float c, float eps)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * d4) return;
    const int64_t r = i / d4;
    const int col = 4 * (int)(i - r * d4);
    const float4 xv = *reinterpret_cast<const float4*>(x + r * ldx + col);
    float4 o = make_float4(a * xv.x, a * xv.y, a * xv.z, a * xv.w);
    if (y != nullptr) {
        const float4 yv = *reinterpret_cast<const float4*>(y + r * ldy + col);
        o.x = fmaf(b, yv.x, o.x); o.y = fmaf(b, yv.y, o.y); o.z = fmaf(b, yv.z, o.z); o.w = fmaf(b, yv.w, o.w);
    }
    if (z != nullptr) {
        float p[kMaxPeers];
#pragma unroll
        for (int k = 0; k < kMaxPeers; ++k) p[k] = k < n_parts ? ss_parts[k * ss_stride + r] : 0.f;
        // balanced tree, neighbours first (n_parts is a power of two)
#pragma unroll
        for (int s = 1; s < kMaxPeers; s <<= 1)
#pragma unroll
            for (int k = 0; k < kMaxPeers; k += 2 * s)
                if (k + s < n_parts) p[k] += p[k + s];
        const float inv = c / fmaxf(sqrtf(p[0]), eps);
        const float4 zv = *reinterpret_cast<const float4*>(z + r * ldz + col);
        o.x = fmaf(inv, zv.x, o.x); o.y = fmaf(inv, zv.y, o.y); o.z = fmaf(inv, zv.z, o.z); o.w = fmaf(inv, zv.w, o.w);
    }
    *reinterpret_cast<float4*>(out + r * ldo + col) = o;
}

// Cross-rank barrier through peer memory.  state[0] = barriers passed (local), state[1] = timeouts seen (local).
// Thread t tells rank t "rank `me` has arrived at barrier number e" (a store into rank t's flag array, after a system-scope
// fence: everything this rank's earlier kernels stored into peer buffers is visible first) and waits until rank t has said
// the same here.  A peer that never arrives is given up on after ~4 s (counted in state[1]; the host checks it).
__global__ void __launch_bounds__(32)
    peer_barrier_kernel(uint32_t* const* __restrict__ flags_peers, int32_t me, int32_t n_ranks, uint32_t* __restrict__ state)
{
    const int t = threadIdx.x;
    const uint32_t e = state[0] + 1u;
    __syncwarp();
    if (t < n_ranks) {
        __threadfence_system();
        volatile uint32_t* theirs = flags_peers[t] + me;
        *theirs = e;
        volatile uint32_t* mine = flags_peers[me] + t;
        unsigned long long t0 = 0, now = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int32_t)(*mine - e) < 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 4000000000ull) {
                atomicAdd(&state[1], 1u);
                break;
            }
        }
        __threadfence_system();
    }
    __syncwarp();
    if (t == 0) state[0] = e;
}

}  // namespace gmr

extern "C" int gmr_cols_push_f32(const float* src, int64_t ld, int64_t n_rows, int32_t dc, int64_t src_col0,
                                 int64_t src_col_step, float* const* y_peers, int32_t n_peers, const int64_t* row_bounds,
                                 int64_t dst_row_offset, int64_t ldy, int64_t dst_col0, int32_t n_seg, int64_t src_seg_step,
                                 int64_t dst_seg_step, void* stream)
{
    GMR_REQUIRE(n_seg >= 1 && src_seg_step >= 0 && dst_seg_step >= 0 && src_seg_step % 4 == 0 && dst_seg_step % 4 == 0,
                "gmr_cols_push_f32: bad segment description (%d segments)", n_seg);
    GMR_REQUIRE(y_peers != nullptr && n_peers >= 1 && n_peers <= gmr::kMaxPeers, "gmr_cols_push_f32: 1..%d destinations (got %d)",
                gmr::kMaxPeers, n_peers);
    GMR_REQUIRE(n_rows >= 0 && dc >= 4 && dc % 4 == 0 && ld % 4 == 0 && ldy % 4 == 0 && src_col0 % 4 == 0 && src_col_step % 4 == 0 &&
                    dst_col0 % 4 == 0 && src_col0 >= 0 && src_col_step >= 0 && dst_col0 >= 0 && dst_row_offset >= 0,
                "gmr_cols_push_f32: widths, offsets and leading dimensions must be non-negative multiples of 4");
    if (n_rows == 0) return GMR_OK;
    GMR_REQUIRE(src != nullptr && (uintptr_t)src % 16 == 0, "gmr_cols_push_f32: src must be 16-byte aligned");
    const int64_t total = n_rows * (dc / 4) * n_seg;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)gmr::sm_count() * (n_peers == 1 ? 8 : 1);   // remote stores: NVLink-bound, leave the SMs alone
    if (blocks > cap) blocks = cap;
    if (row_bounds != nullptr)
        gmr::cols_push_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(src), ld / 4, n_rows, dc / 4, src_col0 / 4, src_col_step / 4, y_peers, n_peers,
            row_bounds, dst_row_offset, ldy / 4, dst_col0 / 4, n_seg, src_seg_step / 4, dst_seg_step / 4);
    else
        gmr::cols_push_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(src), ld / 4, n_rows, dc / 4, src_col0 / 4, src_col_step / 4, y_peers, n_peers,
            nullptr, dst_row_offset, ldy / 4, dst_col0 / 4, n_seg, src_seg_step / 4, dst_seg_step / 4);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

extern "C" int gmr_slabs_to_rows_f32(const float* slabs, int32_t n_slabs, int64_t slab_stride, int64_t n_rows, int32_t dc, float* out,
                                     int64_t ldo, void* stream)
{
    GMR_REQUIRE(n_slabs >= 1 && n_rows >= 0 && dc >= 4 && dc % 4 == 0 && slab_stride % 4 == 0 && ldo % 4 == 0 && ldo >= (int64_t)n_slabs * dc,
                "gmr_slabs_to_rows_f32: bad shape (slabs=%d dc=%d ldo=%lld)", n_slabs, dc, (long long)ldo);
    if (n_rows == 0) return GMR_OK;
    GMR_REQUIRE(slabs != nullptr && out != nullptr && (uintptr_t)slabs % 16 == 0 && (uintptr_t)out % 16 == 0,
                "gmr_slabs_to_rows_f32: null or unaligned operand");
    const int64_t total = n_rows * n_slabs * (dc / 4);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)gmr::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    gmr::slabs_to_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(slabs), n_slabs, slab_stride / 4, n_rows, dc / 4, reinterpret_cast<float4*>(out), ldo / 4);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

extern "C" int gmr_rows_sumsq_f32(const float* z, int64_t ldz, int64_t n_rows, int32_t D, float* out, void* stream)
{
    GMR_REQUIRE(D == 4 || D == 8 || D == 16 || D == 32 || D == 64, "gmr_rows_sumsq_f32: D must be 4, 8, 16, 32 or 64 (got %d)", D);
    GMR_REQUIRE(n_rows >= 0 && ldz >= D && ldz % 4 == 0, "gmr_rows_sumsq_f32: bad shape");
    if (n_rows == 0) return GMR_OK;
    GMR_REQUIRE(z != nullptr && out != nullptr && (uintptr_t)z % 16 == 0, "gmr_rows_sumsq_f32: null or unaligned operand");
    const int lpr = D / 4;
    const unsigned grid = (unsigned)((n_rows * lpr + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    switch (lpr) {
        case 1: gmr::rows_sumsq_kernel<1><<<grid, 256, 0, st>>>(z, ldz, n_rows, out); break;
        case 2: gmr::rows_sumsq_kernel<2><<<grid, 256, 0, st>>>(z, ldz, n_rows, out); break;
        case 4: gmr::rows_sumsq_kernel<4><<<grid, 256, 0, st>>>(z, ldz, n_rows, out); break;
        case 8: gmr::rows_sumsq_kernel<8><<<grid, 256, 0, st>>>(z, ldz, n_rows, out); break;
        default: gmr::rows_sumsq_kernel<16><<<grid, 256, 0, st>>>(z, ldz, n_rows, out); break;
    }
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

extern "C" int gmr_rows_axpby_ss_f32(const float* x, int64_t ldx, const float* y, int64_t ldy, const float* z, int64_t ldz,
                                     const float* ss_parts, int32_t n_parts, int64_t ss_stride, float* out, int64_t ldo,
                                     int64_t n_rows, int32_t D, float a, float b, float c, float eps, void* stream)
{
    GMR_REQUIRE(x != nullptr && out != nullptr, "gmr_rows_axpby_ss_f32: null x/out");
    GMR_REQUIRE(D >= 4 && D % 4 == 0 && n_rows >= 0, "gmr_rows_axpby_ss_f32: D must be a multiple of 4 (got %d)", D);
    auto al16 = [](const float* p, int64_t ld) { return p == nullptr || ((uintptr_t)p % 16 == 0 && ld % 4 == 0); };
    GMR_REQUIRE(al16(x, ldx) && al16(y, ldy) && al16(z, ldz) && al16(out, ldo), "gmr_rows_axpby_ss_f32: rows must be 16-byte aligned");
    if (z != nullptr)
        GMR_REQUIRE(ss_parts != nullptr && n_parts >= 1 && n_parts <= gmr::kMaxPeers && (n_parts & (n_parts - 1)) == 0,
                    "gmr_rows_axpby_ss_f32: the number of norm parts must be a power of two <= %d (got %d)", gmr::kMaxPeers, n_parts);
    if (n_rows == 0) return GMR_OK;
    const int64_t total = n_rows * (D / 4);
    gmr::rows_axpby_ss_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, y, ldy, z, ldz, ss_parts, n_parts, ss_stride, out, ldo, n_rows, D / 4, a, b, c, eps);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

extern "C" int gmr_peer_barrier(uint32_t* const* flags_peers, int32_t my_rank, int32_t n_ranks, uint32_t* state, void* stream)
{
    GMR_REQUIRE(flags_peers != nullptr && state != nullptr, "gmr_peer_barrier: null argument");
    GMR_REQUIRE(n_ranks >= 1 && n_ranks <= 32 && my_rank >= 0 && my_rank < n_ranks, "gmr_peer_barrier: bad rank %d of %d", my_rank,
                n_ranks);
    gmr::peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags_peers, my_rank, n_ranks, state);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}
