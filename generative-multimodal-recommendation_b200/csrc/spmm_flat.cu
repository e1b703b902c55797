// spmm_flat.cu -- K1b: column-blocked, nonzero-centric SpMM  Y = alpha * A * X + beta * Y  (fp32, sm_100a)
//
// Same operator as spmm.cu (torch.sparse.mm at GenMMRec/src/models/diffmm.py:136-152,284-285 and siblings), built
// for graphs whose gathered operand X does not fit the L2 (the 1M x 500k shape: 256-384 MB against an effective
// ~60 MB of L2 per die).  The row-centric kernel of spmm.cu is then bound by DRAM gather misses (7x the algorithmic
// traffic, profiles/r01_ncu_full_step_final_summary.txt).  Here:
//
//   * the PLAN re-lays the matrix out once: columns are cut into blocks whose slice of X fits the L2, the nonzeros
//     are stably sorted by column block (cub radix sort on the device) and stored as three flat arrays
//     (row | col | val) in (block, row, original order) order.  A call runs one pass per block, back to back on the
//     stream, so every pass gathers from an L2-resident slice and Y is accumulated across passes.
//   * the KERNEL is nonzero-centric: a half-warp (D <= 64) or warp owns a TILE of ~64 consecutive nonzeros, stages
//     them global -> shared with 16-byte cp.async (double-buffered, next tile in flight), and runs one continuous
//     gather pipeline over the tile regardless of where rows begin and end; a row's sum is flushed when the row id
//     changes.  Short rows no longer cost a latency chain each, which is what made blocked passes lose with the
//     row-centric kernel (profiles/r01_spmm_block_probe.json).
//   * summation order is CANONICAL per row: the nonzeros of a (row, block) segment are cut into pieces of kT counted
//     from the segment start, a piece is summed sequentially, pieces of long segments go through fp32 slots that a
//     second kernel adds in (block, piece) order, and blocks are added in ascending order.  Tile boundaries only fall
//     on piece boundaries, so the result of a row depends on the row alone (not on its neighbours or on how rows are
//     sharded over GPUs) and is run-to-run deterministic.
//
// Bound: the L2 -> SM gather path for nnz * 4 D bytes.  Measured on B200 (profiles/r02_spmm_table_sweep.log): both this
// kernel and K1 gather 16.5 TB/s from a table of <= 48 MB and 8.9 TB/s from 256 MB (uniform graph); on the power-law
// 1M x 500k graph this kernel reaches 10.6 TB/s with 48 MB blocks -- the same as K1 without blocking, because the short
// per-block segments of the tail rows add a flush + Y read-modify-write every ~8 nonzeros.  It is therefore OPT-IN
// (GMR_SPMM_BLOCKED=1); what it adds over K1 is the shard-independent summation order and 4.5x less DRAM traffic
// (DESIGN.md section 3).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <climits>
#include <vector>

#include "common.cuh"

namespace gmr {

constexpr int kT = 64;          // piece length == tile granularity (nonzeros)
constexpr int kTileMax = 128;   // a tile holds fewer than 2 kT nonzeros ...
constexpr int kRunCap = 8;      // ... and at most this many runs
constexpr int kBatch = 8;       // nonzeros per inner-loop batch; tiles are stored padded to a multiple of it
constexpr int kStageInts = 2 * kTileMax + 16;  // staged per tile: cols | vals | bitmap[4] of run ends | run destinations [8] | pad
constexpr int kRecInts = 16;    // tile record: {offset, padded n, runs, n, bitmap[4], destination of each run [8]}
constexpr int kFlatThreads = 256;
constexpr int kFirstBit = (int)0x80000000u;  // run destination: Y row, first writer of that row in the call
constexpr int kSlotBit = 0x40000000;         // run destination: partial-sum slot (low 30 bits) instead of a Y row

}  // namespace gmr

// RUN = maximal stretch of one row inside a tile = one piece (see flat_boundary_kernel).  The kernel never looks at a
// per-nonzero row id: a tile record carries a 128-bit map of the positions that end a run and the destination of each
// of its (at most kRunCap) runs.  Tiles are stored 16-byte aligned and padded to a multiple of kBatch nonzeros by
// repeating their last entry (whatever the padding adds lands in the accumulator after the tile's last run was flushed).
struct gmr_spmm_bplan {
    int64_t n_rows = 0, n_cols = 0, nnz = 0, block_cols = 0;
    int32_t n_blocks = 0;
    int64_t n_tiles = 0, n_runs = 0, n_slots = 0, n_red_small = 0, n_red_big = 0, n_segments = 0, n_long_segments = 0;
    int64_t n_padded = 0;             // entries of d_col / d_val / d_perm
    std::vector<int64_t> blk_tile0;   // [n_blocks + 1] first tile of each block
    std::vector<int64_t> blk_ntiles;  // [n_blocks]
    // device arrays
    int32_t* d_col = nullptr;         // [n_padded] tile by tile
    float* d_val = nullptr;           // [n_padded]
    int32_t* d_perm = nullptr;        // [n_padded] padded position -> CSR position (value refresh)
    int32_t* d_tile_rec = nullptr;    // [n_tiles][kRecInts]
    int32_t* d_red_row = nullptr;     // [n_red] row | first-touch bit; small entries first, then big ones
    int32_t* d_red_lo = nullptr;      // [n_red] slots [lo, hi) of each entry
    int32_t* d_red_hi = nullptr;
};

namespace gmr {

// ---------------------------------------------------------------------------------------------------------------
// plan construction kernels.  Segment = the nonzeros of one row inside one column block (contiguous after the
// sort).  Segment ids are 1-BASED (they come out of an inclusive scan of the start flags); every per-segment
// array therefore has n_seg + 2 elements with element 0 unused and seg_pos[n_seg + 1] == nnz as a terminator.
// ---------------------------------------------------------------------------------------------------------------

constexpr int32_t kNoBlock = 0x7f7f7f7f;  // first_block[] value of a row without any short segment (memset 0x7f)

__global__ void flat_expand_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                                   int64_t nnz, int64_t block_cols, int32_t* __restrict__ row_of,
                                   unsigned char* __restrict__ key, int32_t* __restrict__ ident)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nnz) return;
    int64_t lo = 0, hi = n_rows;  // last row r with rowptr[r] <= j (rows may be empty)
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)rowptr[mid] <= j) lo = mid; else hi = mid;
    }
    row_of[j] = (int32_t)lo;
    key[j] = (unsigned char)((int64_t)col[j] / block_cols);
    ident[j] = (int32_t)j;
}

__global__ void flat_block_bounds_kernel(const unsigned char* __restrict__ skey, int64_t nnz, int32_t n_blocks,
                                         int32_t* __restrict__ cs)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;  // cs[b] = first position whose key is >= b
    if (b > n_blocks) return;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int)skey[mid] < b) lo = mid + 1; else hi = mid;
    }
    cs[b] = (int32_t)lo;
}

__global__ void flat_gather_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ row_of,
                                   const int32_t* __restrict__ col, const float* __restrict__ val,
                                   const unsigned char* __restrict__ skey, int64_t nnz, int32_t* __restrict__ e_row,
                                   int32_t* __restrict__ e_col, float* __restrict__ e_val, int32_t* __restrict__ flag)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int32_t j = perm[p];
    const int32_t r = row_of[j];
    e_row[p] = r;
    e_col[p] = col[j];
    e_val[p] = val[j];
    bool start = (p == 0);
    if (!start) start = (skey[p - 1] != skey[p]) || (row_of[perm[p - 1]] != r);
    flag[p] = start ? 1 : 0;
}

__global__ void flat_values_kernel(const int32_t* __restrict__ perm, const float* __restrict__ val, int64_t nnz,
                                   float* __restrict__ e_val)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) e_val[p] = val[perm[p]];
}

__global__ void flat_seg_pos_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ seg_id, int64_t nnz,
                                    int32_t n_seg, int32_t* __restrict__ seg_pos)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz && flag[p]) seg_pos[seg_id[p]] = (int32_t)p;
    if (p == 0) {
        seg_pos[0] = 0;
        seg_pos[n_seg + 1] = (int32_t)nnz;
    }
}

// per segment s in [1, n_seg]: slots it needs (0 = short: summed by one group and written straight to Y); short
// segments also bid for the first-touch of their row (the smallest block wins: passes run in block order)
__global__ void flat_seg_info_kernel(const int32_t* __restrict__ seg_pos, const int32_t* __restrict__ e_row,
                                     const unsigned char* __restrict__ skey, int32_t n_seg, int32_t* __restrict__ seg_np,
                                     int32_t* __restrict__ first_block)
{
    const int s = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 1) seg_np[0] = 0;
    if (s > n_seg) return;
    const int32_t p = seg_pos[s], len = seg_pos[s + 1] - p;
    if (len > kT) {
        seg_np[s] = (len + kT - 1) / kT;
    } else {
        seg_np[s] = 0;
        atomicMin(first_block + e_row[p], (int32_t)skey[p]);
    }
}

__global__ void flat_long_keys_kernel(const int32_t* __restrict__ long_ids, const int32_t* __restrict__ seg_pos,
                                      const int32_t* __restrict__ e_row, int32_t n_long, int32_t* __restrict__ long_row,
                                      int32_t* __restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_long) return;
    long_row[i] = e_row[seg_pos[long_ids[i]]];
    idx[i] = i;
}

__global__ void flat_long_np_kernel(const int32_t* __restrict__ sorted_idx, const int32_t* __restrict__ long_ids,
                                    const int32_t* __restrict__ seg_np, int32_t n_long, int32_t* __restrict__ np_sorted)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_long) np_sorted[i] = seg_np[long_ids[sorted_idx[i]]];
}

// after the stable sort of the long segments by row and the exclusive scan of their slot counts
__global__ void flat_long_scatter_kernel(const int32_t* __restrict__ sorted_idx, const int32_t* __restrict__ sorted_row,
                                         const int32_t* __restrict__ long_ids, const int32_t* __restrict__ slot_base_sorted,
                                         int32_t n_long, int32_t* __restrict__ seg_slot, int32_t* __restrict__ head)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_long) return;
    seg_slot[long_ids[sorted_idx[i]]] = slot_base_sorted[i];
    head[i] = (i == 0 || sorted_row[i] != sorted_row[i - 1]) ? 1 : 0;
}

// one reduce entry per row that owns long segments: its slots are contiguous [lo, hi)
__global__ void flat_red_entries_kernel(const int32_t* __restrict__ head, const int32_t* __restrict__ head_scan,
                                        const int32_t* __restrict__ sorted_row, const int32_t* __restrict__ slot_base_sorted,
                                        const int32_t* __restrict__ np_sorted, const int32_t* __restrict__ first_block,
                                        int32_t n_long, int32_t* __restrict__ red_row, int32_t* __restrict__ red_lo,
                                        int32_t* __restrict__ red_hi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_long) return;
    if (head[i]) {
        const int e = head_scan[i] - 1;
        const int32_t r = sorted_row[i];
        red_row[e] = r | (first_block[r] == kNoBlock ? kFirstBit : 0);
        red_lo[e] = slot_base_sorted[i];
    }
    if (i + 1 == n_long || head[i + 1]) red_hi[head_scan[i] - 1] = slot_base_sorted[i] + np_sorted[i];
}

// rows without any nonzero: Y = beta * Y through a reduce entry with zero slots
__global__ void flat_empty_rows_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows, int32_t* __restrict__ out_row,
                                       int32_t* __restrict__ out_lo, int32_t* __restrict__ out_hi,
                                       int32_t* __restrict__ counter)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows || rowptr[r + 1] != rowptr[r]) return;
    const int k = atomicAdd(counter, 1);
    if (out_row != nullptr) {
        out_row[k] = (int32_t)r | kFirstBit;
        out_lo[k] = 0;
        out_hi[k] = 0;
    }
}

// run starts: segment starts and the starts of the pieces m >= 1 of long segments
__global__ void flat_run_flag_kernel(const int32_t* __restrict__ seg_id, const int32_t* __restrict__ seg_pos, int64_t nnz,
                                     int32_t* __restrict__ rflag)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    rflag[p] = (((int32_t)p - seg_pos[seg_id[p]]) % kT == 0) ? 1 : 0;
}

// destination of every run: the slot of a piece of a long segment, else the row of a whole short segment with the
// first-touch bit on the segment that is the first writer of its row
__global__ void flat_run_rows_kernel(const int32_t* __restrict__ rflag, const int32_t* __restrict__ run_id,
                                     const int32_t* __restrict__ seg_id, const int32_t* __restrict__ seg_pos,
                                     const int32_t* __restrict__ seg_np, const int32_t* __restrict__ seg_slot,
                                     const unsigned char* __restrict__ skey, const int32_t* __restrict__ first_block,
                                     const int32_t* __restrict__ e_row, int64_t nnz, int32_t* __restrict__ run_row)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz || !rflag[p]) return;
    const int32_t sid = seg_id[p];
    int32_t out;
    if (seg_np[sid] > 0) {
        out = kSlotBit | (seg_slot[sid] + ((int32_t)p - seg_pos[sid]) / kT);
    } else {
        const int32_t r = e_row[p];
        out = (first_block[r] == (int32_t)skey[p]) ? (r | kFirstBit) : r;
    }
    run_row[run_id[p] - 1] = out;
}

// tile boundaries.  Piece boundaries are segment start + m * kT.  A position is a tile boundary when it is
//   (1) the last piece boundary at or before a probe position (block start + k * kT), or
//   (2) the start of a piece m >= 1 of a long segment.
//   (3) the start of every kRunCap-th run (global run numbering).
// Probes are kT apart, so a tile holds fewer than 2 kT nonzeros; by (2) a run of equal row ids inside a tile is
// exactly one piece (a piece of a long segment, or a whole short segment); by (3) a tile holds at most kRunCap runs.
// Block starts are boundaries (k = 0), so tiles never straddle blocks.
__global__ void flat_boundary_kernel(const int32_t* __restrict__ seg_id, const int32_t* __restrict__ seg_pos,
                                     const int32_t* __restrict__ rflag, const int32_t* __restrict__ run_id,
                                     const unsigned char* __restrict__ skey, const int32_t* __restrict__ cs, int64_t nnz,
                                     unsigned char* __restrict__ bflag)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int32_t s = seg_pos[seg_id[p]];
    const int32_t d = (int32_t)p - s;
    if (d > 0 && d % kT == 0) bflag[p] = 1;
    if (rflag[p] && (run_id[p] - 1) % kRunCap == 0) bflag[p] = 1;
    if (((int32_t)p - cs[skey[p]]) % kT == 0) bflag[s + (d / kT) * kT] = 1;
}

__global__ void flat_block_tiles_kernel(const int32_t* __restrict__ tile_start, int32_t n_tiles, const int32_t* __restrict__ cs,
                                        int32_t n_blocks, int32_t* __restrict__ tile0)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;  // tile0[b] = first tile starting at or after cs[b]
    if (b > n_blocks) return;
    const int32_t target = cs[b];
    int lo = 0, hi = n_tiles;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tile_start[mid] < target) lo = mid + 1; else hi = mid;
    }
    tile0[b] = lo;
}

// padded length of every tile (input of the exclusive scan that lays the tiles out)
__global__ void flat_tile_len_kernel(const int32_t* __restrict__ tile_start, int32_t n_tiles, int32_t* __restrict__ len)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_tiles) len[g] = (tile_start[g + 1] - tile_start[g] + kBatch - 1) / kBatch * kBatch;
}

// the record of every tile
__global__ void flat_tile_rec_kernel(const int32_t* __restrict__ tile_start, const int32_t* __restrict__ tile_off, int32_t n_tiles,
                                     const int32_t* __restrict__ rflag, const int32_t* __restrict__ run_id,
                                     const int32_t* __restrict__ run_row, int32_t* __restrict__ tile_rec,
                                     int32_t* __restrict__ bad)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_tiles) return;
    const int32_t s = tile_start[g], e = tile_start[g + 1];
    const int n = e - s;
    const int r0 = run_id[s] - 1, nr = run_id[e - 1] - run_id[s] + 1;
    if (n < 1 || n >= kTileMax || nr < 1 || nr > kRunCap || !rflag[s]) atomicAdd(bad, 1);  // construction invariants
    uint32_t bits[4] = {0u, 0u, 0u, 0u};
    for (int j = 0; j < n; ++j)
        if (j + 1 == n || rflag[s + j + 1]) bits[j >> 5] |= 1u << (j & 31);
    int4* out = reinterpret_cast<int4*>(tile_rec + (int64_t)g * kRecInts);
    out[0] = make_int4(tile_off[g], (n + kBatch - 1) / kBatch * kBatch, nr, n);
    out[1] = make_int4((int)bits[0], (int)bits[1], (int)bits[2], (int)bits[3]);
    int dst[kRunCap];
#pragma unroll
    for (int k = 0; k < kRunCap; ++k) dst[k] = (k < nr) ? run_row[r0 + k] : 0;
    out[2] = make_int4(dst[0], dst[1], dst[2], dst[3]);
    out[3] = make_int4(dst[4], dst[5], dst[6], dst[7]);
}

// blocked order -> padded tile layout; one group of 16 threads per tile
__global__ void flat_tile_layout_kernel(const int32_t* __restrict__ tile_start, const int32_t* __restrict__ tile_off, int32_t n_tiles,
                                        const int32_t* __restrict__ b_col, const float* __restrict__ b_val,
                                        const int32_t* __restrict__ b_perm, int32_t* __restrict__ p_col,
                                        float* __restrict__ p_val, int32_t* __restrict__ p_perm)
{
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int sub = threadIdx.x & 15;
    if (g >= n_tiles) return;
    const int32_t s = tile_start[g], n = tile_start[g + 1] - s, off = tile_off[g];
    const int n8 = (n + kBatch - 1) / kBatch * kBatch;
    for (int k = sub; k < n8; k += 16) {
        const int src = s + min(k, n - 1);
        p_col[off + k] = b_col[src];
        p_val[off + k] = b_val[src];
        p_perm[off + k] = b_perm[src];
    }
}

constexpr int kBigSlots = 64;  // reduce entries with more slots than this take the CTA-per-entry kernel

__global__ void flat_red_class_kernel(const int32_t* __restrict__ lo, const int32_t* __restrict__ hi, int32_t n,
                                      int32_t* __restrict__ is_small)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) is_small[i] = (hi[i] - lo[i] <= kBigSlots) ? 1 : 0;
}

// stable partition of the reduce entries into small | big
__global__ void flat_red_partition_kernel(const int32_t* __restrict__ red_row, const int32_t* __restrict__ lo,
                                          const int32_t* __restrict__ hi, const int32_t* __restrict__ is_small,
                                          const int32_t* __restrict__ small_scan, int32_t n, int32_t n_small,
                                          int32_t* __restrict__ out_row, int32_t* __restrict__ out_lo,
                                          int32_t* __restrict__ out_hi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = is_small[i] ? (small_scan[i] - 1) : (n_small + (i - small_scan[i]));
    out_row[k] = red_row[i];
    out_lo[k] = lo[i];
    out_hi[k] = hi[i];
}

__device__ __forceinline__ void fma4(float4& a, float w, const float4& x)
{
    a.x = fmaf(w, x.x, a.x);
    a.y = fmaf(w, x.y, a.y);
    a.z = fmaf(w, x.z, a.z);
    a.w = fmaf(w, x.w, a.w);
}

// One GROUP of LANES lanes (half-warp for row widths <= 64 floats, warp otherwise) per tile; lanes own one float4 column
// each (blockIdx.y selects the column chunk of wider rows).  The two half-warps of a warp run their own trip counts and
// meet at the __syncwarp()s.
//
// Per tile the group stages with 16-byte cp.async (double-buffered, the next tile in flight while this one is multiplied)
// its padded column ids and values, the 128-bit map of run ends and the run destinations; and, into a ring of kRunCap
// rows, the OLD Y row of every run that accumulates into Y (column blocks after a row's first one), issued as soon as the
// previous tile has flushed its last run: the read-modify-write of Y never waits for memory.  Tile records (offset,
// sizes, run destinations) are plain loads issued two tiles ahead.  The inner loop handles batches of eight nonzeros:
// two LDS.128 of column ids, eight 128-bit gathers, two LDS.128 of values and 32 FFMAs when the batch's eight map bits
// are clear; a per-nonzero path with the run flush otherwise.  Latency is hidden by the 24 resident warps per SM, not by
// software pipelining inside a warp (two register sets cost a third of the occupancy and twice the code).
template <int LANES>
__global__ void __launch_bounds__(kFlatThreads, 3)
    spmm_flat_kernel(const int32_t* __restrict__ p_col, const float* __restrict__ p_val, const int32_t* __restrict__ tile_rec,
                     int32_t n_tiles, const float* __restrict__ X, int64_t ldx, float* Y, int64_t ldy,
                     float* __restrict__ partial, int32_t DV, float alpha, float beta)
{
    constexpr int GROUPS = kFlatThreads / LANES;
    extern __shared__ int4 flat_smem4[];
    const int lane = threadIdx.x & 31;
    const int sub = threadIdx.x % LANES, gl = threadIdx.x / LANES;
    const int half = (LANES == 16) ? (lane >> 4) : 0;
    int* stage = reinterpret_cast<int*>(flat_smem4) + gl * (2 * kStageInts);
    float4* ring = reinterpret_cast<float4*>(reinterpret_cast<int*>(flat_smem4) + GROUPS * 2 * kStageInts) +
                   gl * (kRunCap * LANES) + sub;
    const int vcol = blockIdx.y * LANES + sub;
    const bool colok = vcol < DV;
    const int vload = colok ? vcol : DV - 1;             // lanes past the row width gather a valid column and never store
    const uint32_t ldxb = (uint32_t)(ldx * 4);           // row pitches in bytes (< 4 GB): one IMAD.WIDE.U32 per address
    const uint32_t ldyb = (uint32_t)(ldy * 4), ldpb = (uint32_t)DV * 16u;
    const char* Xl = reinterpret_cast<const char*>(reinterpret_cast<const float4*>(X) + vload);
    char* Yl = reinterpret_cast<char*>(reinterpret_cast<float4*>(Y) + vload);
    char* Pl = reinterpret_cast<char*>(reinterpret_cast<float4*>(partial) + vload);
    const uint64_t pol_stream = l2_policy_evict_first();  // CSR stream and Y rows leave the L2 first; gathers use the default
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool read_first = beta != 0.f;

    const int64_t G = (int64_t)gridDim.x * GROUPS;
    int64_t t = (int64_t)blockIdx.x * GROUPS + gl;

    // (offset, padded n, runs) of a tile and, in lane e < kRunCap of the group, the destination of its run e
    auto load_rec = [&](int64_t tt, int& off, int& n8, int& nr, int& dst) {
        off = 0; n8 = 0; nr = 0; dst = 0;
        if (tt < n_tiles) {
            const int4 h = __ldg(reinterpret_cast<const int4*>(tile_rec + tt * kRecInts));
            off = h.x; n8 = h.y; nr = h.z;
            dst = __ldg(tile_rec + tt * kRecInts + 8 + (sub & (kRunCap - 1)));
        }
    };
    auto cp16 = [&](const void* dst, const void* src) {
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                     "l"(src), "l"(pol_stream) : "memory");
    };
    // cp.async group 1 of a tile: column ids, values, run-end map, run destinations
    auto issue_stage = [&](int buf, int64_t tt, int off, int n8) {
        if (n8 > 0) {
            int* sdst = stage + buf * kStageInts;
            for (int k = 4 * sub; k < n8; k += 4 * LANES) {
                cp16(sdst + k, p_col + off + k);
                cp16(sdst + kTileMax + k, p_val + off + k);
            }
            if (sub < 3) cp16(sdst + 2 * kTileMax + 4 * sub, tile_rec + tt * kRecInts + 4 + 4 * sub);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // cp.async group 2 of a tile: the old Y rows of its runs that accumulate.  Each lane copies and later reads back its own
    // 16 bytes, so wait_group alone orders the two.  Called where the warp is converged (full-mask shuffles).
    auto issue_rows = [&](int nr, int dst) {
        int nrmax = nr;
        if (LANES == 16) nrmax = max(nrmax, __shfl_xor_sync(0xffffffffu, nrmax, 16));
        for (int e = 0; e < nrmax; ++e) {
            const int d = __shfl_sync(0xffffffffu, dst, e, LANES);
            if (e < nr && colok && !(d & kSlotBit) && (d >= 0 || read_first))
                cp16(ring + e * LANES, Yl + (uint64_t)(uint32_t)(d & 0x3fffffff) * ldyb);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (t - half >= n_tiles) return;  // whole warp leaves together
    int o0, n0, nr0, d0, o1, n1, nr1, d1;
    load_rec(t, o0, n0, nr0, d0);
    issue_stage(0, t, o0, n0);
    issue_rows(nr0, d0);
    load_rec(t + G, o1, n1, nr1, d1);
    int buf = 0;
    while (true) {
        issue_stage(buf ^ 1, t + G, o1, n1);
        int o2, n2, nr2, d2;
        load_rec(t + 2 * G, o2, n2, nr2, d2);
        asm volatile("cp.async.wait_group 2;" ::: "memory");  // this tile's stage; its Y rows and the next stage may be in flight
        __syncwarp();

        const int* sc = stage + buf * kStageInts;
        const float* sv = reinterpret_cast<const float*>(sc + kTileMax);
        const int* sbits = sc + 2 * kTileMax;
        const int* sdst = sbits + 4;
        float4 acc = zero;
        int k = 0;  // current run of the tile
        bool rows_landed = false;

        auto flush = [&]() {
            if (!rows_landed) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");  // the Y rows of this tile (the next stage may be in flight)
                rows_landed = true;
            }
            const int d = sdst[k];
            const uint32_t idx = (uint32_t)(d & 0x3fffffff);
            if (d & kSlotBit) {
                if (colok) *reinterpret_cast<float4*>(Pl + (uint64_t)idx * ldpb) = acc;
            } else {
                float4 o;
                if (d >= 0) {
                    const float4 y = ring[k * LANES];
                    o = make_float4(fmaf(alpha, acc.x, y.x), fmaf(alpha, acc.y, y.y), fmaf(alpha, acc.z, y.z), fmaf(alpha, acc.w, y.w));
                } else {  // first writer of this row in this call
                    o = make_float4(alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w);
                    if (read_first) {
                        const float4 y = ring[k * LANES];
                        o.x = fmaf(beta, y.x, o.x); o.y = fmaf(beta, y.y, o.y);
                        o.z = fmaf(beta, y.z, o.z); o.w = fmaf(beta, y.w, o.w);
                    }
                }
                if (colok) __stcs(reinterpret_cast<float4*>(Yl + (uint64_t)idx * ldyb), o);
            }
            acc = zero;
            ++k;
        };

        // One batch of eight gathers per lane at a time; latency is covered by the 24 resident warps per SM.  Measured
        // alternative (two batches in flight per lane, 122 registers, 16 warps per SM): same time within 3 %
        // (profiles/r02_spmm_flat_probe_v6.json vs _v7.json) -- the kernel follows the L2 -> SM gather rate, see DESIGN.md.
        for (int j = 0; j < n0; j += kBatch) {
            const int4 ca = *reinterpret_cast<const int4*>(sc + j), cb = *reinterpret_cast<const int4*>(sc + j + 4);
            const int cc[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
            float4 x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                x[q] = __ldg(reinterpret_cast<const float4*>(Xl + (uint64_t)(uint32_t)cc[q] * ldxb));
            const float4 wa = *reinterpret_cast<const float4*>(sv + j), wb = *reinterpret_cast<const float4*>(sv + j + 4);
            const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            const uint32_t m = ((uint32_t)sbits[j >> 5] >> (j & 31)) & 255u;
            if (m == 0u) {
#pragma unroll
                for (int q = 0; q < 8; ++q) fma4(acc, ww[q], x[q]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    fma4(acc, ww[q], x[q]);
                    if ((m >> q) & 1u) flush();
                }
            }
        }
        __syncwarp();  // everyone is done with this tile's buffers: the tile after next may land in the stage, the next
                       // tile's Y rows in the ring
        issue_rows(nr1, d1);

        t += G;
        if (t - half >= n_tiles) break;
        o0 = o1; n0 = n1; nr0 = nr1; d0 = d1;
        o1 = o2; n1 = n2; nr1 = nr2; d1 = d2;
        buf ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Reduce entries with few slots: one half-warp (16 float4 columns at a time) per entry, slots added in order.
__global__ void __launch_bounds__(256)
    flat_reduce_small_kernel(const int32_t* __restrict__ red_row, const int32_t* __restrict__ red_lo,
                             const int32_t* __restrict__ red_hi, int32_t n_entries, const float* __restrict__ partial,
                             float* Y, int64_t ldy, int32_t DV, float alpha, float beta)
{
    const int sub = threadIdx.x & 15;
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (e >= n_entries) return;
    const int rraw = red_row[e];
    const int row = rraw & 0x3fffffff;
    const int first = red_lo[e], n = red_hi[e] - first;
    const float4* base = reinterpret_cast<const float4*>(partial) + (int64_t)first * DV;
    for (int vc = sub; vc < DV; vc += 16) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        int k = 0;
        for (; k + 4 <= n; k += 4) {
            const float4 p0 = base[(int64_t)k * DV + vc], p1 = base[(int64_t)(k + 1) * DV + vc];
            const float4 p2 = base[(int64_t)(k + 2) * DV + vc], p3 = base[(int64_t)(k + 3) * DV + vc];
            a.x = (((a.x + p0.x) + p1.x) + p2.x) + p3.x;
            a.y = (((a.y + p0.y) + p1.y) + p2.y) + p3.y;
            a.z = (((a.z + p0.z) + p1.z) + p2.z) + p3.z;
            a.w = (((a.w + p0.w) + p1.w) + p2.w) + p3.w;
        }
        for (; k < n; ++k) {
            const float4 p = base[(int64_t)k * DV + vc];
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        float4* yp = reinterpret_cast<float4*>(Y + (int64_t)row * ldy) + vc;
        float4 o = make_float4(alpha * a.x, alpha * a.y, alpha * a.z, alpha * a.w);
        if (rraw < 0) {
            if (beta != 0.f) {
                const float4 y = *yp;
                o.x = fmaf(beta, y.x, o.x); o.y = fmaf(beta, y.y, o.y); o.z = fmaf(beta, y.z, o.z); o.w = fmaf(beta, y.w, o.w);
            }
        } else {
            const float4 y = *yp;
            o.x += y.x; o.y += y.y; o.z += y.z; o.w += y.w;
        }
        *yp = o;
    }
}

// Reduce entries with many slots (popular rows): one CTA of 512 threads per entry and chunk of L float4 columns
// (L = 16 for rows of <= 64 floats, else 32).  The 512 / L groups each add the slots of a fixed contiguous sub-range in
// order (loads batched eight deep, adds sequential), then the group sums are added in group order.  The split depends
// on the slot count alone, so the result is deterministic and independent of the other rows.
template <int L>
__global__ void __launch_bounds__(512)
    flat_reduce_big_kernel(const int32_t* __restrict__ red_row, const int32_t* __restrict__ red_lo,
                           const int32_t* __restrict__ red_hi, const float* __restrict__ partial, float* Y, int64_t ldy,
                           int32_t DV, float alpha, float beta)
{
    constexpr int NG = 512 / L;
    __shared__ float4 red[NG][L];
    const int e = blockIdx.x;
    const int rraw = red_row[e];
    const int row = rraw & 0x3fffffff;
    const int first = red_lo[e], n = red_hi[e] - first;
    const int g = threadIdx.x / L, sub = threadIdx.x % L;
    const int vc = blockIdx.y * L + sub;
    const int per = (n + NG - 1) / NG;
    const int lo = min(n, g * per), hi = min(n, lo + per);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vc < DV) {
        const float4* base = reinterpret_cast<const float4*>(partial) + (int64_t)first * DV + vc;
        int k = lo;
        for (; k + 8 <= hi; k += 8) {
            float4 p[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) p[q] = base[(int64_t)(k + q) * DV];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                a.x += p[q].x; a.y += p[q].y; a.z += p[q].z; a.w += p[q].w;
            }
        }
        for (; k < hi; ++k) {
            const float4 p = base[(int64_t)k * DV];
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
    }
    red[g][sub] = a;
    __syncthreads();
    if (g == 0 && vc < DV) {
#pragma unroll 4
        for (int w = 1; w < NG; ++w) {
            const float4 p = red[w][sub];
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        float4* yp = reinterpret_cast<float4*>(Y + (int64_t)row * ldy) + vc;
        float4 o = make_float4(alpha * a.x, alpha * a.y, alpha * a.z, alpha * a.w);
        if (rraw < 0) {
            if (beta != 0.f) {
                const float4 y = *yp;
                o.x = fmaf(beta, y.x, o.x); o.y = fmaf(beta, y.y, o.y); o.z = fmaf(beta, y.z, o.z); o.w = fmaf(beta, y.w, o.w);
            }
        } else {
            const float4 y = *yp;
            o.x += y.x; o.y += y.y; o.z += y.z; o.w += y.w;
        }
        *yp = o;
    }
}

template <typename T>
static cudaError_t dmalloc(T** p, int64_t n)
{
    return cudaMalloc((void**)p, sizeof(T) * (size_t)std::max<int64_t>(n, 1));
}

struct TempPool {  // frees every scratch allocation when plan construction leaves scope (success or error)
    std::vector<void*> ptrs;
    template <typename T>
    cudaError_t get(T** p, int64_t n)
    {
        cudaError_t e = dmalloc(p, n);
        if (e == cudaSuccess) ptrs.push_back((void*)*p);
        return e;
    }
    ~TempPool()
    {
        for (void* p : ptrs) cudaFree(p);
    }
};

static inline unsigned nblk(int64_t n, int threads = 256) { return (unsigned)((std::max<int64_t>(n, 1) + threads - 1) / threads); }

static cudaError_t inclusive_sum(TempPool& tmp, const int32_t* in, int32_t* out, int n, cudaStream_t st)
{
    size_t bytes = 0;
    cudaError_t e = cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, n, st);
    if (e != cudaSuccess) return e;
    unsigned char* t = nullptr;
    if ((e = tmp.get(&t, (int64_t)bytes)) != cudaSuccess) return e;
    return cub::DeviceScan::InclusiveSum(t, bytes, in, out, n, st);
}

static cudaError_t exclusive_sum(TempPool& tmp, const int32_t* in, int32_t* out, int n, cudaStream_t st)
{
    size_t bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, st);
    if (e != cudaSuccess) return e;
    unsigned char* t = nullptr;
    if ((e = tmp.get(&t, (int64_t)bytes)) != cudaSuccess) return e;
    return cub::DeviceScan::ExclusiveSum(t, bytes, in, out, n, st);
}

template <typename K>
static cudaError_t sort_pairs(TempPool& tmp, const K* kin, K* kout, const int32_t* vin, int32_t* vout, int n, int end_bit,
                              cudaStream_t st)
{
    size_t bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit, st);
    if (e != cudaSuccess) return e;
    unsigned char* t = nullptr;
    if ((e = tmp.get(&t, (int64_t)bytes)) != cudaSuccess) return e;
    return cub::DeviceRadixSort::SortPairs(t, bytes, kin, kout, vin, vout, n, 0, end_bit, st);
}

static int build_plan(gmr_spmm_bplan* p, const int32_t* rowptr, const int32_t* col, const float* val, cudaStream_t st)
{
    const int64_t n_rows = p->n_rows, nnz = p->nnz, n_blocks = p->n_blocks;
    TempPool tmp;
    std::vector<int32_t> cs(n_blocks + 1, 0);
    int32_t n_seg = 0, n_long = 0, n_slots = 0, n_rows_red = 0, n_empty = 0;
    int32_t *seg_id = nullptr, *seg_pos = nullptr, *seg_np = nullptr, *seg_slot = nullptr, *first_block = nullptr;
    int32_t *red_row = nullptr, *red_lo = nullptr, *red_hi = nullptr, *d_count = nullptr, *d_cs = nullptr;
    unsigned char* skey = nullptr;
    int32_t *e_row = nullptr, *rflag = nullptr, *run_id = nullptr, *run_row = nullptr, *b_col = nullptr, *b_perm = nullptr;
    float* b_val = nullptr;
    int32_t n_runs = 0;
    const int32_t nnz32 = (int32_t)nnz;
    GMR_CHECK_CUDA(tmp.get(&d_count, 4));
    GMR_CHECK_CUDA(cudaMemsetAsync(d_count, 0, 4 * sizeof(int32_t), st));

    // rows without nonzeros (their Y row still has to become beta * Y)
    flat_empty_rows_kernel<<<nblk(n_rows), 256, 0, st>>>(rowptr, n_rows, nullptr, nullptr, nullptr, d_count);
    GMR_LAUNCH_CHECK();
    GMR_CHECK_CUDA(cudaMemcpyAsync(&n_empty, d_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GMR_CHECK_CUDA(cudaStreamSynchronize(st));

    if (nnz > 0) {
        GMR_CHECK_CUDA(tmp.get(&b_col, nnz));
        GMR_CHECK_CUDA(tmp.get(&b_val, nnz));
        GMR_CHECK_CUDA(tmp.get(&b_perm, nnz));
        GMR_CHECK_CUDA(tmp.get(&e_row, nnz));

        int32_t *row_of = nullptr, *ident = nullptr, *flag = nullptr;
        unsigned char* key = nullptr;
        GMR_CHECK_CUDA(tmp.get(&row_of, nnz));
        GMR_CHECK_CUDA(tmp.get(&ident, nnz));
        GMR_CHECK_CUDA(tmp.get(&key, nnz));
        GMR_CHECK_CUDA(tmp.get(&skey, nnz));
        GMR_CHECK_CUDA(tmp.get(&flag, nnz));
        GMR_CHECK_CUDA(tmp.get(&d_cs, n_blocks + 1));
        GMR_CHECK_CUDA(tmp.get(&first_block, n_rows));
        GMR_CHECK_CUDA(cudaMemsetAsync(first_block, 0x7f, sizeof(int32_t) * (size_t)n_rows, st));  // kNoBlock
        flat_expand_kernel<<<nblk(nnz), 256, 0, st>>>(rowptr, col, n_rows, nnz, p->block_cols, row_of, key, ident);
        GMR_LAUNCH_CHECK();
        int end_bit = 1;
        while ((1 << end_bit) < n_blocks) ++end_bit;
        GMR_CHECK_CUDA(sort_pairs(tmp, key, skey, ident, b_perm, (int)nnz, end_bit, st));  // stable: CSR order inside a block
        flat_block_bounds_kernel<<<nblk(n_blocks + 1), 256, 0, st>>>(skey, nnz, (int32_t)n_blocks, d_cs);
        GMR_LAUNCH_CHECK();
        flat_gather_kernel<<<nblk(nnz), 256, 0, st>>>(b_perm, row_of, col, val, skey, nnz, e_row, b_col, b_val, flag);
        GMR_LAUNCH_CHECK();
        GMR_CHECK_CUDA(tmp.get(&seg_id, nnz));
        GMR_CHECK_CUDA(inclusive_sum(tmp, flag, seg_id, (int)nnz, st));  // 1-based segment ids
        GMR_CHECK_CUDA(cudaMemcpyAsync(&n_seg, seg_id + nnz - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        GMR_CHECK_CUDA(cudaMemcpyAsync(cs.data(), d_cs, sizeof(int32_t) * (n_blocks + 1), cudaMemcpyDeviceToHost, st));
        GMR_CHECK_CUDA(cudaStreamSynchronize(st));

        GMR_CHECK_CUDA(tmp.get(&seg_pos, (int64_t)n_seg + 2));
        GMR_CHECK_CUDA(tmp.get(&seg_np, (int64_t)n_seg + 2));
        GMR_CHECK_CUDA(tmp.get(&seg_slot, (int64_t)n_seg + 2));
        GMR_CHECK_CUDA(cudaMemsetAsync(seg_slot, 0xff, sizeof(int32_t) * ((size_t)n_seg + 2), st));
        flat_seg_pos_kernel<<<nblk(nnz), 256, 0, st>>>(flag, seg_id, nnz, n_seg, seg_pos);
        GMR_LAUNCH_CHECK();
        flat_seg_info_kernel<<<nblk(n_seg), 256, 0, st>>>(seg_pos, e_row, skey, n_seg, seg_np, first_block);
        GMR_LAUNCH_CHECK();

        // long segments -> slots, grouped by row (stable sort by row keeps the blocks ascending inside a row)
        int32_t* long_ids = nullptr;
        GMR_CHECK_CUDA(tmp.get(&long_ids, n_seg));
        {
            thrust::counting_iterator<int32_t> ids(1);
            size_t bytes = 0;
            GMR_CHECK_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, ids, seg_np + 1, long_ids, d_count + 1, n_seg, st));
            unsigned char* t = nullptr;
            GMR_CHECK_CUDA(tmp.get(&t, (int64_t)bytes));
            GMR_CHECK_CUDA(cub::DeviceSelect::Flagged(t, bytes, ids, seg_np + 1, long_ids, d_count + 1, n_seg, st));
        }
        GMR_CHECK_CUDA(cudaMemcpyAsync(&n_long, d_count + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        GMR_CHECK_CUDA(cudaStreamSynchronize(st));
        int32_t *s_row = nullptr, *s_base = nullptr, *s_np = nullptr, *head = nullptr, *head_scan = nullptr;
        if (n_long > 0) {
            int32_t *l_row = nullptr, *l_idx = nullptr, *s_idx = nullptr;
            GMR_CHECK_CUDA(tmp.get(&l_row, n_long));
            GMR_CHECK_CUDA(tmp.get(&l_idx, n_long));
            GMR_CHECK_CUDA(tmp.get(&s_row, n_long));
            GMR_CHECK_CUDA(tmp.get(&s_idx, n_long));
            GMR_CHECK_CUDA(tmp.get(&s_np, n_long));
            GMR_CHECK_CUDA(tmp.get(&s_base, n_long));
            GMR_CHECK_CUDA(tmp.get(&head, n_long));
            GMR_CHECK_CUDA(tmp.get(&head_scan, n_long));
            flat_long_keys_kernel<<<nblk(n_long), 256, 0, st>>>(long_ids, seg_pos, e_row, n_long, l_row, l_idx);
            GMR_LAUNCH_CHECK();
            GMR_CHECK_CUDA(sort_pairs(tmp, l_row, s_row, l_idx, s_idx, n_long, 32, st));
            flat_long_np_kernel<<<nblk(n_long), 256, 0, st>>>(s_idx, long_ids, seg_np, n_long, s_np);
            GMR_LAUNCH_CHECK();
            GMR_CHECK_CUDA(exclusive_sum(tmp, s_np, s_base, n_long, st));
            flat_long_scatter_kernel<<<nblk(n_long), 256, 0, st>>>(s_idx, s_row, long_ids, s_base, n_long, seg_slot, head);
            GMR_LAUNCH_CHECK();
            GMR_CHECK_CUDA(inclusive_sum(tmp, head, head_scan, n_long, st));
            int32_t last_base = 0, last_np = 0;
            GMR_CHECK_CUDA(cudaMemcpyAsync(&n_rows_red, head_scan + n_long - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            GMR_CHECK_CUDA(cudaMemcpyAsync(&last_base, s_base + n_long - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            GMR_CHECK_CUDA(cudaMemcpyAsync(&last_np, s_np + n_long - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            GMR_CHECK_CUDA(cudaStreamSynchronize(st));
            n_slots = last_base + last_np;
        }
        // runs and their rows
        GMR_CHECK_CUDA(tmp.get(&rflag, nnz));
        GMR_CHECK_CUDA(tmp.get(&run_id, nnz));
        flat_run_flag_kernel<<<nblk(nnz), 256, 0, st>>>(seg_id, seg_pos, nnz, rflag);
        GMR_LAUNCH_CHECK();
        GMR_CHECK_CUDA(inclusive_sum(tmp, rflag, run_id, (int)nnz, st));  // 1-based run ids
        GMR_CHECK_CUDA(cudaMemcpyAsync(&n_runs, run_id + nnz - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        GMR_CHECK_CUDA(cudaStreamSynchronize(st));
        GMR_CHECK_CUDA(tmp.get(&run_row, n_runs));
        flat_run_rows_kernel<<<nblk(nnz), 256, 0, st>>>(rflag, run_id, seg_id, seg_pos, seg_np, seg_slot, skey, first_block,
                                                        e_row, nnz, run_row);
        GMR_LAUNCH_CHECK();

        // reduce entries: rows with long segments, then rows without nonzeros
        const int32_t n_red = n_rows_red + n_empty;
        if (n_red > 0) {
            GMR_CHECK_CUDA(tmp.get(&red_row, n_red));
            GMR_CHECK_CUDA(tmp.get(&red_lo, n_red));
            GMR_CHECK_CUDA(tmp.get(&red_hi, n_red));
            if (n_long > 0) {
                flat_red_entries_kernel<<<nblk(n_long), 256, 0, st>>>(head, head_scan, s_row, s_base, s_np, first_block,
                                                                      n_long, red_row, red_lo, red_hi);
                GMR_LAUNCH_CHECK();
            }
            if (n_empty > 0) {
                GMR_CHECK_CUDA(cudaMemsetAsync(d_count + 2, 0, sizeof(int32_t), st));
                flat_empty_rows_kernel<<<nblk(n_rows), 256, 0, st>>>(rowptr, n_rows, red_row + n_rows_red, red_lo + n_rows_red,
                                                                    red_hi + n_rows_red, d_count + 2);
                GMR_LAUNCH_CHECK();
            }
            int32_t *is_small = nullptr, *small_scan = nullptr;
            GMR_CHECK_CUDA(tmp.get(&is_small, n_red));
            GMR_CHECK_CUDA(tmp.get(&small_scan, n_red));
            flat_red_class_kernel<<<nblk(n_red), 256, 0, st>>>(red_lo, red_hi, n_red, is_small);
            GMR_LAUNCH_CHECK();
            GMR_CHECK_CUDA(inclusive_sum(tmp, is_small, small_scan, n_red, st));
            int32_t n_small = 0;
            GMR_CHECK_CUDA(cudaMemcpyAsync(&n_small, small_scan + n_red - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            GMR_CHECK_CUDA(cudaStreamSynchronize(st));
            p->n_red_small = n_small;
            p->n_red_big = n_red - n_small;
            GMR_CHECK_CUDA(dmalloc(&p->d_red_row, n_red));
            GMR_CHECK_CUDA(dmalloc(&p->d_red_lo, n_red));
            GMR_CHECK_CUDA(dmalloc(&p->d_red_hi, n_red));
            flat_red_partition_kernel<<<nblk(n_red), 256, 0, st>>>(red_row, red_lo, red_hi, is_small, small_scan, n_red, n_small,
                                                                   p->d_red_row, p->d_red_lo, p->d_red_hi);
            GMR_LAUNCH_CHECK();
        }
    } else if (n_empty > 0) {  // no nonzeros at all: every row is an empty-row entry
        GMR_CHECK_CUDA(dmalloc(&p->d_red_row, n_empty));
        GMR_CHECK_CUDA(dmalloc(&p->d_red_lo, n_empty));
        GMR_CHECK_CUDA(dmalloc(&p->d_red_hi, n_empty));
        GMR_CHECK_CUDA(cudaMemsetAsync(d_count + 2, 0, sizeof(int32_t), st));
        flat_empty_rows_kernel<<<nblk(n_rows), 256, 0, st>>>(rowptr, n_rows, p->d_red_row, p->d_red_lo, p->d_red_hi, d_count + 2);
        GMR_LAUNCH_CHECK();
        p->n_red_small = n_empty;
    }

    // tiles
    if (nnz > 0) {
        unsigned char* bflag = nullptr;
        int32_t *d_tile0 = nullptr, *tile_start = nullptr;
        GMR_CHECK_CUDA(tmp.get(&bflag, nnz));
        GMR_CHECK_CUDA(tmp.get(&d_tile0, n_blocks + 1));
        GMR_CHECK_CUDA(cudaMemsetAsync(bflag, 0, (size_t)nnz, st));
        flat_boundary_kernel<<<nblk(nnz), 256, 0, st>>>(seg_id, seg_pos, rflag, run_id, skey, d_cs, nnz, bflag);
        GMR_LAUNCH_CHECK();
        // upper bound of the tile count: probes + pieces of long segments + run-cap boundaries
        int64_t cap = n_slots + n_runs / kRunCap + 2;
        for (int64_t b = 0; b < n_blocks; ++b) cap += ((int64_t)cs[b + 1] - cs[b] + kT - 1) / kT;
        GMR_CHECK_CUDA(tmp.get(&tile_start, cap + 1));
        {
            thrust::counting_iterator<int32_t> pos(0);
            size_t bytes = 0;
            GMR_CHECK_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, pos, bflag, tile_start, d_count + 3, (int)nnz, st));
            unsigned char* t = nullptr;
            GMR_CHECK_CUDA(tmp.get(&t, (int64_t)bytes));
            GMR_CHECK_CUDA(cub::DeviceSelect::Flagged(t, bytes, pos, bflag, tile_start, d_count + 3, (int)nnz, st));
        }
        int32_t n_tiles = 0;
        GMR_CHECK_CUDA(cudaMemcpyAsync(&n_tiles, d_count + 3, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        GMR_CHECK_CUDA(cudaStreamSynchronize(st));
        if (n_tiles > cap) {
            set_error("gmr_spmm_blocked_plan_create: internal error, %d tiles exceed the bound %lld", n_tiles, (long long)cap);
            return GMR_ERR_INVALID;
        }
        GMR_CHECK_CUDA(cudaMemcpyAsync(tile_start + n_tiles, &nnz32, sizeof(int32_t), cudaMemcpyHostToDevice, st));
        GMR_CHECK_CUDA(dmalloc(&p->d_tile_rec, (int64_t)kRecInts * n_tiles));
        flat_block_tiles_kernel<<<nblk(n_blocks + 1), 256, 0, st>>>(tile_start, n_tiles, d_cs, (int32_t)n_blocks, d_tile0);
        GMR_LAUNCH_CHECK();
        // padded, 16-byte aligned tile layout
        int32_t *tile_len = nullptr, *tile_off = nullptr;
        GMR_CHECK_CUDA(tmp.get(&tile_len, n_tiles));
        GMR_CHECK_CUDA(tmp.get(&tile_off, n_tiles));
        flat_tile_len_kernel<<<nblk(n_tiles), 256, 0, st>>>(tile_start, n_tiles, tile_len);
        GMR_LAUNCH_CHECK();
        GMR_CHECK_CUDA(exclusive_sum(tmp, tile_len, tile_off, n_tiles, st));
        int32_t last_off = 0, last_len = 0;
        GMR_CHECK_CUDA(cudaMemcpyAsync(&last_off, tile_off + n_tiles - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        GMR_CHECK_CUDA(cudaMemcpyAsync(&last_len, tile_len + n_tiles - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        GMR_CHECK_CUDA(cudaStreamSynchronize(st));
        p->n_padded = (int64_t)last_off + last_len;
        if (p->n_padded >= (int64_t)0x7fffffff) {
            set_error("gmr_spmm_blocked_plan_create: padded layout of %lld entries exceeds int32", (long long)p->n_padded);
            return GMR_ERR_UNSUPPORTED;
        }
        GMR_CHECK_CUDA(dmalloc(&p->d_col, p->n_padded));
        GMR_CHECK_CUDA(dmalloc(&p->d_val, p->n_padded));
        GMR_CHECK_CUDA(dmalloc(&p->d_perm, p->n_padded));
        flat_tile_layout_kernel<<<nblk((int64_t)n_tiles * 16), 256, 0, st>>>(tile_start, tile_off, n_tiles, b_col, b_val, b_perm,
                                                                            p->d_col, p->d_val, p->d_perm);
        GMR_LAUNCH_CHECK();
        GMR_CHECK_CUDA(cudaMemsetAsync(d_count + 3, 0, sizeof(int32_t), st));
        flat_tile_rec_kernel<<<nblk(n_tiles), 256, 0, st>>>(tile_start, tile_off, n_tiles, rflag, run_id, run_row, p->d_tile_rec,
                                                           d_count + 3);
        GMR_LAUNCH_CHECK();
        std::vector<int32_t> t0(n_blocks + 1, 0);
        int32_t bad = 0;
        GMR_CHECK_CUDA(cudaMemcpyAsync(t0.data(), d_tile0, sizeof(int32_t) * (n_blocks + 1), cudaMemcpyDeviceToHost, st));
        GMR_CHECK_CUDA(cudaMemcpyAsync(&bad, d_count + 3, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        GMR_CHECK_CUDA(cudaStreamSynchronize(st));
        if (bad != 0) {
            set_error("gmr_spmm_blocked_plan_create: internal error, %d tiles violate the tiling invariants", bad);
            return GMR_ERR_INVALID;
        }
        for (int64_t b = 0; b <= n_blocks; ++b) p->blk_tile0[b] = t0[b];
        for (int64_t b = 0; b < n_blocks; ++b) p->blk_ntiles[b] = t0[b + 1] - t0[b];
        p->n_tiles = n_tiles;
    }
    if (n_slots >= kSlotBit) {
        set_error("gmr_spmm_blocked_plan_create: %d partial-sum slots exceed 2^30", n_slots);
        return GMR_ERR_UNSUPPORTED;
    }
    p->n_slots = n_slots;
    p->n_runs = n_runs;
    p->n_segments = n_seg;
    p->n_long_segments = n_long;
    GMR_CHECK_CUDA(cudaStreamSynchronize(st));  // scratch arrays die with `tmp`
    return GMR_OK;
}

template <int LANES>
static int launch_flat(const gmr_spmm_bplan* p, const float* X, int64_t ldx, float* Y, int64_t ldy, float* partial,
                       int32_t DV, float alpha, float beta, cudaStream_t st)
{
    constexpr int GROUPS = kFlatThreads / LANES;
    const size_t smem = (size_t)GROUPS * (2 * kStageInts * sizeof(int) + (size_t)kRunCap * LANES * sizeof(float4));
    GMR_REQUIRE(ldx * 4 < (int64_t)0xffffffffll, "gmr_spmm_blocked_f32: row pitch of X exceeds 4 GB");
    auto kern = spmm_flat_kernel<LANES>;
    GMR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t cap = (int64_t)sm_count() * 3;
    for (int32_t b = 0; b < p->n_blocks; ++b) {
        const int64_t nt = p->blk_ntiles[b];
        if (nt == 0) continue;
        const int64_t want = (nt + GROUPS - 1) / GROUPS;
        dim3 grid((unsigned)std::min(want, cap), (unsigned)((DV + LANES - 1) / LANES));
        kern<<<grid, kFlatThreads, smem, st>>>(p->d_col, p->d_val, p->d_tile_rec + (int64_t)kRecInts * p->blk_tile0[b], (int32_t)nt, X,
                                                ldx, Y, ldy, partial, DV, alpha, beta);
        GMR_LAUNCH_CHECK();
    }
    return GMR_OK;
}

}  // namespace gmr

// ---- C ABI -------------------------------------------------------------------------------------------------------

extern "C" int gmr_spmm_blocked_plan_destroy(gmr_spmm_bplan_t* p)
{
    if (p == nullptr) return GMR_OK;
    cudaFree(p->d_col);
    cudaFree(p->d_val);
    cudaFree(p->d_perm);
    cudaFree(p->d_tile_rec);
    cudaFree(p->d_red_row);
    cudaFree(p->d_red_lo);
    cudaFree(p->d_red_hi);
    delete p;
    return GMR_OK;
}

extern "C" int gmr_spmm_blocked_plan_create(gmr_spmm_bplan_t** out, const int32_t* rowptr, const int32_t* col,
                                            const float* val, int64_t n_rows, int64_t n_cols, int64_t block_cols,
                                            void* stream)
{
    GMR_REQUIRE(out != nullptr, "gmr_spmm_blocked_plan_create: null output handle");
    *out = nullptr;
    GMR_REQUIRE(n_rows >= 0 && n_cols >= 0, "gmr_spmm_blocked_plan_create: negative shape");
    GMR_REQUIRE(n_rows < (int64_t)0x40000000 && n_cols < (int64_t)0x7fffffff,
                "gmr_spmm_blocked_plan_create: shape exceeds the index range (rows < 2^30, columns < 2^31)");
    GMR_REQUIRE(n_rows == 0 || rowptr != nullptr, "gmr_spmm_blocked_plan_create: null rowptr");
    if (block_cols <= 0 || block_cols > n_cols) block_cols = std::max<int64_t>(n_cols, 1);
    int64_t n_blocks = n_cols > 0 ? (n_cols + block_cols - 1) / block_cols : 1;
    if (n_blocks > 256) {  // one byte of sort key: widen the blocks instead of failing
        block_cols = (n_cols + 255) / 256;
        n_blocks = (n_cols + block_cols - 1) / block_cols;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int32_t nnz32 = 0;
    if (n_rows > 0) {
        GMR_CHECK_CUDA(cudaMemcpyAsync(&nnz32, rowptr + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        GMR_CHECK_CUDA(cudaStreamSynchronize(st));
    }
    GMR_REQUIRE(nnz32 >= 0, "gmr_spmm_blocked_plan_create: rowptr[n_rows] is negative");
    GMR_REQUIRE(nnz32 == 0 || (col != nullptr && val != nullptr), "gmr_spmm_blocked_plan_create: null col/val with nnz > 0");

    gmr_spmm_bplan* p = new gmr_spmm_bplan();
    p->n_rows = n_rows;
    p->n_cols = n_cols;
    p->nnz = nnz32;
    p->block_cols = block_cols;
    p->n_blocks = (int32_t)n_blocks;
    p->blk_tile0.assign(n_blocks + 1, 0);
    p->blk_ntiles.assign(n_blocks, 0);
    const int rc = gmr::build_plan(p, rowptr, col, val, st);
    if (rc != GMR_OK) {
        gmr_spmm_blocked_plan_destroy(p);
        return rc;
    }
    *out = p;
    return GMR_OK;
}

extern "C" int gmr_spmm_blocked_plan_set_values(gmr_spmm_bplan_t* p, const float* val, void* stream)
{
    GMR_REQUIRE(p != nullptr, "gmr_spmm_blocked_plan_set_values: null plan");
    if (p->nnz == 0) return GMR_OK;
    GMR_REQUIRE(val != nullptr, "gmr_spmm_blocked_plan_set_values: null values");
    gmr::flat_values_kernel<<<gmr::nblk(p->n_padded), 256, 0, (cudaStream_t)stream>>>(p->d_perm, val, p->n_padded, p->d_val);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

extern "C" int gmr_spmm_blocked_plan_stats(const gmr_spmm_bplan_t* p, int64_t* stats /* [8] */)
{
    GMR_REQUIRE(p != nullptr && stats != nullptr, "gmr_spmm_blocked_plan_stats: null argument");
    stats[0] = p->n_blocks;
    stats[1] = p->block_cols;
    stats[2] = p->n_tiles;
    stats[3] = p->n_segments;
    stats[4] = p->n_long_segments;
    stats[5] = p->n_slots;
    stats[6] = p->n_red_small;
    stats[7] = p->n_red_big;
    return GMR_OK;
}

extern "C" int64_t gmr_spmm_blocked_workspace_bytes(const gmr_spmm_bplan_t* p, int32_t D)
{
    if (p == nullptr || D < 1) return 0;
    return gmr::align_up(p->n_slots * (int64_t)D * (int64_t)sizeof(float), 256);
}

extern "C" int gmr_spmm_blocked_f32(const gmr_spmm_bplan_t* p, const float* X, int64_t ldx, float* Y, int64_t ldy,
                                    int32_t D, float alpha, float beta, void* workspace, int64_t workspace_bytes,
                                    void* stream)
{
    using namespace gmr;
    GMR_REQUIRE(p != nullptr, "gmr_spmm_blocked_f32: plan is null");
    GMR_REQUIRE(D >= 4 && D % 4 == 0, "gmr_spmm_blocked_f32: D must be a positive multiple of 4 (got %d)", D);
    GMR_REQUIRE(ldx >= D && ldy >= D && ldx % 4 == 0 && ldy % 4 == 0,
                "gmr_spmm_blocked_f32: leading dimensions (%lld, %lld) must be multiples of 4 and >= D=%d", (long long)ldx,
                (long long)ldy, D);
    if (p->n_rows == 0) return GMR_OK;
    GMR_REQUIRE(X != nullptr && Y != nullptr, "gmr_spmm_blocked_f32: null operand");
    GMR_REQUIRE((uintptr_t)X % 16 == 0 && (uintptr_t)Y % 16 == 0, "gmr_spmm_blocked_f32: X and Y must be 16-byte aligned");
    const int64_t need = gmr_spmm_blocked_workspace_bytes(p, D);
    if (need > 0 && (workspace == nullptr || workspace_bytes < need)) {
        set_error("gmr_spmm_blocked_f32: workspace of %lld bytes required, %lld given", (long long)need,
                  (long long)workspace_bytes);
        return GMR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = (float*)workspace;
    const int32_t DV = D / 4;
    int rc = GMR_OK;
    if (p->nnz > 0)
        rc = (DV <= 16) ? launch_flat<16>(p, X, ldx, Y, ldy, partial, DV, alpha, beta, st)
                        : launch_flat<32>(p, X, ldx, Y, ldy, partial, DV, alpha, beta, st);
    if (rc != GMR_OK) return rc;
    if (p->n_red_small > 0) {
        flat_reduce_small_kernel<<<nblk(p->n_red_small * 16), 256, 0, st>>>(p->d_red_row, p->d_red_lo, p->d_red_hi,
                                                                             (int32_t)p->n_red_small, partial, Y, ldy, DV,
                                                                             alpha, beta);
        GMR_LAUNCH_CHECK();
    }
    if (p->n_red_big > 0) {
        const int32_t *rr = p->d_red_row + p->n_red_small, *lo = p->d_red_lo + p->n_red_small, *hi = p->d_red_hi + p->n_red_small;
        if (DV <= 16)
            flat_reduce_big_kernel<16><<<dim3((unsigned)p->n_red_big, 1), 512, 0, st>>>(rr, lo, hi, partial, Y, ldy, DV, alpha, beta);
        else
            flat_reduce_big_kernel<32><<<dim3((unsigned)p->n_red_big, (unsigned)((DV + 31) / 32)), 512, 0, st>>>(
                rr, lo, hi, partial, Y, ldy, DV, alpha, beta);
        GMR_LAUNCH_CHECK();
    }
    return GMR_OK;
}
