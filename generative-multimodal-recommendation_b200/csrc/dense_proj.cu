// dense_proj.cu -- C[M, 64] = A[M, K] * W[K, 64] in fp32 accuracy on the 5th-gen tensor cores (sm_100a).
//
// Replaces the dense modality projections of the propagation step
//   torch.mm(self.image_embedding.weight, self.image_trans)   GenMMRec/src/models/diffmm.py:115-121 (getImageFeats)
//   torch.mm(self.text_embedding.weight,  self.text_trans)    GenMMRec/src/models/diffmm.py:123-127 (getTextFeats)
// (M = n_items, K = 4096 / 384 feature columns, N = 64): a tall-skinny fp32 GEMM that is HBM-bound on reading A once
// (8.2 GB at the 1M-user shape = 1.26 ms) but runs at CUDA-core speed through cuBLAS (4.4 ms, 60 TFLOP/s fp32 SIMT).
//
// Scheme: split-TF32 with fp32-level accuracy.  Every fp32 operand x is written as hi + lo with hi = x rounded to
// TF32's 11 significant bits and lo = x - hi (exact in fp32, then cut to TF32 as well); the product sums the three
// terms hi.hi + hi.lo + lo.hi on tcgen05.mma.kind::tf32 (3xTF32), so the error is the truncation of lo plus the
// dropped lo.lo term (each <= 2^-22 relative per product, below the fp32 rounding of the sum) plus fp32 accumulation.  Accumulation is two-level: TMEM holds the sum of one K chunk
// (128 columns = 48 MMAs), the epilogue warps add the chunk sums into fp32 registers with IEEE adds, so the
// tensor-core accumulator never carries more than 48 partial products.
//
// Measured (1M-user shape, M = 500k, K = 4096 + 384): 2.33 ms for both projections = 3.9 TB/s of feature bytes (60 % of
// the measured HBM copy peak), against 4.8 ms for cuBLAS fp32 SIMT and a 1.4 ms HBM floor; ncu: tensor pipe 35 % active,
// DRAM 48 %.  History: A in shared memory with 4 terms was shared-memory-bandwidth bound (every MMA re-read its 4 KB A
// slice: 2.9 ms); moving A to tensor memory gave 2.7 ms, dropping lo.lo 2.33 ms.  Tried without effect: 8 splitter warps
// + 4 TMEM A stages, a deeper HBM ring, L2 prefetch of the adjacent K atoms (DRAM page locality).
//
// One persistent CTA per SM, 128 rows of A per tile, warp-specialised:
//   warp 0     TMA producer: per K atom (32 fp32 = 128 B per row) the raw A tile (16 KB) and the hi / lo atoms of W^T
//              (8 KB each, L2-resident) into a 6-stage ring, 128-byte swizzle
//   warps 2-5  splitter (thread = row): read the row's 32 floats from the swizzled tile, split, and store hi / lo as 32
//              TMEM columns each (tcgen05.st) -- the A operand lives in TENSOR MEMORY, so the 16 MMAs of an atom do not
//              re-read 64 KB of A from shared memory (shared-memory bandwidth was the limiter of the first version)
//   warp 1     MMA issuer: 16 tcgen05.mma (M 128, N 64, K 8; A from TMEM, W from shared memory) per atom into a
//              double-buffered TMEM accumulator
//   warps 6-9  epilogue: tcgen05.ld the chunk sums, add into 64 fp32 registers per row, store the row at tile end
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace gmr {

constexpr int kPM = 128;          // rows of A per tile (UMMA M)
constexpr int kPN = 64;           // output columns (UMMA N)
constexpr int kPStages = 6;
constexpr int kPChunkAtoms = 4;   // K atoms per TMEM accumulation chunk (K = 128)
constexpr int kPThreads = 320;
constexpr uint32_t kPAtomA = kPM * 128;   // 16 KB
constexpr uint32_t kPAtomW = kPN * 128;   // 8 KB
constexpr uint32_t kPStageBytes = kPAtomA + 2 * kPAtomW;      // raw A, W hi, W lo = 32 KB

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo)
{
    // hi: round-to-nearest at 11 significant bits (low 13 mantissa bits cleared); lo: exact remainder, cut likewise
    const uint32_t b = __float_as_uint(x);
    hi = __uint_as_float((b + 0x00001000u) & 0xFFFFE000u);
    const float r = x - hi;
    lo = __uint_as_float(__float_as_uint(r) & 0xFFFFE000u);
}

// W [K, N] row-major -> W^T hi / lo, [N, K] each (K-major B operand of the MMA)
__global__ void __launch_bounds__(256)
    proj_split_w_kernel(const float* __restrict__ W, int64_t ldw, int32_t K, int32_t N, float* __restrict__ wt_hi,
                        float* __restrict__ wt_lo)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)K * N) return;
    const int k = (int)(idx / N), n = (int)(idx - (int64_t)k * N);
    float hi, lo;
    split_tf32(W[(int64_t)k * ldw + n], hi, lo);
    wt_hi[(int64_t)n * K + k] = hi;
    wt_lo[(int64_t)n * K + k] = lo;
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// kind::tf32 instruction descriptor: D = F32, A = B = TF32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    // A operand read from tensor memory (lane = row, one 32-bit column per K element), B from shared memory
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_st_32x32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TMEM columns: [0, 128) two accumulators, [128, 256) two A stages of (hi 32 | lo 32) columns
constexpr uint32_t kPTmemCols = 256;
constexpr uint32_t kPTmemA = 2 * kPN;

__global__ void __launch_bounds__(kPThreads, 1)
    dense_proj_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_wh,
                      const __grid_constant__ CUtensorMap map_wl, float* __restrict__ C, int64_t ldc, int32_t M, int32_t K)
{
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kPStages * kPStageBytes);
    uint64_t* full = bars;                        // [kPStages] TMA landed (raw A atom + W hi / lo atoms)
    uint64_t* empty = bars + kPStages;            // [kPStages] MMAs that read the stage's W atoms retired
    uint64_t* split = bars + 2 * kPStages;        // [2] A hi / lo of the TMEM stage written (128 arrivals)
    uint64_t* a_free = bars + 2 * kPStages + 2;   // [2] MMAs that read the TMEM stage retired
    uint64_t* acc_full = bars + 2 * kPStages + 4;     // [2]
    uint64_t* acc_empty = bars + 2 * kPStages + 6;    // [2] (128 arrivals)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPStages + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (M + kPM - 1) / kPM;
    const int n_atoms = K / 32;
    const int n_chunks = (n_atoms + kPChunkAtoms - 1) / kPChunkAtoms;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kPStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&split[s], 128);
            mbar_init(&a_free[s], 1);
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                     "r"(kPTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int k = 0; k < n_atoms; ++k, ++g) {
                    const uint32_t s = g % kPStages, ph = (g / kPStages) & 1u;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + (size_t)s * kPStageBytes;
                    mbar_expect_tx(&full[s], kPAtomA + 2 * kPAtomW);
                    tma_load_2d(st, &map_a, &full[s], k * 32, tile * kPM);
                    tma_load_2d(st + kPAtomA, &map_wh, &full[s], k * 32, 0);
                    tma_load_2d(st + kPAtomA + kPAtomW, &map_wl, &full[s], k * 32, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: A (hi / lo) from tensor memory, W^T (hi / lo) from shared memory =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(kPM, kPN);
            uint32_t g = 0, c = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int ch = 0; ch < n_chunks; ++ch, ++c) {
                    const uint32_t buf = c & 1u, bph = (c >> 1) & 1u;
                    mbar_wait(&acc_empty[buf], bph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * kPN;
                    const int k_end = min(n_atoms, (ch + 1) * kPChunkAtoms);
                    for (int k = ch * kPChunkAtoms; k < k_end; ++k, ++g) {
                        const uint32_t s = g % kPStages, ph = (g / kPStages) & 1u;
                        const uint32_t ta = g & 1u, tph = (g >> 1) & 1u;
                        mbar_wait(&full[s], ph);       // W atoms in shared memory
                        mbar_wait(&split[ta], tph);    // A hi / lo in tensor memory
                        tc_fence_after();
                        const uint32_t w_hi = smem_u32(smem + (size_t)s * kPStageBytes + kPAtomA), w_lo = w_hi + kPAtomW;
                        const uint32_t a_hi = tmem_base + kPTmemA + ta * 64, a_lo = a_hi + 32;
                        const bool first = (k == ch * kPChunkAtoms);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {  // 4 x (K = 8): 8 TMEM columns of A, 32 B of each W row
                            const uint32_t o = kk * 32;
                            // lo.lo (<= 2^-22 of the product, below the fp32 rounding of the sum) is not issued
                            tc_mma_tf32_ts(d_tmem, a_lo + kk * 8, umma_desc_sw128(w_hi + o), idesc, (first && kk == 0) ? 0u : 1u);
                            tc_mma_tf32_ts(d_tmem, a_hi + kk * 8, umma_desc_sw128(w_lo + o), idesc, 1u);
                            tc_mma_tf32_ts(d_tmem, a_hi + kk * 8, umma_desc_sw128(w_hi + o), idesc, 1u);
                        }
                        tc_commit(&empty[s]);
                        tc_commit(&a_free[ta]);
                    }
                    tc_commit(&acc_full[buf]);
                }
            }
        }
    } else if (warp < 6) {
        // ===== splitter: thread = row.  raw A atom (shared memory, 128-byte swizzle) -> hi / lo columns in TMEM =====
        const int q = warp & 3;                     // TMEM lane quarter this warp may touch
        const int row = q * 32 + lane;
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int k = 0; k < n_atoms; ++k, ++g) {
                const uint32_t s = g % kPStages, ph = (g / kPStages) & 1u;
                const uint32_t ta = g & 1u, tph = (g >> 1) & 1u;
                mbar_wait(&full[s], ph);
                const uint8_t* rowp = smem + (size_t)s * kPStageBytes + (size_t)row * 128;
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {       // logical 16-byte chunk j sits at physical chunk j ^ (row & 7)
                    const float4 x = *reinterpret_cast<const float4*>(rowp + ((j ^ (row & 7)) << 4));
                    float h, l;
                    split_tf32(x.x, h, l); hi[4 * j + 0] = __float_as_uint(h); lo[4 * j + 0] = __float_as_uint(l);
                    split_tf32(x.y, h, l); hi[4 * j + 1] = __float_as_uint(h); lo[4 * j + 1] = __float_as_uint(l);
                    split_tf32(x.z, h, l); hi[4 * j + 2] = __float_as_uint(h); lo[4 * j + 2] = __float_as_uint(l);
                    split_tf32(x.w, h, l); hi[4 * j + 3] = __float_as_uint(h); lo[4 * j + 3] = __float_as_uint(l);
                }
                mbar_wait(&a_free[ta], tph ^ 1);    // the MMAs that read this TMEM stage last have retired
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kPTmemA + ta * 64;
                tc_st_32x32(taddr, hi);
                tc_st_32x32(taddr + 32, lo);
                tc_st_wait();
                tc_fence_before();
                mbar_arrive(&split[ta]);
            }
        }
    } else {
        // ===== epilogue: thread = row; chunk sums promoted into fp32 registers =====
        const int q = warp & 3;
        const int row_in_tile = q * 32 + lane;
        uint32_t c = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            float acc[kPN];
#pragma unroll
            for (int j = 0; j < kPN; ++j) acc[j] = 0.f;
            for (int ch = 0; ch < n_chunks; ++ch, ++c) {
                const uint32_t buf = c & 1u, bph = (c >> 1) & 1u;
                mbar_wait(&acc_full[buf], bph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kPN;
                uint32_t v0[32], v1[32];
                tc_ld_32x32(taddr, v0);
                tc_ld_32x32(taddr + 32, v1);
                tc_ld_wait();
                tc_fence_before();
                mbar_arrive(&acc_empty[buf]);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    acc[j] += __uint_as_float(v0[j]);
                    acc[32 + j] += __uint_as_float(v1[j]);
                }
            }
            const int64_t row = (int64_t)tile * kPM + row_in_tile;
            if (row < M) {
                float4* out = reinterpret_cast<float4*>(C + row * ldc);
#pragma unroll
                for (int j = 0; j < kPN / 4; ++j) out[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kPTmemCols) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnP)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFnP proj_encode_fn()
{
    static EncodeTiledFnP fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFnP)p;
    }
    return fn;
}

// [rows, k] fp32 with row pitch ld (elements), boxes of 32 (K) x box_rows, 128-byte swizzle, OOB rows read as zero
static bool proj_make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t k, int64_t ld, int box_rows)
{
    EncodeTiledFnP fn = proj_encode_fn();
    if (fn == nullptr) return false;
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gmr

extern "C" int64_t gmr_dense_proj_workspace_bytes(int32_t K, int32_t N)
{
    if (K <= 0 || N <= 0) return 0;
    return 2 * gmr::align_up((int64_t)K * N * 4, 256);
}

extern "C" int gmr_dense_proj_f32(const float* A, int64_t lda, int32_t M, int32_t K, const float* W, int64_t ldw, int32_t N,
                                  float* C, int64_t ldc, void* workspace, int64_t workspace_bytes, void* stream)
{
    using namespace gmr;
    GMR_REQUIRE(M >= 0 && K >= 1 && N >= 1, "gmr_dense_proj_f32: bad shape M=%d K=%d N=%d", M, K, N);
    if (M == 0) return GMR_OK;
    GMR_REQUIRE(A && W && C, "gmr_dense_proj_f32: null operand");
    if (N != kPN || K % 32 != 0 || lda % 4 != 0 || ldc % 4 != 0 || (uintptr_t)A % 16 != 0 || (uintptr_t)C % 16 != 0 ||
        lda < K || ldw < N || ldc < N) {
        set_error("gmr_dense_proj_f32: needs N == 64, K %% 32 == 0 and 16-byte aligned rows (got M=%d K=%d N=%d lda=%lld ldc=%lld)",
                  M, K, N, (long long)lda, (long long)ldc);
        return GMR_ERR_UNSUPPORTED;
    }
    const int64_t need = gmr_dense_proj_workspace_bytes(K, N);
    if (workspace == nullptr || workspace_bytes < need || (uintptr_t)workspace % 256 != 0) {
        set_error("gmr_dense_proj_f32: 256-byte aligned workspace of %lld bytes required, %lld given", (long long)need,
                  (long long)workspace_bytes);
        return GMR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* wt_hi = (float*)workspace;
    float* wt_lo = (float*)((uint8_t*)workspace + need / 2);
    {
        const int64_t n = (int64_t)K * N;
        proj_split_w_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, ldw, K, N, wt_hi, wt_lo);
        GMR_LAUNCH_CHECK();
    }
    CUtensorMap map_a, map_wh, map_wl;
    if (!proj_make_map(&map_a, A, M, K, lda, kPM) || !proj_make_map(&map_wh, wt_hi, N, K, K, kPN) ||
        !proj_make_map(&map_wl, wt_lo, N, K, K, kPN)) {
        set_error("gmr_dense_proj_f32: cuTensorMapEncodeTiled unavailable or failed");
        return GMR_ERR_CUDA;
    }
    const size_t smem = (size_t)kPStages * kPStageBytes + (2 * kPStages + 10) * sizeof(uint64_t) + 1024;
    GMR_CHECK_CUDA(cudaFuncSetAttribute(dense_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_tiles = (M + kPM - 1) / kPM;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    dense_proj_kernel<<<grid, kPThreads, smem, st>>>(map_a, map_wh, map_wl, C, ldc, M, K);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}
