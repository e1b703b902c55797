// score_topk_tc.cu -- K2 + K3: fused score + mask + top-K with the dense contraction on the 5th-gen
// tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed operands), sm_100a only.
//
// Replaces the same reference call chain as score_topk_simt.cu
//   torch.matmul(u_e[user], i_e.T) -> scores[mask] = -1e10 -> torch.topk   (models/diffmm.py:276-278,
//   common/trainer.py:381-386) and returns THE SAME ids/scores as the fp32 path:
//
//   1. split: every fp32 operand x is written as bf16 terms hi = bf16(x), mid = bf16(x - hi).  With
//      A' = [hi | hi | mid] (users) and B' = [hi | mid | hi] (items) along K, one bf16 GEMM over
//      K' = 3 D yields  hi.hi + hi.mid + mid.hi,  i.e. the fp32 product up to
//      |err| <= kSplitErr * |u|_2 * max_i |e_i|_2  (the dropped terms are O(2^-16) relative).
//   2. a persistent, warp-specialised CTA per 128-user tile sweeps all item tiles:
//        warp 0   TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, mbarrier tx)
//        warp 1   MMA issuer     (one elected thread: 3*D/16 tcgen05.mma per 128 x 128 tile into a
//                                 double-buffered TMEM accumulator; tcgen05.commit -> mbarriers)
//        warps 2-5 epilogue      (tcgen05.ld 32 lanes x 32 columns; thread = user row: bias, running
//                                 threshold filter, mask lookup only for survivors, append to the
//                                 row's key slots; warp-cooperative bitonic compaction)
//      The [B, I] score matrix exists only as 128 x 128 fp32 tiles in TMEM.
//   3. at the end of a user tile the epilogue warps re-score the KP >= K + 8 surviving candidates of
//      each row with the exact fp32 fmaf chain, sort them by (score desc, id asc) and write the top K.
//      A row is CERTIFIED exact when  t + eps < s_K  (t = approximate score of its weakest kept
//      candidate, eps the bound above, s_K its K-th exact score): no item outside the candidate list
//      can then reach the top K.  Uncertified rows (rare: needs >= KP - K near-ties) are queued
//      and redone by the fp32 kernel, so both precision modes return identical results.
//   4. wide operands (D > 256: the modality feature tables of the kNN item-item graphs, D = 4096 / 384,
//      GenMMRec/src/utils/utils.py:147-197) do not fit a resident A' tile: in STREAM_A mode every ring stage carries
//      one K-atom of A' AND one of B' (32 KB), both TMA-fed, and the accumulator integrates 3 D / 64 atoms per tile
//      in TMEM -- a K-chunked GEMM with the same fused top-K epilogue.  This is the tensor-core kNN builder
//      (SURVEY.md section 8f rank 1).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "topk_select.cuh"

namespace gmr {

// fp32 kernel entry used for the uncertified rows (score_topk_simt.cu)
int score_topk_simt_launch(const float* Eu, int64_t lde_u, const int64_t* users, const int32_t* row_map,
                           int32_t n_rows, const float* Ei, int64_t lde_i, const float* bias, int32_t I, int32_t D,
                           const int64_t* mask_rowptr, const int32_t* mask_items, int32_t K, int32_t* out_ids,
                           float* out_scores, void* workspace, cudaStream_t st, int grid_override);
int64_t score_simt_workspace_bytes(int32_t B, int32_t K);
void score_simt_set_dynamic_rows(const int32_t* n_rows_dev);

constexpr int kTM = 128;          // users per tile (UMMA M)
constexpr int kTN = 128;          // items per tile (UMMA N)
constexpr int kMaxStages = 12;    // B' K-atom ring (as many 16 KB stages as fit beside the A' tile)
constexpr int kTcThreads = 192;   // 6 warps
constexpr float kSplitErr = 6.2e-5f;  // 3 * 2^-16 (dropped split terms) + fp32 accumulation slack

// ---- 1. operand split -----------------------------------------------------------------------------
// out[r, :] = [hi | hi | mid] (is_a) or [hi | mid | hi] over 3 * D bf16; rows >= n_rows are zero.
__global__ void __launch_bounds__(256)
    split_bf16_kernel(const float* __restrict__ E, int64_t lde, const int64_t* __restrict__ rows, int32_t n_rows,
                      int32_t n_rows_pad, int32_t D, int is_a, __nv_bfloat16* __restrict__ out,
                      float* __restrict__ row_norm, uint32_t* __restrict__ max_norm_bits)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows_pad) return;
    __nv_bfloat16* o = out + r * (3 * (int64_t)D);
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
        float x = 0.f;
        if (r < n_rows) x = E[(rows ? rows[r] : r) * lde + d];
        const __nv_bfloat16 hi = __float2bfloat16_rn(x);
        const __nv_bfloat16 mid = __float2bfloat16_rn(x - __bfloat162float(hi));
        o[d] = hi;
        o[D + d] = is_a ? hi : mid;
        o[2 * D + d] = is_a ? mid : hi;
        ss = fmaf(x, x, ss);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
    if (lane == 0) {
        const float nrm = sqrtf(ss) * 1.0000002f;  // round up: this feeds an error BOUND
        if (row_norm != nullptr && r < n_rows) row_norm[r] = nrm;
        if (max_norm_bits != nullptr) atomicMax(max_norm_bits, __float_as_uint(nrm));
    }
}

// ---- 2. + 3. the fused kernel ---------------------------------------------------------------------
struct TcArgs {
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
    uint64_t* slots;          // [grid][kTM][CAP]
    const float* a_norm;      // [B]
    const uint32_t* b_max_norm_bits;
    int32_t* fallback_rows;   // [B]
    int32_t* fallback_count;  // [1]
    int32_t n_stages;
    float split_err;  // error bound of the split product, relative to |u| max|e| (grows with the accumulation length)
    uint32_t* dbg;  // misc counters
    int32_t debug;  // GMR_TC_DEBUG: 1 = epilogue only drains TMEM, 2 = filter without appends (timing experiments)
};

// dynamic shared memory layout (1024-byte aligned): A' atoms (resident per user tile) | ring of B'
// atoms (one 128 x 64 bf16 K-atom = 16 KB per stage) | bias tile | barriers | row states
template <int NPL, bool HAS_BIAS, bool STREAM_A>
__global__ void __launch_bounds__(kTcThreads, STREAM_A ? 1 : 2)
    score_topk_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TcArgs a)
{
    constexpr int CAP = 32 * NPL;
    constexpr int KP = CAP / 4;
    constexpr int TRIG = CAP - 32;

    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem_raw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int n_atoms = (3 * a.D) / 64;                       // K' / 64
    const int n_stages = a.n_stages;
    const uint32_t atom_a_bytes = kTM * 128, atom_b_bytes = kTN * 128;
    // resident mode: A' atoms | ring of B' atoms.  STREAM_A: ring of (A' atom | B' atom) pairs.
    const uint32_t stage_bytes = STREAM_A ? (atom_a_bytes + atom_b_bytes) : atom_b_bytes;
    uint8_t* sm_a = smem_raw;
    uint8_t* sm_b = STREAM_A ? smem_raw : sm_a + (size_t)n_atoms * atom_a_bytes;   // ring base
    uint8_t* tail = sm_b + (size_t)n_stages * stage_bytes;
    float* sm_bias = reinterpret_cast<float*>(tail);          // [2][kTN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 2 * kTN * sizeof(float));
    uint64_t* full_b = bars;                    // [kMaxStages]
    uint64_t* empty_b = bars + kMaxStages;      // [kMaxStages]
    uint64_t* tmem_full = bars + 2 * kMaxStages;      // [2]
    uint64_t* tmem_empty = bars + 2 * kMaxStages + 2; // [2]
    uint64_t* a_full = bars + 2 * kMaxStages + 4;
    uint64_t* a_empty = bars + 2 * kMaxStages + 5;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 6);
    RowState* rows = reinterpret_cast<RowState*>(bars + 2 * kMaxStages + 8);  // [kTM]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_utiles = (a.B + kTM - 1) / kTM;
    const int n_itiles = (a.I + kTN - 1) / kTN;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&empty_b[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 128);
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                     "r"(2 * kTN)
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
            int stage = 0;
            uint32_t phase = 0, a_phase = 0;
            for (int ut = blockIdx.x; ut < n_utiles; ut += gridDim.x) {
                if (!STREAM_A) {
                    mbar_wait(a_empty, a_phase ^ 1);  // first pass: passes immediately (fresh barrier)
                    mbar_expect_tx(a_full, (uint32_t)n_atoms * atom_a_bytes);
                    for (int k = 0; k < n_atoms; ++k) tma_load_2d(sm_a + (size_t)k * atom_a_bytes, &map_a, a_full, k * 64, ut * kTM);
                    a_phase ^= 1;
                }
                for (int it = 0; it < n_itiles; ++it) {
                    for (int k = 0; k < n_atoms; ++k) {
                        mbar_wait(&empty_b[stage], phase ^ 1);
                        mbar_expect_tx(&full_b[stage], stage_bytes);
                        uint8_t* sb = sm_b + (size_t)stage * stage_bytes;
                        if (STREAM_A) {
                            tma_load_2d(sb, &map_a, &full_b[stage], k * 64, ut * kTM);
                            sb += atom_a_bytes;
                        }
                        tma_load_2d(sb, &map_b, &full_b[stage], k * 64, it * kTN);
                        if (++stage == n_stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread) =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kTM, kTN);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, a_phase = 0;
            for (int ut = blockIdx.x; ut < n_utiles; ut += gridDim.x) {
                if (!STREAM_A) {
                    mbar_wait(a_full, a_phase);
                    a_phase ^= 1;
                }
                for (int it = 0; it < n_itiles; ++it) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)acc * kTN;
                    for (int k = 0; k < n_atoms; ++k) {
                        mbar_wait(&full_b[stage], phase);
                        tc_fence_after();
                        const uint32_t st_base = smem_u32(sm_b + (size_t)stage * stage_bytes);
                        const uint32_t a_atom = STREAM_A ? st_base : smem_u32(sm_a) + k * atom_a_bytes;
                        const uint32_t b_base = STREAM_A ? st_base + atom_a_bytes : st_base;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {  // 4 x (K = 16 bf16 = 32 B) per 128-byte swizzle row
                            const uint64_t ad = umma_desc_sw128(a_atom + kk * 32);
                            const uint64_t bd = umma_desc_sw128(b_base + kk * 32);
                            tc_mma_bf16(d_tmem, ad, bd, idesc, (k | kk) ? 1u : 0u);
                        }
                        tc_commit(&empty_b[stage]);  // this B' atom may be overwritten once its MMAs retire
                        if (++stage == n_stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    tc_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
                    if (++acc == 2) {
                        acc = 0;
                        acc_phase ^= 1;
                    }
                }
                if (!STREAM_A) tc_commit(a_empty);  // A' tile free once every MMA of this user tile retired
            }
        }
    } else {
        // ===== epilogue: thread = user row (TMEM lane), warps 2..5 own lane quarters (warp % 4) =====
        const int q = warp & 3;
        const int row = q * 32 + lane;  // row inside the user tile == TMEM lane
        uint64_t* cta_slots = a.slots + (int64_t)blockIdx.x * kTM * CAP;
        uint64_t* my_slots = cta_slots + (int64_t)row * CAP;
        RowState* st = &rows[row];
        const int et = threadIdx.x - 64;  // 0..127 among epilogue threads
        const float b_max = __uint_as_float(*a.b_max_norm_bits);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int ut = blockIdx.x; ut < n_utiles; ut += gridDim.x) {
            const int b = ut * kTM + row;
            const bool valid = b < a.B;
            int64_t mlo = 0, mhi = 0;
            if (valid && a.mask_rowptr != nullptr) {
                mlo = a.mask_rowptr[b];
                mhi = a.mask_rowptr[b + 1];
            }
            int cnt = 0;
            float thr = -INFINITY;
            // train-history mask: the item tiles are swept in ascending id order and the row's mask list is
            // ascending too, so a cursor replaces any search; masked columns are never appended.  A row that
            // would have fewer than KP unmasked items is left to the exact fp32 kernel (it also owns the
            // "masked items surface once the unmasked run out" semantics of trainer.py:384).
            int64_t mcur = mlo;
            int mnext = (mcur < mhi) ? a.mask_items[mcur] : 0x7fffffff;
            const bool force_exact = valid && ((int64_t)a.I - (mhi - mlo) < (int64_t)KP);
            st->cnt = 0;
            st->thr_score = -INFINITY;
            st->thr_key = 0ull;
            __syncwarp();

            for (int it = 0; it < n_itiles; ++it) {
                const int i0 = it * kTN;
                if (HAS_BIAS) {
                    // stage the bias tile (double-buffered with the accumulator) and sync the 4 warps
                    sm_bias[acc * kTN + et] = (i0 + et < a.I) ? a.bias[i0 + et] : 0.f;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * kTN;
                // 32 accumulator columns of this thread's row: hierarchical max (4 groups of 8) against
                // the row threshold, scan only the groups that can contain a survivor
                // One chunk = this thread's row x 32 accumulator columns.  The loop body is kept SMALL on
                // purpose (not unrolled over chunks): a filter unrolled over chunks and bias variants
                // is tens of KB of SASS and starves on instruction fetch.
#pragma unroll 1
                for (int c0 = 0; c0 < kTN; c0 += 32) {
                    uint32_t v[32];
                    tc_ld_32x32(taddr + c0, v);
                    tc_ld_wait();
                    if (a.debug == 1) continue;
                    const int cbase = i0 + c0;
                    uint32_t mbits = 0u;  // masked columns of this chunk
                    while (mnext < cbase + 32) {
                        if (mnext >= cbase) mbits |= 1u << (mnext - cbase);
                        ++mcur;
                        mnext = (mcur < mhi) ? a.mask_items[mcur] : 0x7fffffff;
                    }
                    // hierarchical max (4 groups of 8 columns) against the row threshold: only groups that can
                    // hold a survivor are scanned, branch-free (predicated store + counter bump)
                    float gm[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float m = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float sj = __uint_as_float(v[g * 8 + j]);
                            if (HAS_BIAS) {
                                sj += sm_bias[acc * kTN + c0 + g * 8 + j];
                                v[g * 8 + j] = __float_as_uint(sj);
                            }
                            m = fmaxf(m, sj);
                        }
                        gm[g] = m;
                    }
                    const float mx = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
                    if (valid && mx >= thr && a.debug != 2) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (gm[g] >= thr) {
                                const int ibase = cbase + g * 8;
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float sj = __uint_as_float(v[g * 8 + j]);
                                    const bool take = (sj >= thr) && (ibase + j < a.I) && !((mbits >> (g * 8 + j)) & 1u);
                                    const uint64_t key = make_key(sj, ibase + j);
                                    if (take) my_slots[cnt] = key;
                                    cnt += take ? 1 : 0;
                                }
                            }
                        }
                    }
                    // rows about to run out of slots: cheap conservative pruning (exact compaction only when
                    // ties defeat it)
                    unsigned need = __ballot_sync(0xffffffffu, cnt >= TRIG);
                    if (need) {
                        st->cnt = cnt;
                        __syncwarp();
                        while (need) {
                            const int l = __ffs(need) - 1;
                            need &= need - 1;
                            uint64_t* s_row = cta_slots + (int64_t)(q * 32 + l) * CAP;
                            if (a.debug == 5 || !prune_row<NPL>(s_row, &rows[q * 32 + l], lane))
                                compact_row<NPL>(s_row, &rows[q * 32 + l], KP - 1, lane);
                        }
                        cnt = st->cnt;
                        thr = st->thr_score;
                    }
                }
                tc_fence_before();
                mbar_arrive(&tmem_empty[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }

            // ---- user tile done: exact re-score, sort, certify, write (warp handles its 32 rows) ----
            st->cnt = cnt;
            __syncwarp();
            for (int l = 0; l < 32; ++l) {
                const int r = q * 32 + l;
                const int rb = ut * kTM + r;
                if (rb >= a.B) break;  // rows are contiguous: warp-uniform exit
                uint64_t* s_row = cta_slots + (int64_t)r * CAP;
                compact_row<NPL>(s_row, &rows[r], KP - 1, lane);
                const int n_cand = rows[r].cnt;  // <= KP, sorted by approximate key
                const float t_approx = (n_cand >= KP) ? key_score(s_row[KP - 1]) : -INFINITY;
                const float* u = a.Eu + (a.users ? a.users[rb] : (int64_t)rb) * a.lde_u;
                uint64_t ek[NPL / 4];
#pragma unroll
                for (int c = 0; c < NPL / 4; ++c) {
                    const int j = c * 32 + lane;
                    uint64_t key = 0ull;
                    if (j < n_cand) {
                        const uint64_t ak = s_row[j];
                        const int item = key_id(ak);
                        float s;
                        if (key_score(ak) == -1e10f) {
                            s = -1e10f;  // masked (trainer.py:384): exact by definition
                        } else {
                            const float* e = a.Ei + (int64_t)item * a.lde_i;
                            s = a.bias ? a.bias[item] : 0.f;
                            for (int d = 0; d < a.D; ++d) s = fmaf(u[d], e[d], s);
                        }
                        key = make_key(s, item);
                    }
                    ek[c] = key;
                }
                warp_bitonic_sort_desc<NPL / 4>(ek, lane);
                // K-th exact key -> certification
                const int kth = a.K - 1;
                uint64_t kth_key = 0ull;
#pragma unroll
                for (int c = 0; c < NPL / 4; ++c) {
                    const uint64_t cand = __shfl_sync(0xffffffffu, ek[c], kth & 31);
                    if (c == (kth >> 5)) kth_key = cand;
                }
                bool certified = true;
                if (n_cand >= KP) {  // otherwise every item that could matter is in the list
                    const float eps = a.split_err * a.a_norm[rb] * b_max + 4.8e-7f * fabsf(t_approx);
                    certified = (kth_key != 0ull) && (t_approx + eps < key_score(kth_key));
                }
#pragma unroll
                for (int c = 0; c < NPL / 4; ++c) {
                    const int j = c * 32 + lane;
                    if (j < a.K) {
                        const bool ok = ek[c] != 0ull;
                        a.out_ids[(int64_t)rb * a.K + j] = ok ? key_id(ek[c]) : -1;
                        if (a.out_scores) a.out_scores[(int64_t)rb * a.K + j] = ok ? key_score(ek[c]) : -INFINITY;
                    }
                }
                if (__shfl_sync(0xffffffffu, (int)force_exact, l)) certified = false;
                if (!certified && lane == 0) a.fallback_rows[atomicAdd(a.fallback_count, 1)] = rb;
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kTN) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
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

// [rows, kd] bf16 row-major, boxes of 64 (K) x box_rows, 128-byte swizzle
static bool make_map(CUtensorMap* m, void* base, int64_t rows, int64_t kd, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    cuuint64_t dims[2] = {(cuuint64_t)kd, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kd * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int tc_stages_for(int32_t D);
static int tc_kp(int32_t K) { return K + 8 <= 64 ? 64 : (K + 8 <= 128 ? 128 : 256); }

constexpr int32_t kTcMaxD = 8192;     // widest operand of the streamed (K-chunked) mode
constexpr int32_t kTcResidentD = 256; // up to here the A' tile stays resident in shared memory
static bool tc_stream(int32_t D) { return D > kTcResidentD; }

bool score_tc_supported(int32_t D, int32_t K) { return D % 64 == 0 && D >= 64 && D <= kTcMaxD && K + 8 <= 256 && tc_stages_for(D) >= 2; }

static bool tc_stream(int32_t D);
static int tc_grid(int32_t B, int32_t D)
{
    const int tiles = (B + kTM - 1) / kTM;
    const int cap = (tc_stream(D) ? 1 : 2) * sm_count();   // resident CTAs per SM: 1 in the streamed mode
    return tiles < cap ? tiles : cap;
}

struct TcLayout {
    int64_t a_split, b_split, a_norm, misc, fallback, slots, simt, total;
    int32_t b_pad, i_pad;
};

static TcLayout tc_layout(int32_t B, int32_t I, int32_t D, int32_t K)
{
    TcLayout L;
    L.b_pad = (B + kTM - 1) / kTM * kTM;
    L.i_pad = (I + kTN - 1) / kTN * kTN;
    const int cap = 4 * tc_kp(K);
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        const int64_t o = off;
        off += align_up(bytes, 256);
        return o;
    };
    L.a_split = take((int64_t)L.b_pad * 3 * D * 2);
    L.b_split = take((int64_t)L.i_pad * 3 * D * 2);
    L.a_norm = take((int64_t)L.b_pad * 4);
    L.misc = take(256);  // [0] max item norm bits, [1] fallback count
    L.fallback = take((int64_t)B * 4);
    L.slots = take((int64_t)tc_grid(B, D) * kTM * cap * 8);
    L.simt = take(score_simt_workspace_bytes(B, K));
    L.total = off;
    return L;
}

int64_t score_tc_workspace_bytes(int32_t B, int32_t I, int32_t D, int32_t K) { return tc_layout(B, I, D, K).total; }


static size_t tc_smem_fixed(int32_t D)
{
    const int n_atoms = tc_stream(D) ? 0 : 3 * D / 64;   // streamed mode keeps no resident A' tile
    return (size_t)n_atoms * kTM * 128 + 2 * kTN * sizeof(float) + (2 * kMaxStages + 8) * sizeof(uint64_t) +
           kTM * sizeof(RowState) + 1024;
}
static int tc_stages(int32_t D)
{
    if (tc_stream(D)) {  // one CTA per SM, ring of (A' atom | B' atom) pairs
        const int s = (int)(((int64_t)227 * 1024 - (int64_t)tc_smem_fixed(D)) / ((kTM + kTN) * 128));
        return s > kMaxStages ? kMaxStages : s;
    }
    // two CTAs per SM when they fit (D = 64): 8 epilogue warps per SM hide each other's latencies and
    // the two MMA streams share the tensor pipe; otherwise one CTA with a deep ring
    int64_t room = (int64_t)113 * 1024 - (int64_t)tc_smem_fixed(D);
    if (room < 3 * kTN * 128) room = (int64_t)227 * 1024 - (int64_t)tc_smem_fixed(D);
    const int s = (int)(room / (kTN * 128));
    return s > kMaxStages ? kMaxStages : s;
}

static int tc_stages_for(int32_t D) { return tc_stages(D); }

// rows the last tensor-core call could not certify (they were redone on the fp32 path); synchronises
int score_tc_fallback_count(const void* workspace, int32_t B, int32_t I, int32_t D, int32_t K, int32_t* count_host,
                            cudaStream_t st)
{
    const TcLayout L = tc_layout(B, I, D, K);
    GMR_CHECK_CUDA(cudaMemcpyAsync(count_host, (const uint8_t*)workspace + L.misc + 4, sizeof(int32_t),
                                   cudaMemcpyDeviceToHost, st));
    GMR_CHECK_CUDA(cudaStreamSynchronize(st));
    return GMR_OK;
}

int score_topk_tc_launch(const float* Eu, int64_t lde_u, const int64_t* users, int32_t B, const float* Ei,
                         int64_t lde_i, const float* bias, int32_t I, int32_t D, const int64_t* mask_rowptr,
                         const int32_t* mask_items, int32_t K, int32_t* out_ids, float* out_scores, void* workspace,
                         int64_t workspace_bytes, cudaStream_t st)
{
    const TcLayout L = tc_layout(B, I, D, K);
    if (workspace_bytes < L.total) {
        set_error("score_topk_tc: workspace of %lld bytes required, %lld given", (long long)L.total, (long long)workspace_bytes);
        return GMR_ERR_WORKSPACE;
    }
    if ((uintptr_t)workspace % 256 != 0) {
        set_error("score_topk_tc: workspace must be 256-byte aligned");
        return GMR_ERR_INVALID;
    }
    const int n_stages = tc_stages(D);
    if (n_stages < 2) {
        set_error("score_topk_tc: D=%d leaves no room for the operand ring in shared memory", D);
        return GMR_ERR_UNSUPPORTED;
    }
    const bool stream = tc_stream(D);
    const size_t smem = tc_smem_fixed(D) + (size_t)n_stages * (stream ? (kTM + kTN) * 128 : kTN * 128);
    uint8_t* ws = (uint8_t*)workspace;
    __nv_bfloat16* a_split = (__nv_bfloat16*)(ws + L.a_split);
    __nv_bfloat16* b_split = (__nv_bfloat16*)(ws + L.b_split);
    float* a_norm = (float*)(ws + L.a_norm);
    uint32_t* misc = (uint32_t*)(ws + L.misc);
    int32_t* fallback = (int32_t*)(ws + L.fallback);

    GMR_CHECK_CUDA(cudaMemsetAsync(misc, 0, 256, st));
    const int wpb = 8;
    split_bf16_kernel<<<(L.b_pad + wpb - 1) / wpb, wpb * 32, 0, st>>>(Eu, lde_u, users, B, L.b_pad, D, 1, a_split, a_norm, nullptr);
    GMR_LAUNCH_CHECK();
    split_bf16_kernel<<<(L.i_pad + wpb - 1) / wpb, wpb * 32, 0, st>>>(Ei, lde_i, nullptr, I, L.i_pad, D, 0, b_split, nullptr, misc);
    GMR_LAUNCH_CHECK();

    CUtensorMap map_a, map_b;
    if (!make_map(&map_a, a_split, L.b_pad, 3 * (int64_t)D, kTM) || !make_map(&map_b, b_split, L.i_pad, 3 * (int64_t)D, kTN)) {
        set_error("score_topk_tc: cuTensorMapEncodeTiled unavailable or failed");
        return GMR_ERR_CUDA;
    }
    TcArgs a;
    a.Eu = Eu; a.lde_u = lde_u; a.users = users; a.B = B; a.Ei = Ei; a.lde_i = lde_i; a.bias = bias; a.I = I; a.D = D;
    a.mask_rowptr = mask_rowptr; a.mask_items = mask_items; a.K = K; a.out_ids = out_ids; a.out_scores = out_scores;
    a.slots = (uint64_t*)(ws + L.slots); a.a_norm = a_norm; a.b_max_norm_bits = misc;
    a.fallback_rows = fallback; a.fallback_count = (int32_t*)(misc + 1); a.n_stages = n_stages;
    // split error: 3 dropped 2^-16 terms (+5 %) and, for long accumulations, the worst-case fp32 summation error of the
    // 3 D products; D <= 256 keeps the constant its certification tests were run with
    a.split_err = stream ? (4.8e-5f + 3.0f * (float)D * 5.97e-8f) : kSplitErr;
    a.dbg = misc;
    a.debug = getenv("GMR_TC_DEBUG") ? atoi(getenv("GMR_TC_DEBUG")) : 0;
    const int grid = tc_grid(B, D);
    const int kp = tc_kp(K);
#define GMR_TC_LAUNCH1(NPL, BIAS, STREAM)                                                                          \
    do {                                                                                                           \
        GMR_CHECK_CUDA(cudaFuncSetAttribute(score_topk_tc_kernel<NPL, BIAS, STREAM>,                               \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        score_topk_tc_kernel<NPL, BIAS, STREAM><<<grid, kTcThreads, smem, st>>>(map_a, map_b, a);                  \
    } while (0)
#define GMR_TC_LAUNCH(NPL)                                                                                         \
    do {                                                                                                           \
        if (stream) {                                                                                              \
            if (bias != nullptr) GMR_TC_LAUNCH1(NPL, true, true); else GMR_TC_LAUNCH1(NPL, false, true);           \
        } else {                                                                                                   \
            if (bias != nullptr) GMR_TC_LAUNCH1(NPL, true, false); else GMR_TC_LAUNCH1(NPL, false, false);         \
        }                                                                                                          \
    } while (0)
    if (kp == 64)
        GMR_TC_LAUNCH(8);
    else if (kp == 128)
        GMR_TC_LAUNCH(16);
    else
        GMR_TC_LAUNCH(32);
#undef GMR_TC_LAUNCH
#undef GMR_TC_LAUNCH1
    GMR_LAUNCH_CHECK();
    // uncertified rows: exact fp32 kernel, row count read on the device (no host synchronisation)
    score_simt_set_dynamic_rows((const int32_t*)(misc + 1));
    const int fb_grid = 2 * sm_count() < (B + 127) / 128 ? 2 * sm_count() : (B + 127) / 128;
    int rc = score_topk_simt_launch(Eu, lde_u, users, fallback, B, Ei, lde_i, bias, I, D, mask_rowptr, mask_items, K,
                                    out_ids, out_scores, ws + L.simt, st, fb_grid);
    score_simt_set_dynamic_rows(nullptr);
    if (a.debug == 9) {
        uint32_t h[4];
        cudaMemcpyAsync(h, misc, sizeof(h), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr, "[gmr tc debug] appends=%u compactions=%u fallback=%u rows=%d items=%d\n", h[2], h[3], h[1], B, I);
    }
    return rc;
}

}  // namespace gmr
