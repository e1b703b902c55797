// train_ops.cu -- training-side fusion: the BPR gathers of calculate_loss.
//
// Replaces, in every model's objective,
//     anc = usr[users]; pos = itm[pos_items]; neg = itm[neg_items]
//     pos_score = (anc * pos).sum(-1); neg_score = (anc * neg).sum(-1)
//   GenMMRec/src/models/diffmm.py:211-219, vbpr.py:84-93, gume.py:281-292,369-371, genrecv1.py:362-372,
//   lightgcn.py:136-146
// and the matching backward (three index_put_(accumulate) scatters of [B, D] gradients).  The reference's form writes
// three [B, D] gathers, two products and -- for the backward -- three more [B, D] gradient tensors; here a warp reads
// the three embedding rows of a triple once and produces two scalars, and the backward accumulates straight into the
// gradient tables.  HBM-bound: 3 rows read per triple forward, 3 rows read + 3 rows accumulated backward.
//
// The backward accumulates with fp32 atomic adds (a user / item may appear in many triples of a batch), like
// PyTorch's own index_put_(accumulate=True) on CUDA: results are reproducible to fp32 rounding of the summation order.
#include "common.cuh"

namespace gmr {

constexpr int kTrainWarps = 8;

// one warp per triple; D floats per row, float4 lanes when VEC4
template <bool VEC4>
__global__ void __launch_bounds__(kTrainWarps * 32)
    bpr_scores_kernel(const float* __restrict__ Eu, int64_t ldu, const float* __restrict__ Ei, int64_t ldi,
                      const int64_t* __restrict__ users, const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                      int32_t B, int32_t D, float* __restrict__ pos_score, float* __restrict__ neg_score)
{
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * kTrainWarps + (threadIdx.x >> 5);
    if (t >= B) return;
    const float* u = Eu + users[t] * ldu;
    const float* p = Ei + pos[t] * ldi;
    const float* n = Ei + neg[t] * ldi;
    float sp = 0.f, sn = 0.f;
    if (VEC4) {
        for (int d = lane * 4; d < D; d += 128) {
            const float4 a = *reinterpret_cast<const float4*>(u + d);
            const float4 b = *reinterpret_cast<const float4*>(p + d);
            const float4 c = *reinterpret_cast<const float4*>(n + d);
            sp = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, sp))));
            sn = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, fmaf(a.w, c.w, sn))));
        }
    } else {
        for (int d = lane; d < D; d += 32) {
            const float a = u[d];
            sp = fmaf(a, p[d], sp);
            sn = fmaf(a, n[d], sn);
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        sp += __shfl_xor_sync(0xffffffffu, sp, m);
        sn += __shfl_xor_sync(0xffffffffu, sn, m);
    }
    if (lane == 0) {
        pos_score[t] = sp;
        neg_score[t] = sn;
    }
}

// dEu[u] += gp * Ei[p] + gn * Ei[n];  dEi[p] += gp * Eu[u];  dEi[n] += gn * Eu[u]
__global__ void __launch_bounds__(kTrainWarps * 32)
    bpr_scores_backward_kernel(const float* __restrict__ Eu, int64_t ldu, const float* __restrict__ Ei, int64_t ldi,
                               const int64_t* __restrict__ users, const int64_t* __restrict__ pos,
                               const int64_t* __restrict__ neg, int32_t B, int32_t D, const float* __restrict__ g_pos,
                               const float* __restrict__ g_neg, float* dEu, int64_t lddu, float* dEi, int64_t lddi)
{
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * kTrainWarps + (threadIdx.x >> 5);
    if (t >= B) return;
    const int64_t iu = users[t], ip = pos[t], in_ = neg[t];
    const float gp = g_pos[t], gn = g_neg[t];
    const float* u = Eu + iu * ldu;
    const float* p = Ei + ip * ldi;
    const float* n = Ei + in_ * ldi;
    float* du = dEu + iu * lddu;
    float* dp = dEi + ip * lddi;
    float* dn = dEi + in_ * lddi;
    for (int d = lane; d < D; d += 32) {
        const float a = u[d];
        atomicAdd(du + d, fmaf(gp, p[d], gn * n[d]));
        atomicAdd(dp + d, gp * a);
        atomicAdd(dn + d, gn * a);
    }
}

}  // namespace gmr

static bool aligned16(const void* p, int64_t ld) { return ((uintptr_t)p % 16 == 0) && (ld % 4 == 0); }

extern "C" int gmr_bpr_scores_f32(const float* Eu, int64_t ldu, const float* Ei, int64_t ldi, const int64_t* users,
                                  const int64_t* pos, const int64_t* neg, int32_t B, int32_t D, float* pos_score,
                                  float* neg_score, void* stream)
{
    GMR_REQUIRE(B >= 0 && D >= 1, "gmr_bpr_scores_f32: bad shape (B=%d, D=%d)", B, D);
    if (B == 0) return GMR_OK;
    GMR_REQUIRE(Eu && Ei && users && pos && neg && pos_score && neg_score, "gmr_bpr_scores_f32: null operand");
    GMR_REQUIRE(ldu >= D && ldi >= D, "gmr_bpr_scores_f32: leading dimensions smaller than D=%d", D);
    const unsigned grid = (unsigned)((B + gmr::kTrainWarps - 1) / gmr::kTrainWarps);
    if (D % 4 == 0 && aligned16(Eu, ldu) && aligned16(Ei, ldi))
        gmr::bpr_scores_kernel<true><<<grid, gmr::kTrainWarps * 32, 0, (cudaStream_t)stream>>>(Eu, ldu, Ei, ldi, users, pos, neg, B,
                                                                                             D, pos_score, neg_score);
    else
        gmr::bpr_scores_kernel<false><<<grid, gmr::kTrainWarps * 32, 0, (cudaStream_t)stream>>>(Eu, ldu, Ei, ldi, users, pos, neg,
                                                                                              B, D, pos_score, neg_score);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}

extern "C" int gmr_bpr_scores_backward_f32(const float* Eu, int64_t ldu, const float* Ei, int64_t ldi, const int64_t* users,
                                           const int64_t* pos, const int64_t* neg, int32_t B, int32_t D, const float* g_pos,
                                           const float* g_neg, float* dEu, int64_t lddu, float* dEi, int64_t lddi,
                                           void* stream)
{
    GMR_REQUIRE(B >= 0 && D >= 1, "gmr_bpr_scores_backward_f32: bad shape (B=%d, D=%d)", B, D);
    if (B == 0) return GMR_OK;
    GMR_REQUIRE(Eu && Ei && users && pos && neg && g_pos && g_neg && dEu && dEi, "gmr_bpr_scores_backward_f32: null operand");
    GMR_REQUIRE(ldu >= D && ldi >= D && lddu >= D && lddi >= D, "gmr_bpr_scores_backward_f32: leading dimensions smaller than D=%d", D);
    const unsigned grid = (unsigned)((B + gmr::kTrainWarps - 1) / gmr::kTrainWarps);
    gmr::bpr_scores_backward_kernel<<<grid, gmr::kTrainWarps * 32, 0, (cudaStream_t)stream>>>(Eu, ldu, Ei, ldi, users, pos, neg, B, D,
                                                                                             g_pos, g_neg, dEu, lddu, dEi, lddi);
    GMR_LAUNCH_CHECK();
    return GMR_OK;
}
