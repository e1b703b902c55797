// api.cu -- error state, device queries, K2 entry point dispatch and peer-memory plumbing of libgmr.
#include <cstring>

#include "common.cuh"

namespace gmr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// score_topk_simt.cu
int64_t score_simt_workspace_bytes(int32_t B, int32_t K);
int score_topk_simt_launch(const float* Eu, int64_t lde_u, const int64_t* users, const int32_t* row_map,
                           int32_t n_rows, const float* Ei, int64_t lde_i, const float* bias, int32_t I, int32_t D,
                           const int64_t* mask_rowptr, const int32_t* mask_items, int32_t K, int32_t* out_ids,
                           float* out_scores, void* workspace, cudaStream_t st, int grid_override);
int scores_dense_launch(const float* Eu, int64_t lde_u, const int64_t* users, int32_t B, const float* Ei,
                        int64_t lde_i, const float* bias, int32_t I, int32_t D, float* out, int64_t ldo,
                        cudaStream_t st);
// score_topk_tc.cu
bool score_tc_supported(int32_t D, int32_t K);
int64_t score_tc_workspace_bytes(int32_t B, int32_t I, int32_t D, int32_t K);
int score_topk_tc_launch(const float* Eu, int64_t lde_u, const int64_t* users, int32_t B, const float* Ei,
                         int64_t lde_i, const float* bias, int32_t I, int32_t D, const int64_t* mask_rowptr,
                         const int32_t* mask_items, int32_t K, int32_t* out_ids, float* out_scores, void* workspace,
                         int64_t workspace_bytes, cudaStream_t st);

int score_tc_fallback_count(const void* workspace, int32_t B, int32_t I, int32_t D, int32_t K, int32_t* count_host,
                            cudaStream_t st);
// score_topk_screen.cu
bool score_screen_supported(int32_t D, int32_t K);
int64_t score_screen_workspace_bytes(int32_t B, int32_t I, int32_t D, int32_t K);
int score_topk_screen_launch(const float* Eu, int64_t lde_u, const int64_t* users, int32_t B, const float* Ei,
                             int64_t lde_i, const float* bias, int32_t I, int32_t D, const int64_t* mask_rowptr,
                             const int32_t* mask_items, int32_t K, int32_t* out_ids, float* out_scores, void* workspace,
                             int64_t workspace_bytes, cudaStream_t st);
int score_screen_fallback_count(const void* workspace, int32_t B, int32_t I, int32_t D, int32_t K, int32_t* count_host,
                                cudaStream_t st);
int score_screen_stats(const void* workspace, int32_t B, int32_t I, int32_t D, int32_t K, uint64_t* stats_host,
                       cudaStream_t st);

}  // namespace gmr

extern "C" const char* gmr_last_error(void) { return gmr::g_err; }

extern "C" int gmr_abi_version(void) { return GMR_ABI_VERSION; }

extern "C" int gmr_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* l2_bytes)
{
    int dev = 0;
    GMR_CHECK_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) {
        GMR_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sm_count = v;
    }
    if (cc_major) {
        GMR_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        *cc_major = v;
    }
    if (cc_minor) {
        GMR_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        *cc_minor = v;
    }
    if (l2_bytes) {
        GMR_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev));
        *l2_bytes = v;
    }
    return GMR_OK;
}

extern "C" int64_t gmr_score_topk_workspace_bytes(int32_t B, int32_t I, int32_t D, int32_t K, int32_t precision)
{
    if (B <= 0 || I <= 0 || D <= 0 || K <= 0 || K > GMR_MAX_TOPK) return 0;
    int64_t simt = gmr::score_simt_workspace_bytes(B, K);
    if (precision == GMR_SCORE_TC && gmr::score_screen_supported(D, K)) return gmr::score_screen_workspace_bytes(B, I, D, K);
    if (precision == GMR_SCORE_TC_SPLIT && gmr::score_tc_supported(D, K)) return gmr::score_tc_workspace_bytes(B, I, D, K);
    return gmr::align_up(simt, 256);
}

extern "C" int gmr_score_mask_topk_f32(const float* Eu, int64_t lde_u, const int64_t* users, int32_t B,
                                       const float* Ei, int64_t lde_i, const float* bias, int32_t I, int32_t D,
                                       const int64_t* mask_rowptr, const int32_t* mask_items, int32_t K,
                                       int32_t precision, int32_t* out_ids, float* out_scores, void* workspace,
                                       int64_t workspace_bytes, void* stream)
{
    GMR_REQUIRE(B >= 0 && I >= 1 && D >= 1, "gmr_score_mask_topk_f32: bad shape B=%d I=%d D=%d", B, I, D);
    GMR_REQUIRE(K >= 1 && K <= GMR_MAX_TOPK, "gmr_score_mask_topk_f32: K=%d outside [1, %d]", K, GMR_MAX_TOPK);
    GMR_REQUIRE(precision == GMR_SCORE_FP32 || precision == GMR_SCORE_TC || precision == GMR_SCORE_TC_SPLIT,
                "gmr_score_mask_topk_f32: unknown precision %d", precision);
    if (B == 0) return GMR_OK;
    GMR_REQUIRE(Eu && Ei && out_ids, "gmr_score_mask_topk_f32: null operand");
    GMR_REQUIRE(lde_u >= D && lde_i >= D, "gmr_score_mask_topk_f32: leading dimension smaller than D");
    GMR_REQUIRE(mask_rowptr == nullptr || mask_items != nullptr, "gmr_score_mask_topk_f32: mask_rowptr without mask_items");
    const int64_t need = gmr_score_topk_workspace_bytes(B, I, D, K, precision);
    if (workspace == nullptr || workspace_bytes < need) {
        gmr::set_error("gmr_score_mask_topk_f32: workspace of %lld bytes required, %lld given", (long long)need,
                       (long long)workspace_bytes);
        return GMR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == GMR_SCORE_TC) {
        if (!gmr::score_screen_supported(D, K)) {
            gmr::set_error("gmr_score_mask_topk_f32: GMR_SCORE_TC needs D %% 64 == 0, D <= 256 and K <= 256 (got D=%d, K=%d)", D, K);
            return GMR_ERR_UNSUPPORTED;
        }
        return gmr::score_topk_screen_launch(Eu, lde_u, users, B, Ei, lde_i, bias, I, D, mask_rowptr, mask_items, K,
                                             out_ids, out_scores, workspace, workspace_bytes, st);
    }
    if (precision == GMR_SCORE_TC_SPLIT) {
        if (!gmr::score_tc_supported(D, K)) {
            gmr::set_error("gmr_score_mask_topk_f32: GMR_SCORE_TC_SPLIT needs D %% 64 == 0, D <= 192 and K <= 248 (got D=%d, K=%d)", D, K);
            return GMR_ERR_UNSUPPORTED;
        }
        return gmr::score_topk_tc_launch(Eu, lde_u, users, B, Ei, lde_i, bias, I, D, mask_rowptr, mask_items, K,
                                         out_ids, out_scores, workspace, workspace_bytes, st);
    }
    return gmr::score_topk_simt_launch(Eu, lde_u, users, nullptr, B, Ei, lde_i, bias, I, D, mask_rowptr, mask_items,
                                       K, out_ids, out_scores, workspace, st, 0);
}

extern "C" int gmr_score_tc_fallback_rows(const void* workspace, int32_t B, int32_t I, int32_t D, int32_t K,
                                          int32_t precision, int32_t* count_host, void* stream)
{
    GMR_REQUIRE(workspace && count_host, "gmr_score_tc_fallback_rows: null argument");
    *count_host = 0;
    if (precision == GMR_SCORE_TC && gmr::score_screen_supported(D, K))
        return gmr::score_screen_fallback_count(workspace, B, I, D, K, count_host, (cudaStream_t)stream);
    if (precision == GMR_SCORE_TC_SPLIT && gmr::score_tc_supported(D, K))
        return gmr::score_tc_fallback_count(workspace, B, I, D, K, count_host, (cudaStream_t)stream);
    return GMR_OK;
}

extern "C" int gmr_score_tc_stats(const void* workspace, int32_t B, int32_t I, int32_t D, int32_t K,
                                  uint64_t* stats_host, void* stream)
{
    GMR_REQUIRE(workspace && stats_host, "gmr_score_tc_stats: null argument");
    for (int i = 0; i < 8; ++i) stats_host[i] = 0;
    if (!gmr::score_screen_supported(D, K)) return GMR_OK;
    return gmr::score_screen_stats(workspace, B, I, D, K, stats_host, (cudaStream_t)stream);
}

extern "C" int gmr_scores_f32(const float* Eu, int64_t lde_u, const int64_t* users, int32_t B, const float* Ei,
                              int64_t lde_i, const float* bias, int32_t I, int32_t D, float* out, int64_t ldo,
                              void* stream)
{
    GMR_REQUIRE(B >= 0 && I >= 1 && D >= 1, "gmr_scores_f32: bad shape B=%d I=%d D=%d", B, I, D);
    if (B == 0) return GMR_OK;
    GMR_REQUIRE(Eu && Ei && out, "gmr_scores_f32: null operand");
    GMR_REQUIRE(lde_u >= D && lde_i >= D && ldo >= I, "gmr_scores_f32: leading dimension too small");
    return gmr::scores_dense_launch(Eu, lde_u, users, B, Ei, lde_i, bias, I, D, out, ldo, (cudaStream_t)stream);
}

// ---- peer memory (CUDA IPC) ---------------------------------------------------------------------

static_assert(sizeof(cudaIpcMemHandle_t) == GMR_PEER_HANDLE_BYTES, "IPC handle size");

extern "C" int gmr_peer_alloc(void** ptr, int64_t bytes)
{
    GMR_REQUIRE(ptr != nullptr && bytes > 0, "gmr_peer_alloc: bad arguments");
    GMR_CHECK_CUDA(cudaMalloc(ptr, (size_t)bytes));
    return GMR_OK;
}

extern "C" int gmr_peer_free(void* ptr)
{
    if (ptr) GMR_CHECK_CUDA(cudaFree(ptr));
    return GMR_OK;
}

extern "C" int gmr_peer_export(void* ptr, uint8_t* handle_host)
{
    GMR_REQUIRE(ptr && handle_host, "gmr_peer_export: null argument");
    cudaIpcMemHandle_t h;
    GMR_CHECK_CUDA(cudaIpcGetMemHandle(&h, ptr));
    std::memcpy(handle_host, &h, sizeof(h));
    return GMR_OK;
}

extern "C" int gmr_peer_open(const uint8_t* handle_host, void** ptr)
{
    GMR_REQUIRE(ptr && handle_host, "gmr_peer_open: null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle_host, sizeof(h));
    GMR_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GMR_OK;
}

extern "C" int gmr_peer_close(void* ptr)
{
    if (ptr) GMR_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return GMR_OK;
}
