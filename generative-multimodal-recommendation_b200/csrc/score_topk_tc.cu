// score_topk_tc.cu -- K2/K3 (tensor-core path) -- placeholder until the tcgen05 kernel lands.
#include "common.cuh"

namespace gmr {

bool score_tc_supported(int32_t, int32_t) { return false; }
int64_t score_tc_workspace_bytes(int32_t, int32_t, int32_t, int32_t) { return 0; }
int score_topk_tc_launch(const float*, int64_t, const int64_t*, int32_t, const float*, int64_t, const float*, int32_t,
                         int32_t, const int64_t*, const int32_t*, int32_t, int32_t*, float*, void*, int64_t,
                         cudaStream_t)
{
    set_error("tensor-core scoring path not built");
    return GMR_ERR_UNSUPPORTED;
}

}  // namespace gmr
