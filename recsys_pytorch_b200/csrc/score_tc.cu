// tcgen05 scoring path - placeholder until the tensor-core kernel lands; refuses loudly.
#include "common.cuh"
namespace b200 {
int64_t score_topk_tc_workspace(int, int, int, int) { return 0; }
int score_topk_tc(const float *, const float *, int, int, const int32_t *, int, int, const int64_t *, const int32_t *,
                  int, int32_t *, float *, void *, int64_t, cudaStream_t) {
    set_error("score_topk: B200REC_SCORE_TC is not built in this revision");
    return B200REC_EUNSUPPORTED;
}
}  // namespace b200
