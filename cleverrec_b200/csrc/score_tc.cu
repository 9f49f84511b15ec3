// placeholder until the tcgen05 path lands (replaced in the next commit)
#include "score_common.cuh"
int crb_score_topk_tc(crb_handle* h, int32_t kind, const float* P, const float* Q, const float* hvec, int64_t n_items, int32_t dim,
                      const int32_t* users, const int32_t* hist_users, int64_t n_users, int32_t K, int32_t* topk_items,
                      float* topk_scores, cudaStream_t s) {
    crb_set_error("tensor-core scoring not built yet");
    return CRB_ERR_UNSUPPORTED;
}
