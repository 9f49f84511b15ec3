// crb_score_topk: full-rank evaluation entry point (reference: test_model_rs, model/RankingRecommender.py:198-247).
//   exact = 1  canonical fp32 on CUDA cores (score.cu: fullrank_exact_kernel)
//   exact = 0  bf16 tcgen05 tensor-core candidate pass + canonical fp32 re-scoring with a per-user certificate
//              (score_tc.cu); users whose certificate fails are re-run through the exact kernel on the GPU.
#include "score_common.cuh"

int crb_score_topk_tc(crb_handle* h, int32_t kind, const float* P, const float* Q, const float* hvec, int64_t n_items, int32_t dim,
                      const int32_t* users, const int32_t* hist_users, int64_t n_users, int32_t K, int32_t* topk_items,
                      float* topk_scores, cudaStream_t s);

extern "C" int crb_score_topk(crb_handle* h, int32_t kind, const float* P, const float* Q, const float* hvec, int64_t n_items,
                              int32_t dim, const int32_t* users, const int32_t* hist_users, int64_t n_users, int32_t K,
                              int32_t exact, int32_t* topk_items, float* topk_scores, void* stream) {
    CRB_CHECK_ARG(h && P && Q && users && topk_items, "null argument");
    CRB_CHECK_ARG(n_items > 0 && dim > 0 && n_users >= 0 && K >= 1, "sizes");
    CRB_CHECK_ARG(kind == CRB_SCORE_DOT || kind == CRB_SCORE_SQDIST || hvec, "this score kind needs hvec");
    if (!h->seen_rowptr) { crb_set_error("crb_score_topk before crb_set_history"); return CRB_ERR_STATE; }
    if (n_users == 0) return CRB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CUDA(cudaSetDevice(h->device));
    // stage host buffers (users / hist_users in, ids / scores out) through a small private allocation
    const bool u_dev = crb_is_device_ptr(users), hu_dev = !hist_users || crb_is_device_ptr(hist_users);
    const bool oi_dev = crb_is_device_ptr(topk_items), os_dev = !topk_scores || crb_is_device_ptr(topk_scores);
    int32_t *d_users = nullptr, *d_hist = nullptr, *d_items = nullptr;
    float* d_scores = nullptr;
    void* tmp = nullptr;
    int64_t bytes = 0;
    if (!u_dev) bytes += 4 * n_users;
    if (!hu_dev) bytes += 4 * n_users;
    if (!oi_dev) bytes += 4 * n_users * K;
    if (!os_dev) bytes += 4 * n_users * K;
    if (bytes) {
        CRB_CUDA(cudaMallocAsync(&tmp, bytes, s));
        char* p = (char*)tmp;
        if (!u_dev) { d_users = (int32_t*)p; p += 4 * n_users; CRB_CUDA(cudaMemcpyAsync(d_users, users, 4 * n_users, cudaMemcpyHostToDevice, s)); }
        if (!hu_dev) { d_hist = (int32_t*)p; p += 4 * n_users; CRB_CUDA(cudaMemcpyAsync(d_hist, hist_users, 4 * n_users, cudaMemcpyHostToDevice, s)); }
        if (!oi_dev) { d_items = (int32_t*)p; p += 4 * n_users * K; }
        if (!os_dev) { d_scores = (float*)p; p += 4 * n_users * K; }
    }
    const int32_t* du = u_dev ? users : d_users;
    const int32_t* dh = hist_users ? (hu_dev ? hist_users : d_hist) : nullptr;
    int32_t* di = oi_dev ? topk_items : d_items;
    float* ds = topk_scores ? (os_dev ? topk_scores : d_scores) : nullptr;
    int rc;
    if (exact) {
        rc = crb_launch_fullrank_exact(h, kind, P, Q, hvec, n_items, dim, du, dh, nullptr, n_users, K, di, ds, s);
    } else {
        rc = crb_score_topk_tc(h, kind, P, Q, hvec, n_items, dim, du, dh, n_users, K, di, ds, s);
    }
    if (rc) { if (tmp) cudaFreeAsync(tmp, s); return rc; }
    bool sync = false;
    if (!oi_dev) { CRB_CUDA(cudaMemcpyAsync(topk_items, di, 4 * n_users * K, cudaMemcpyDeviceToHost, s)); sync = true; }
    if (topk_scores && !os_dev) { CRB_CUDA(cudaMemcpyAsync(topk_scores, ds, 4 * n_users * K, cudaMemcpyDeviceToHost, s)); sync = true; }
    if (tmp) CRB_CUDA(cudaFreeAsync(tmp, s));
    if (sync) CRB_CUDA(cudaStreamSynchronize(s));
    return CRB_OK;
}

extern "C" int crb_score_topk_stats(crb_handle* h, int64_t stats[4]) {
    CRB_CHECK_ARG(h && stats, "null argument");
    for (int k = 0; k < 4; ++k) stats[k] = h->topk_stats[k];
    return CRB_OK;
}
