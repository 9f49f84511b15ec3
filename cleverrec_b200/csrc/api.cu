// Handle, workspace and error plumbing of libcleverrec_b200.so.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void crb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* crb_last_error(void) { return g_err; }
extern "C" int crb_abi_version(void) { return CRB_ABI_VERSION; }

extern "C" int crb_create(int device, crb_handle** out) {
    CRB_CHECK_ARG(out, "out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        crb_set_error("no CUDA device available (%s); cleverrec_b200 has no CPU path", e == cudaSuccess ? "count == 0" : cudaGetErrorString(e));
        return CRB_ERR_CUDA;
    }
    CRB_CHECK_ARG(device >= 0 && device < n, "device index out of range");
    CRB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CRB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        crb_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return CRB_ERR_UNSUPPORTED;
    }
    {
        // the library's temporaries (history / preprocessing sorts, host-feed staging) come from the stream-ordered allocator; by
        // default its pool hands physical memory back at every synchronisation, so each call paid the driver's allocation again
        // (1.5 s of a 1e8-row history build): keep up to 8 GB cached
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = 8ULL << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    crb_handle* h = (crb_handle*)calloc(1, sizeof(crb_handle));
    if (!h) {
        crb_set_error("out of host memory");
        return CRB_ERR_ARG;
    }
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->loss_blocks = h->sm_count * 8;
    CRB_CUDA(cudaMalloc(&h->ctr, sizeof(crb_step_ctr)));
    CRB_CUDA(cudaMemset(h->ctr, 0, sizeof(crb_step_ctr)));
    CRB_CUDA(cudaMalloc(&h->block_loss, sizeof(double) * h->loss_blocks));
    CRB_CUDA(cudaMemset(h->block_loss, 0, sizeof(double) * h->loss_blocks));
    CRB_CUDA(cudaMalloc(&h->lrt, sizeof(float) * CRB_LRT_TABLE));
    CRB_CUDA(cudaMalloc(&h->dense_loss, sizeof(double) * 4 * h->loss_blocks));
    CRB_CUDA(cudaMalloc(&h->dyn_dev, sizeof(uint32_t) * 8));
    CRB_CUDA(cudaMemset(h->dyn_dev, 0, sizeof(uint32_t) * 8));
    h->lrt_lr = -1.0;
    *out = h;
    return CRB_OK;
}

void crb_alt_swap(crb_handle* h) {
#define CRB_SWAP(a, b) do { auto t__ = (a); (a) = (b); (b) = t__; } while (0)
    for (int k = 0; k < 3; ++k) { CRB_SWAP(h->idx[k], h->alt.idx[k]); CRB_SWAP(h->rank[k], h->alt.rank[k]); }
    for (int k = 0; k < 2; ++k) { CRB_SWAP(h->meta[k], h->alt.meta[k]); CRB_SWAP(h->meta_rows[k], h->alt.meta_rows[k]); CRB_SWAP(h->sb[k], h->alt.sb[k]); }
    CRB_SWAP(h->ctr, h->alt.ctr);
    CRB_SWAP(h->ctr_zeroed, h->alt.ctr_zeroed);
    CRB_SWAP(h->dup_rows, h->alt.dup_rows);
    CRB_SWAP(h->work, h->alt.work);
    CRB_SWAP(h->multi, h->alt.multi);
#undef CRB_SWAP
    h->alt_active ^= 1;
}

static void free_alt_ws(crb_handle* h) {
    for (int k = 0; k < 3; ++k) { cudaFree(h->alt.idx[k]); h->alt.idx[k] = nullptr; cudaFree(h->alt.rank[k]); h->alt.rank[k] = nullptr; }
    for (int k = 0; k < 2; ++k) { cudaFree(h->alt.sb[k]); h->alt.sb[k] = nullptr; }
    cudaFree(h->alt.dup_rows); h->alt.dup_rows = nullptr;
    cudaFree(h->alt.work); h->alt.work = nullptr;
    cudaFree(h->alt.multi); h->alt.multi = nullptr;
    h->alt.cap_batch = 0;
}

int crb_alt_reserve(crb_handle* h, cudaStream_t s) {
    if (h->alt_active) crb_alt_swap(h);
    if (!h->aux_stream) {
        CRB_CUDA(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
        CRB_CUDA(cudaEventCreateWithFlags(&h->ev_entry, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) {
            CRB_CUDA(cudaEventCreateWithFlags(&h->ev_prep[k], cudaEventDisableTiming));
            CRB_CUDA(cudaEventCreateWithFlags(&h->ev_done[k], cudaEventDisableTiming));
        }
        CRB_CUDA(cudaMalloc(&h->alt.ctr, sizeof(crb_step_ctr)));
        CRB_CUDA(cudaMemset(h->alt.ctr, 0, sizeof(crb_step_ctr)));
    }
    if (h->alt.cap_batch != h->cap_batch) {
        CRB_CUDA(cudaStreamSynchronize(s));
        CRB_CUDA(cudaStreamSynchronize(h->aux_stream));
        free_alt_ws(h);
        const int64_t nb = h->cap_batch, occ = 3 * nb;
        for (int k = 0; k < 3; ++k) {
            CRB_CUDA(cudaMalloc(&h->alt.idx[k], sizeof(int32_t) * nb));
            CRB_CUDA(cudaMalloc(&h->alt.rank[k], sizeof(uint32_t) * nb));
        }
        for (int k = 0; k < 2; ++k) CRB_CUDA(cudaMalloc(&h->alt.sb[k], sizeof(int32_t) * nb));
        CRB_CUDA(cudaMalloc(&h->alt.dup_rows, sizeof(crb_dup_row) * (occ / 2 + 1)));
        CRB_CUDA(cudaMalloc(&h->alt.work, sizeof(crb_work) * (occ / 2 + occ / CRB_DUP_CHUNK + 2)));
        CRB_CUDA(cudaMalloc(&h->alt.multi, sizeof(unsigned int) * (occ / CRB_DUP_CHUNK + 2)));
        h->alt.cap_batch = nb;
    }
    for (int w = 0; w < 2; ++w) {
        if (h->alt.meta_rows[w] == h->meta_rows[w]) continue;
        CRB_CUDA(cudaStreamSynchronize(s));
        CRB_CUDA(cudaStreamSynchronize(h->aux_stream));
        cudaFree(h->alt.meta[w]);
        h->alt.meta[w] = nullptr;
        h->alt.meta_rows[w] = 0;
        if (h->meta_rows[w] > 0) {
            CRB_CUDA(cudaMalloc(&h->alt.meta[w], sizeof(unsigned long long) * h->meta_rows[w]));
            CRB_CUDA(cudaMemsetAsync(h->alt.meta[w], 0, sizeof(unsigned long long) * h->meta_rows[w], s));
            h->alt.meta_rows[w] = h->meta_rows[w];
        }
    }
    return CRB_OK;
}

static void free_ws(crb_handle* h) {
    h->ws_generation++;
    if (h->aux_stream) cudaStreamSynchronize(h->aux_stream);
    if (h->alt_active) crb_alt_swap(h);
    free_alt_ws(h);
    h->prep_valid = 0;
    cudaFree(h->dup_src); h->dup_src = nullptr;
    for (int k = 0; k < 4; ++k) { cudaFree(h->idx[k]); h->idx[k] = nullptr; }
    for (int k = 0; k < 3; ++k) { cudaFree(h->rank[k]); h->rank[k] = nullptr; }
    for (int k = 0; k < 2; ++k) { cudaFree(h->sb[k]); h->sb[k] = nullptr; }
    cudaFree(h->stage); h->stage = nullptr; h->stage_rows = 0; h->stage_dim = 0;
    cudaFree(h->yv); h->yv = nullptr;
    cudaFree(h->dup_grad); h->dup_grad = nullptr;
    cudaFree(h->dup_t); h->dup_t = nullptr;
    cudaFree(h->dup_rows); h->dup_rows = nullptr;
    cudaFree(h->work); h->work = nullptr;
    cudaFree(h->multi); h->multi = nullptr;
    cudaFree(h->partial); h->partial = nullptr;
    h->cap_batch = 0;
    h->cap_dim = 0;
}

extern "C" int crb_destroy(crb_handle* h) {
    if (!h) return CRB_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    free_ws(h);
    cudaFree(h->meta[0]);
    cudaFree(h->meta[1]);
    cudaFree(h->alt.meta[0]);
    cudaFree(h->alt.meta[1]);
    cudaFree(h->alt.ctr);
    if (h->aux_stream) {
        cudaStreamDestroy(h->aux_stream);
        cudaEventDestroy(h->ev_entry);
        for (int k = 0; k < 2; ++k) { cudaEventDestroy(h->ev_prep[k]); cudaEventDestroy(h->ev_done[k]); }
    }
    if (h->epoch_graph) cudaGraphExecDestroy((cudaGraphExec_t)h->epoch_graph);
    cudaFree(h->dyn_dev);
    cudaFree(h->ctr);
    cudaFree(h->block_loss);
    cudaFree(h->loss_dev);
    cudaFree(h->lrt);
    cudaFree(h->dense_loss);
    cudaFree(h->np_state); cudaFree(h->np_raw); cudaFree(h->np_scratch); cudaFree(h->np_sort_tmp); cudaFree(h->np_result);
    cudaFree(h->dense_grad);
    cudaFree(h->bloom);
    cudaFree(h->eval_ws);
    for (int t = 0; t < CRB_PROF_TAGS; ++t)
        if (h->prof_ev[t]) { for (int k = 0; k < 2 * CRB_PROF_CAP; ++k) cudaEventDestroy(h->prof_ev[t][k]); free(h->prof_ev[t]); }
    free(h);
    return CRB_OK;
}

extern "C" int64_t crb_launch_count(crb_handle* h) { return h ? h->launches : -1; }

// ---------------------------------------------------------------------------------------------- seen-item Bloom filters
__global__ void __launch_bounds__(256) bloom_build_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ cols, int64_t n_users,
                                                          uint32_t* bloom, int shift, int64_t stride, int exact) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n_users; u += n_warps) {
        const int64_t lo = rowptr[u], hi = rowptr[u + 1];
        uint32_t* mine = bloom + u * stride;
        for (int64_t p = lo + lane; p < hi; p += 32) {
            const uint32_t bit = exact ? (uint32_t)cols[p] : crb_bloom_bit((uint32_t)cols[p], shift);
            atomicOr(mine + (bit >> 5), 1u << (bit & 31));
        }
    }
}

// (Re)builds the per-user Bloom filters for the history just installed.  ~8 bits per seen entry on average, 32..2048 bits per user;
// skipped (bloom = NULL, exact search only) when it would not fit in a quarter of the free device memory or CRB_NO_BLOOM is set.
// Small catalogues (users x items bits <= 64 MB: every dataset the reference ships) get the EXACT bitmap instead (bit = item id): a
// set bit then rejects the candidate without the dependent binary search over the history, which at ml-1m's density (165 of 3706
// items seen on average, filters 8-40 % full) one lane of almost every warp had to take -- the sampler was the longest kernel of
// the step there.  Same accept / reject decisions either way (CRB_NO_EXACT_BITMAP=1 keeps the hashed filter: A/B and the test).
int crb_bloom_build(crb_handle* h, cudaStream_t s) {
    CRB_CUDA(cudaSetDevice(h->device));
    int64_t n_seen = 0;
    CRB_CUDA(cudaMemcpyAsync(&n_seen, h->seen_rowptr + h->n_users, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CRB_CUDA(cudaStreamSynchronize(s));
    const bool off = getenv("CRB_NO_BLOOM") != nullptr || n_seen <= 0 || !h->seen_cols;
    int shift = 0;
    if (!off) {
        const double bits = 8.0 * (double)n_seen / (double)h->n_users;
        while (shift < 6 && (double)(32u << shift) < bits) ++shift;
        size_t free_b = 0, total_b = 0;
        CRB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        while (shift > 0 && (size_t)(h->n_users << shift) * 4 > free_b / 4) --shift;
        if ((size_t)(h->n_users << shift) * 4 > free_b / 4) shift = -1;
    }
    if (off || shift < 0) {
        if (h->bloom) { CRB_CUDA(cudaStreamSynchronize(s)); cudaFree(h->bloom); }
        h->bloom = nullptr; h->bloom_words = 0; h->bloom_shift = 0; h->bloom_stride = 0; h->bloom_exact = 0;
        return CRB_OK;
    }
    const int64_t exact_stride = (h->n_items + 31) / 32;
    const bool exact = getenv("CRB_NO_EXACT_BITMAP") == nullptr && h->n_users * exact_stride * 4 <= ((int64_t)64 << 20);
    const int64_t stride = exact ? exact_stride : ((int64_t)1 << shift);
    const int64_t words = h->n_users * stride;
    if (words > h->bloom_words) {
        CRB_CUDA(cudaStreamSynchronize(s));
        cudaFree(h->bloom);
        h->bloom = nullptr; h->bloom_words = 0;
        CRB_CUDA(cudaMalloc(&h->bloom, sizeof(uint32_t) * words));
        h->bloom_words = words;
    }
    h->bloom_shift = shift;
    h->bloom_stride = stride;
    h->bloom_exact = exact ? 1 : 0;
    CRB_CUDA(cudaMemsetAsync(h->bloom, 0, sizeof(uint32_t) * words, s));
    int64_t grid = (h->n_users + 7) / 8;
    if (grid > (int64_t)h->sm_count * 32) grid = (int64_t)h->sm_count * 32;
    bloom_build_kernel<<<(int)grid, 256, 0, s>>>(h->seen_rowptr, h->seen_cols, h->n_users, h->bloom, shift, stride, exact ? 1 : 0);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

extern "C" int crb_set_history(crb_handle* h, int64_t n_users, int64_t n_items, int64_t n_pos, const int32_t* pos_user,
                               const int32_t* pos_item, const int64_t* seen_rowptr, const int32_t* seen_cols, void* stream) {
    CRB_CHECK_ARG(h, "null handle");
    CRB_CHECK_ARG(n_users > 0 && n_items > 0 && n_pos >= 0, "sizes");
    CRB_CHECK_ARG(n_users < 0x7fffffffLL && n_items < 0x7fffffffLL, "row ids are int32");
    CRB_CHECK_ARG(crb_is_device_ptr(seen_rowptr), "seen_rowptr must be a device pointer");
    CRB_CHECK_ARG(n_pos == 0 || (crb_is_device_ptr(pos_user) && crb_is_device_ptr(pos_item)), "pos_user/pos_item must be device pointers");
    h->n_users = n_users;
    h->n_items = n_items;
    h->n_pos = n_pos;
    h->pos_user = pos_user;
    h->pos_item = pos_item;
    h->seen_rowptr = seen_rowptr;
    h->seen_cols = seen_cols;
    h->list_start = nullptr;
    h->list_len = nullptr;
    return crb_bloom_build(h, (cudaStream_t)stream);
}

int crb_meta_reserve(crb_handle* h, int which, int64_t rows, cudaStream_t s) {
    if (h->alt_active) crb_alt_swap(h);
    if (h->meta_rows[which] >= rows) return CRB_OK;
    CRB_CUDA(cudaStreamSynchronize(s));
    h->ws_generation++;
    cudaFree(h->meta[which]);
    h->meta[which] = nullptr;
    h->meta_rows[which] = 0;
    CRB_CUDA(cudaMalloc(&h->meta[which], sizeof(unsigned long long) * rows));
    CRB_CUDA(cudaMemsetAsync(h->meta[which], 0, sizeof(unsigned long long) * rows, s));
    h->meta_rows[which] = rows;
    return CRB_OK;
}

int crb_ws_reserve(crb_handle* h, int64_t batch, int32_t dim, int64_t steps, cudaStream_t s) {
    if (steps > h->cap_steps) {
        CRB_CUDA(cudaStreamSynchronize(s));
        h->ws_generation++;
        cudaFree(h->loss_dev);
        h->loss_dev = nullptr;
        CRB_CUDA(cudaMalloc(&h->loss_dev, sizeof(double) * steps));
        h->cap_steps = steps;
    }
    if (batch <= h->cap_batch && dim <= h->cap_dim) return CRB_OK;
    CRB_CUDA(cudaStreamSynchronize(s));
    int64_t nb = batch > h->cap_batch ? batch : h->cap_batch;
    int32_t nd = dim > h->cap_dim ? dim : h->cap_dim;
    free_ws(h);
    const int64_t occ = 3 * nb;
    for (int k = 0; k < 4; ++k) CRB_CUDA(cudaMalloc(&h->idx[k], sizeof(int32_t) * nb));
    for (int k = 0; k < 3; ++k) CRB_CUDA(cudaMalloc(&h->rank[k], sizeof(uint32_t) * nb));
    for (int k = 0; k < 2; ++k) CRB_CUDA(cudaMalloc(&h->sb[k], sizeof(int32_t) * nb));
    CRB_CUDA(cudaMalloc(&h->yv, sizeof(float) * nb));
    CRB_CUDA(cudaMalloc(&h->dup_grad, sizeof(float) * occ * nd));
    CRB_CUDA(cudaMalloc(&h->dup_t, sizeof(uint32_t) * occ));
    CRB_CUDA(cudaMalloc(&h->dup_src, sizeof(uint32_t) * occ));
    CRB_CUDA(cudaMalloc(&h->dup_rows, sizeof(crb_dup_row) * (occ / 2 + 1)));
    CRB_CUDA(cudaMalloc(&h->work, sizeof(crb_work) * (occ / 2 + occ / CRB_DUP_CHUNK + 2)));
    CRB_CUDA(cudaMalloc(&h->multi, sizeof(unsigned int) * (occ / CRB_DUP_CHUNK + 2)));
    h->cap_partial = 2 * (occ / CRB_DUP_CHUNK) + 2;
    CRB_CUDA(cudaMalloc(&h->partial, sizeof(float) * h->cap_partial * nd));
    h->cap_batch = nb;
    h->cap_dim = nd;
    return CRB_OK;
}

extern "C" int crb_eval_cache_invalidate(crb_handle* h) {
    CRB_CHECK_ARG(h, "null handle");
    h->evq_valid = 0;
    return CRB_OK;
}

int crb_eval_ws_reserve(crb_handle* h, int64_t bytes) {
    if (bytes <= h->eval_ws_bytes) return CRB_OK;
    h->evq_valid = 0;
    CRB_CUDA(cudaDeviceSynchronize());
    cudaFree(h->eval_ws);
    h->eval_ws = nullptr;
    h->eval_ws_bytes = 0;
    CRB_CUDA(cudaMalloc(&h->eval_ws, bytes));
    h->eval_ws_bytes = bytes;
    return CRB_OK;
}

// lr_t(s) = lr * sqrt(1 - beta2^s) / (1 - beta1^s) evaluated in double like TF / the torch restatement, rounded to fp32
int crb_lrt_prepare(crb_handle* h, const crb_opt* opt, cudaStream_t s) {
    if (h->lrt_lr == opt->lr && h->lrt_b1 == opt->beta1 && h->lrt_b2 == opt->beta2) return CRB_OK;
    float* host = (float*)malloc(sizeof(float) * CRB_LRT_TABLE);
    if (!host) { crb_set_error("out of host memory"); return CRB_ERR_ARG; }
    const double lr = opt->lr, b1 = opt->beta1, b2 = opt->beta2;
    host[0] = 0.f;
    for (int t = 1; t < CRB_LRT_TABLE; ++t) host[t] = (float)(lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t)));
    cudaError_t e = cudaMemcpyAsync(h->lrt, host, sizeof(float) * CRB_LRT_TABLE, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    free(host);
    CRB_CUDA(e);
    h->lrt_lr = opt->lr;
    h->lrt_b1 = opt->beta1;
    h->lrt_b2 = opt->beta2;
    return CRB_OK;
}

// ------------------------------------------------------------------------------------------------ profiling hook
static int prof_drain(crb_handle* h, int tag) {
    for (int k = 0; k < h->prof_n[tag]; ++k) {
        float ms = 0.f;
        CRB_CUDA(cudaEventSynchronize(h->prof_ev[tag][2 * k + 1]));
        CRB_CUDA(cudaEventElapsedTime(&ms, h->prof_ev[tag][2 * k], h->prof_ev[tag][2 * k + 1]));
        h->prof_ms[tag] += (double)ms;
        h->prof_launches[tag]++;
    }
    h->prof_n[tag] = 0;
    return CRB_OK;
}

int crb_prof_begin(crb_handle* h, cudaStream_t s, int tag) {
    if (!h->prof_on) return CRB_OK;
    if (h->prof_n[tag] == CRB_PROF_CAP) { int rc = prof_drain(h, tag); if (rc) return rc; }
    CRB_CUDA(cudaEventRecord(h->prof_ev[tag][2 * h->prof_n[tag]], s));
    return CRB_OK;
}

int crb_prof_end(crb_handle* h, cudaStream_t s, int tag) {
    if (!h->prof_on) return CRB_OK;
    CRB_CUDA(cudaEventRecord(h->prof_ev[tag][2 * h->prof_n[tag] + 1], s));
    h->prof_n[tag]++;
    return CRB_OK;
}

extern "C" int crb_profile_enable(crb_handle* h, int32_t on) {
    CRB_CHECK_ARG(h, "null handle");
    if (on && !h->prof_ev[0]) {
        for (int t = 0; t < CRB_PROF_TAGS; ++t) {
            h->prof_ev[t] = (cudaEvent_t*)calloc(2 * CRB_PROF_CAP, sizeof(cudaEvent_t));
            for (int k = 0; k < 2 * CRB_PROF_CAP; ++k) CRB_CUDA(cudaEventCreate(&h->prof_ev[t][k]));
        }
    }
    if (!on && h->prof_on)
        for (int t = 0; t < CRB_PROF_TAGS; ++t) { int rc = prof_drain(h, t); if (rc) return rc; }
    h->prof_on = on ? 1 : 0;
    return CRB_OK;
}

extern "C" int crb_profile_read_tag(crb_handle* h, int32_t tag, double* ms, int64_t* n_launches) {
    CRB_CHECK_ARG(h && ms && n_launches && tag >= 0 && tag < CRB_PROF_TAGS, "bad argument");
    if (h->prof_ev[0]) { int rc = prof_drain(h, tag); if (rc) return rc; }
    *ms = h->prof_ms[tag];
    *n_launches = h->prof_launches[tag];
    h->prof_ms[tag] = 0.0;
    h->prof_launches[tag] = 0;
    return CRB_OK;
}

extern "C" int crb_profile_read(crb_handle* h, double* step_kernel_ms, int64_t* n_launches) {
    return crb_profile_read_tag(h, 0, step_kernel_ms, n_launches);
}
