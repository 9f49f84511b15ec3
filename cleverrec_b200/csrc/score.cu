// Evaluation kernels in canonical fp32 arithmetic (reference: each model's `_predict` + the ranking part of
// model/RankingRecommender.py:198-299).  "Canonical" = one sequential fp32 fma chain over k = 0..d-1 per
// (user, item) pair, the definition oracle/crb_oracle.c restates, so that ranks and top-K ids are bit-exact.
//   score_pairs_kernel      sess.run(pre_scores, {u_idx, i_idx})        test_model_loo :257-278
//   topk_segments_kernel    np.argsort(-scores_u)[:K] per user           test_model_loo :281-288
//   fullrank_exact_kernel   matmul + argsort + seen filter + first K     test_model_rs  :203-240
#include <cuda_pipeline.h>
#include <stdlib.h>

#include "score_common.cuh"

template <int KIND>
__global__ void __launch_bounds__(256) score_pairs_kernel(const float* __restrict__ P, const float* __restrict__ Q,
                                                          const float* __restrict__ hvec, int dim, const int32_t* __restrict__ u,
                                                          const int32_t* __restrict__ it, int64_t n, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const int32_t item = it[k];
        out[k] = canonical_score<KIND>(P + (int64_t)u[k] * dim, Q + (int64_t)item * dim, hvec, item, dim);
    }
}

// The same scores for large calls (test_model_loo over every test user: (1 + neg_samples) pairs per user).  A thread still owns one pair
// and runs the canonical sequential fma chain -- that is what makes the result bit-identical to score_pairs_kernel and to the oracle --
// but the rows are no longer read by the thread that consumes them (32 scattered 16-byte reads per warp instruction, the chain stalling
// on each): a warp stages a 64-column chunk of its 32 item rows (and of the user row, once if the warp's pairs share the user) in
// shared memory with coalesced 128-byte reads, all issued before the first use, then every thread walks its own row out of shared
// memory (row stride 65: conflict free).  Bound: HBM, (1 + neg_samples) * 4 * d bytes per user.
#define SP_CH 64
#define SP_WARPS 4
// rows `row` (one per lane) of `table`, columns [kc, kc+len): row r lands at dst[r * (SP_CH + 1) ...]; lane l copies columns l and l + 32
// of every row (128-byte coalesced) with cp.async (LDGSTS): global -> shared without passing through registers, so all 64 copies of
// the chunk are in flight at once (written as register loads + stores, ptxas re-schedules them to ~6 in flight to save registers)
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ table, int32_t row, int dim, int kc, int len, int lane) {
    const bool in0 = lane < len, in1 = lane + 32 < len;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const float* src = table + (int64_t)__shfl_sync(0xffffffffu, row, r) * dim + kc;
        if (in0) __pipeline_memcpy_async(dst + r * (SP_CH + 1) + lane, src + lane, 4);
        if (in1) __pipeline_memcpy_async(dst + r * (SP_CH + 1) + lane + 32, src + lane + 32, 4);
    }
    __pipeline_commit();
}

template <int KIND>
__global__ void __launch_bounds__(SP_WARPS * 32) score_pairs_tiled_kernel(const float* __restrict__ P, const float* __restrict__ Q,
                                                                          const float* __restrict__ hvec, int dim, const int32_t* __restrict__ u,
                                                                          const int32_t* __restrict__ it, int64_t n, float* __restrict__ out) {
    extern __shared__ float sp_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sQ = sp_sm + warp * (2 * 32 * (SP_CH + 1));
    float* sP = sQ + 32 * (SP_CH + 1);
    const int64_t n_batches = (n + 31) / 32;
    const int64_t gw = (int64_t)blockIdx.x * SP_WARPS + warp, nw = (int64_t)gridDim.x * SP_WARPS;
    for (int64_t b = gw; b < n_batches; b += nw) {
        const int64_t k = b * 32 + lane;
        const bool valid = k < n;
        const int64_t kk = valid ? k : n - 1;
        const int32_t item = it[kk], usr = u[kk];
        const int32_t usr0 = __shfl_sync(0xffffffffu, usr, 0);
        const bool uniform = __all_sync(0xffffffffu, usr == usr0);
        float acc = 0.f;
        for (int kc = 0; kc < dim; kc += SP_CH) {
            const int len = dim - kc < SP_CH ? dim - kc : SP_CH;
            stage_rows(sQ, Q, item, dim, kc, len, lane);
            if (uniform) {
                for (int c = lane; c < len; c += 32) sP[c] = P[(int64_t)usr0 * dim + kc + c];
            } else {
                stage_rows(sP, P, usr, dim, kc, len, lane);
            }
            __pipeline_wait_prior(0);
            __syncwarp();
            const float* q = sQ + lane * (SP_CH + 1);
            const float* p = uniform ? sP : sP + lane * (SP_CH + 1);
            for (int c = 0; c < len; ++c) {
                const float a = p[c], bq = q[c];
                if (KIND == CRB_SCORE_DOT || KIND == CRB_SCORE_DOT_BIAS) acc = fmaf(a, bq, acc);
                else if (KIND == CRB_SCORE_GMF) acc = fmaf(__fmul_rn(a, bq), __ldg(hvec + kc + c), acc);
                else { const float dd = __fsub_rn(a, bq); acc = fmaf(dd, dd, acc); }
            }
            __syncwarp();
        }
        if (KIND == CRB_SCORE_DOT_BIAS) acc = __fadd_rn(acc, hvec[item]);
        if (valid) out[k] = acc;
    }
}

template <int KIND>
static int launch_score_pairs(crb_handle* h, const float* P, const float* Q, const float* hvec, int dim, const int32_t* u, const int32_t* it,
                              int64_t n, float* out, cudaStream_t s) {
    if (n >= 8192 && !getenv("CRB_SCORE_PAIRS_SIMPLE")) {
        const size_t smem = sizeof(float) * SP_WARPS * 2 * 32 * (SP_CH + 1);
        CRB_CUDA(cudaFuncSetAttribute(score_pairs_tiled_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t grid = (n + 32 * SP_WARPS - 1) / (32 * SP_WARPS);
        if (grid > (int64_t)h->sm_count * 3) grid = (int64_t)h->sm_count * 3;
        score_pairs_tiled_kernel<KIND><<<(int)grid, SP_WARPS * 32, smem, s>>>(P, Q, hvec, dim, u, it, n, out);
    } else {
        int64_t grid = (n + 255) / 256;
        if (grid > (int64_t)h->sm_count * 8) grid = (int64_t)h->sm_count * 8;
        score_pairs_kernel<KIND><<<(int)(grid < 1 ? 1 : grid), 256, 0, s>>>(P, Q, hvec, dim, u, it, n, out);
    }
    return CRB_OK;
}

// One warp per user segment: K rounds of arg-best over the remaining candidates.
__global__ void __launch_bounds__(256) topk_segments_kernel(const float* __restrict__ scores, const int64_t* __restrict__ offsets,
                                                            int64_t n_users, int K, int ascending, int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t usr = warp; usr < n_users; usr += n_warps) {
        const int64_t lo = offsets[usr], hi = offsets[usr + 1];
        unsigned long long prev = ~0ULL;  // keys are strictly decreasing from round to round
        for (int r = 0; r < K; ++r) {
            unsigned long long best = 0ULL;
            for (int64_t p = lo + lane; p < hi; p += 32) {
                const unsigned long long key = rank_key(scores[p], (uint32_t)(p - lo), ascending);
                if (key < prev && key > best) best = key;
            }
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other > best ? other : best;
            }
            if (lane == 0) out[usr * K + r] = best ? (int32_t)key_index(best) : -1;
            if (!best) {  // fewer than K candidates
                for (int q = r + 1 + lane; q < K; q += 32) out[usr * K + q] = -1;
                break;
            }
            prev = best;
        }
    }
}

// ---------------------------------------------------------------------------------------- exact full-rank top-K
// One CTA per user.  Threads stride over the catalogue, compute the canonical score and push (key) into a shared
// candidate buffer when it beats the CTA's current K-th key; a full buffer is compacted by a bitonic sort.
#define FR_CAP 1024
template <int KIND>
__global__ void __launch_bounds__(256) fullrank_exact_kernel(const float* __restrict__ P, const float* __restrict__ Q,
                                                             const float* __restrict__ hvec, int64_t n_items, int dim,
                                                             const int32_t* __restrict__ users, const int32_t* __restrict__ hist_users,
                                                             const int32_t* __restrict__ todo, int64_t n_todo,
                                                             const int64_t* __restrict__ seen_rowptr, const int32_t* __restrict__ seen_cols,
                                                             int K, int32_t* __restrict__ out_items, float* __restrict__ out_scores,
                                                             int n_chunks, unsigned long long* __restrict__ part) {
    // n_chunks > 1 (few users, e.g. the tensor-core path's uncertified ones): the catalogue is cut into n_chunks ranges, one CTA
    // per (user, range) writes the range's best K keys to `part`, and fullrank_merge_kernel takes the best K of those -- the same
    // ids as one sweep, because ranking keys are a strict total order.
    extern __shared__ float s_user[];  // dim floats
    __shared__ unsigned long long buf[FR_CAP];
    __shared__ int s_count;
    __shared__ unsigned long long s_thr;
    constexpr int ASC = KIND == CRB_SCORE_SQDIST ? 1 : 0;
    const int64_t chunk_len = (n_items + n_chunks - 1) / n_chunks;
    for (int64_t w = blockIdx.x; w < n_todo * n_chunks; w += gridDim.x) {
        const int64_t k = todo ? todo[w / n_chunks] : w / n_chunks;
        const int64_t item_lo = (w % n_chunks) * chunk_len, item_hi = min(n_items, item_lo + chunk_len);
        const int32_t urow = users[k];
        const int32_t hu = hist_users ? hist_users[k] : urow;
        __syncthreads();
        for (int c = threadIdx.x; c < dim; c += blockDim.x) s_user[c] = P[(int64_t)urow * dim + c];
        if (threadIdx.x == 0) { s_count = 0; s_thr = 0ULL; }
        __syncthreads();
        const int64_t h_lo = hu >= 0 ? seen_rowptr[hu] : 0, h_hi = hu >= 0 ? seen_rowptr[hu + 1] : 0;
        for (int64_t base = item_lo; base < item_hi; base += blockDim.x) {
            const int64_t item = base + threadIdx.x;
            unsigned long long key = 0ULL;
            if (item < item_hi) {
                const float sc = canonical_score<KIND>(s_user, Q + item * dim, hvec, (int32_t)item, dim);
                key = rank_key(sc, (uint32_t)item, ASC);
                if (key <= s_thr) key = 0ULL;
                else if (sorted_contains(seen_cols, h_lo, h_hi, (int32_t)item)) key = 0ULL;
            }
            // the buffer may overflow inside one sweep of 256 items: compact first when fewer than 256 slots remain
            // (the vote makes the decision uniform: the last thread to arrive has seen every push of the previous sweep)
            if (__syncthreads_or(s_count > FR_CAP - 256)) {
                block_sort_desc<FR_CAP>(buf, s_count);
                if (threadIdx.x == 0) {
                    if (s_count > K) s_count = K;
                    if (s_count == K) s_thr = buf[K - 1];
                }
                __syncthreads();
            }
            if (key) buf[atomicAdd(&s_count, 1)] = key;
        }
        __syncthreads();
        block_sort_desc<FR_CAP>(buf, s_count);
        for (int r = threadIdx.x; r < K; r += blockDim.x) {
            const bool ok = r < s_count;
            if (n_chunks > 1) { part[w * K + r] = ok ? buf[r] : 0ULL; continue; }
            out_items[k * K + r] = ok ? (int32_t)key_index(buf[r]) : -1;
            if (out_scores) out_scores[k * K + r] = ok ? key_score(buf[r], ASC) : 0.f;
        }
    }
}

// one warp per user: K rounds of arg-best over the user's n_chunks * K partial keys
__global__ void __launch_bounds__(32) fullrank_merge_kernel(const unsigned long long* __restrict__ part, const int32_t* __restrict__ todo,
                                                            int n_chunks, int K, int ascending, int32_t* __restrict__ out_items,
                                                            float* __restrict__ out_scores) {
    const int lane = threadIdx.x;
    const int64_t w = blockIdx.x, k = todo ? todo[w] : w;
    const unsigned long long* keys = part + w * n_chunks * K;
    const int n = n_chunks * K;
    unsigned long long prev = ~0ULL;
    for (int r = 0; r < K; ++r) {
        unsigned long long best = 0ULL;
        for (int q = lane; q < n; q += 32) {
            const unsigned long long key = keys[q];
            if (key < prev && key > best) best = key;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) {
            out_items[k * K + r] = best ? (int32_t)key_index(best) : -1;
            if (out_scores) out_scores[k * K + r] = best ? key_score(best, ascending) : 0.f;
        }
        prev = best ? best : 0ULL;
    }
}

static int grid_for(crb_handle* h, int64_t n, int per_block) {
    int64_t b = (n + per_block - 1) / per_block;
    int64_t cap = (int64_t)h->sm_count * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// host<->device staging helper for the evaluation entry points
struct Stage {
    crb_handle* h;
    cudaStream_t s;
    char* base;
    int64_t used, cap;
    void* take(int64_t bytes) {
        bytes = (bytes + 255) & ~(int64_t)255;
        if (used + bytes > cap) return nullptr;
        void* p = base + used;
        used += bytes;
        return p;
    }
};

template <typename T>
static int stage_in(Stage& st, const T* src, int64_t n, const T** out) {
    if (!src) { *out = nullptr; return CRB_OK; }
    if (crb_is_device_ptr(src)) { *out = src; return CRB_OK; }
    void* d = st.take(sizeof(T) * n);
    if (!d) { crb_set_error("evaluation workspace too small"); return CRB_ERR_ARG; }
    CRB_CUDA(cudaMemcpyAsync(d, src, sizeof(T) * n, cudaMemcpyHostToDevice, st.s));
    *out = (const T*)d;
    return CRB_OK;
}

template <typename T>
static int stage_out(Stage& st, T* dst, int64_t n, T** dev) {
    if (!dst) { *dev = nullptr; return CRB_OK; }
    if (crb_is_device_ptr(dst)) { *dev = dst; return CRB_OK; }
    void* d = st.take(sizeof(T) * n);
    if (!d) { crb_set_error("evaluation workspace too small"); return CRB_ERR_ARG; }
    *dev = (T*)d;
    return CRB_OK;
}

template <typename T>
static int stage_back(Stage& st, T* dst, const T* dev, int64_t n, bool* need_sync) {
    if (!dst || dst == dev) return CRB_OK;
    CRB_CUDA(cudaMemcpyAsync(dst, dev, sizeof(T) * n, cudaMemcpyDeviceToHost, st.s));
    *need_sync = true;
    return CRB_OK;
}

extern "C" int crb_score_pairs(crb_handle* h, int32_t kind, const float* P, const float* Q, const float* hvec, int32_t dim,
                               const int32_t* u, const int32_t* i, int64_t n, float* scores, void* stream) {
    CRB_CHECK_ARG(h && P && Q && u && i && scores, "null argument");
    CRB_CHECK_ARG(dim > 0 && n >= 0, "sizes");
    CRB_CHECK_ARG(kind == CRB_SCORE_DOT || kind == CRB_SCORE_SQDIST || hvec, "this score kind needs hvec");
    if (n == 0) return CRB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CUDA(cudaSetDevice(h->device));
    int rc = crb_eval_ws_reserve(h, 12 * n + 4096);
    if (rc) return rc;
    h->evq_valid = 0;   // the staging area below overlays the cached bf16 item table
    Stage st = {h, s, (char*)h->eval_ws, 0, h->eval_ws_bytes};
    const int32_t *du, *di;
    float* ds;
    if ((rc = stage_in(st, u, n, &du))) return rc;
    if ((rc = stage_in(st, i, n, &di))) return rc;
    if ((rc = stage_out(st, scores, n, &ds))) return rc;
    switch (kind) {
        case CRB_SCORE_DOT: rc = launch_score_pairs<CRB_SCORE_DOT>(h, P, Q, hvec, dim, du, di, n, ds, s); break;
        case CRB_SCORE_GMF: rc = launch_score_pairs<CRB_SCORE_GMF>(h, P, Q, hvec, dim, du, di, n, ds, s); break;
        case CRB_SCORE_SQDIST: rc = launch_score_pairs<CRB_SCORE_SQDIST>(h, P, Q, hvec, dim, du, di, n, ds, s); break;
        case CRB_SCORE_DOT_BIAS: rc = launch_score_pairs<CRB_SCORE_DOT_BIAS>(h, P, Q, hvec, dim, du, di, n, ds, s); break;
        default: crb_set_error("unknown score kind %d", kind); return CRB_ERR_ARG;
    }
    if (rc) return rc;
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    bool sync = false;
    if ((rc = stage_back(st, scores, ds, n, &sync))) return rc;
    if (sync) CRB_CUDA(cudaStreamSynchronize(s));
    return CRB_OK;
}

extern "C" int crb_topk_segments(crb_handle* h, const float* scores, const int64_t* offsets, int64_t n_users, int32_t K,
                                 int32_t ascending, int32_t* topk_pos, void* stream) {
    CRB_CHECK_ARG(h && scores && offsets && topk_pos, "null argument");
    CRB_CHECK_ARG(K >= 1 && n_users >= 0, "sizes");
    if (n_users == 0) return CRB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CUDA(cudaSetDevice(h->device));
    int64_t total = 0;
    const bool off_dev = crb_is_device_ptr(offsets);
    if (off_dev) {
        CRB_CUDA(cudaMemcpyAsync(&total, offsets + n_users, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        CRB_CUDA(cudaStreamSynchronize(s));
    } else {
        total = offsets[n_users];
    }
    int rc = crb_eval_ws_reserve(h, 4 * total + 8 * (n_users + 1) + 4 * n_users * K + 4096);
    if (rc) return rc;
    h->evq_valid = 0;
    Stage st = {h, s, (char*)h->eval_ws, 0, h->eval_ws_bytes};
    const float* dsc;
    const int64_t* doff;
    int32_t* dout;
    if ((rc = stage_in(st, scores, total, &dsc))) return rc;
    if ((rc = stage_in(st, offsets, n_users + 1, &doff))) return rc;
    if ((rc = stage_out(st, topk_pos, n_users * K, &dout))) return rc;
    topk_segments_kernel<<<grid_for(h, n_users, 8), 256, 0, s>>>(dsc, doff, n_users, K, ascending, dout);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    bool sync = false;
    if ((rc = stage_back(st, topk_pos, dout, n_users * K, &sync))) return rc;
    if (sync) CRB_CUDA(cudaStreamSynchronize(s));
    return CRB_OK;
}

int crb_launch_fullrank_exact(crb_handle* h, int32_t kind, const float* P, const float* Q, const float* hvec, int64_t n_items,
                              int32_t dim, const int32_t* users, const int32_t* hist_users, const int32_t* todo, int64_t n_todo,
                              int32_t K, int32_t* out_items, float* out_scores, cudaStream_t s) {
    if (n_todo <= 0) return CRB_OK;
    CRB_CHECK_ARG(K >= 1 && K <= 256, "K must be in [1,256]");
    // few users: cut the catalogue so that the machine is full (8 CTAs per SM), ranges of at least 4096 items
    int n_chunks = 1;
    if (n_todo < (int64_t)h->sm_count * 4) {
        n_chunks = (int)(((int64_t)h->sm_count * 8 + n_todo - 1) / n_todo);
        const int64_t max_chunks = (n_items + 4095) / 4096;
        if (n_chunks > max_chunks) n_chunks = (int)max_chunks;
        if (n_chunks < 1) n_chunks = 1;
    }
    const int64_t work = n_todo * n_chunks;
    int grid = (int)(work < (int64_t)h->sm_count * 8 ? work : (int64_t)h->sm_count * 8);
    const size_t sm = sizeof(float) * dim;
    unsigned long long* part = nullptr;
    if (n_chunks > 1) CRB_CUDA(cudaMallocAsync(&part, sizeof(unsigned long long) * work * K, s));
#define FR_CASE(KD)                                                                                                          \
    case KD:                                                                                                                 \
        fullrank_exact_kernel<KD><<<grid, 256, sm, s>>>(P, Q, hvec, n_items, dim, users, hist_users, todo, n_todo, h->seen_rowptr, \
                                                        h->seen_cols, K, out_items, out_scores, n_chunks, part);             \
        break;
    switch (kind) {
        FR_CASE(CRB_SCORE_DOT)
        FR_CASE(CRB_SCORE_GMF)
        FR_CASE(CRB_SCORE_SQDIST)
        FR_CASE(CRB_SCORE_DOT_BIAS)
        default: crb_set_error("unknown score kind %d", kind); return CRB_ERR_ARG;
    }
#undef FR_CASE
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    if (n_chunks > 1) {
        fullrank_merge_kernel<<<(int)n_todo, 32, 0, s>>>(part, todo, n_chunks, K, kind == CRB_SCORE_SQDIST ? 1 : 0, out_items, out_scores);
        h->launches++;
        CRB_CUDA(cudaGetLastError());
        CRB_CUDA(cudaFreeAsync(part, s));
    }
    return CRB_OK;
}
