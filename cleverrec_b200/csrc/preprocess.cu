// Device form of model/RankingPreprocess.py (reference): what runs BEFORE the hot path and produces every one of its inputs.
//   crb_prep_filter_reindex   _filter_users / _filter_items (:70-90) + re_index (:41-47): rows of users with fewer than user_min
//                             interactions are dropped, then rows of items with fewer than item_min (counted on what is left);
//                             the surviving raw ids are renumbered 0..n-1
//   crb_prep_split_loo        the leave-one-out split (:96-109): per user, in (time, file) order, the last row goes to the test set
//                             when the user has more than 3 rows
//   crb_prep_eval_negatives   the sampled evaluation negatives (:120-129): neg_samples distinct items outside the user's training
//                             items, per test user
// The reference holds all of this as pandas frames, dicts of lists and Python sets (a dict of 1e7 lists / 1e9 boxed ints cannot be
// materialised); here the interaction log is two int64 columns on the device and every step is a radix sort, a run-length encode, a
// flagged compaction or a thread-per-row kernel on the caller's stream.
//
// Id order.  The reference numbers users / items in the iteration order of a Python set of the raw ids.  For non-negative ids below
// the set's table size that order is ascending, which is what the ascending renumbering here reproduces (ml-100k, ml-1m, Ciao,
// Epinions: checked against the golden splits); for sparse id spaces the two numberings differ by a permutation of the ids, which
// changes no metric.
#include <cub/cub.cuh>

#include "common.cuh"

static int prep_grid(crb_handle* h, int64_t n) {
    int64_t b = (n + 255) / 256, cap = (int64_t)h->sm_count * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// rank[k] = position of raw[k] in the sorted unique array `uniq` (n_uniq entries)
__global__ void __launch_bounds__(256) prep_rank_kernel(const int64_t* __restrict__ raw, int64_t n, const int64_t* __restrict__ uniq,
                                                        const int64_t* __restrict__ n_uniq_dev, int32_t* __restrict__ rank) {
    const int64_t m = *n_uniq_dev;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const int64_t v = raw[k];
        int64_t lo = 0, hi = m;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (uniq[mid] < v) lo = mid + 1; else hi = mid;
        }
        rank[k] = (int32_t)lo;
    }
}

__global__ void __launch_bounds__(256) prep_flag_kernel(const int32_t* __restrict__ rank, const int32_t* __restrict__ counts, int64_t n,
                                                        int32_t minimum, unsigned char* __restrict__ keep) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) keep[k] = counts[rank[k]] >= minimum ? 1 : 0;
}

__global__ void __launch_bounds__(256) prep_iota_kernel(int64_t* __restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[k] = k;
}

struct PrepScratch {
    crb_handle* h;
    cudaStream_t s;
    void* ptrs[64];
    int n_ptrs;
    template <typename T>
    int take(T** out, int64_t count) {
        void* p = nullptr;
        CRB_CUDA(cudaMallocAsync(&p, sizeof(T) * (size_t)(count > 0 ? count : 1), s));
        if (n_ptrs < 64) ptrs[n_ptrs++] = p;
        *out = (T*)p;
        return CRB_OK;
    }
    void release() {
        for (int k = 0; k < n_ptrs; ++k) cudaFreeAsync(ptrs[k], s);
        n_ptrs = 0;
    }
};

// sorted unique values of raw[0..n) with their counts, and every row's rank among them.  Synchronises (n_uniq is a host value).
static int rank_ids(PrepScratch& sc, const int64_t* raw, int64_t n, int64_t* uniq, int32_t* counts, int32_t* rank, int64_t* n_uniq) {
    crb_handle* h = sc.h;
    cudaStream_t s = sc.s;
    int64_t *sorted = nullptr, *n_dev = nullptr;
    void* tmp = nullptr;
    size_t ta = 0, tb = 0;
    int rc;
    if ((rc = sc.take(&sorted, n))) return rc;
    if ((rc = sc.take(&n_dev, 1))) return rc;
    CRB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, ta, raw, sorted, n, 0, 64, s));
    CRB_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tb, sorted, uniq, counts, n_dev, n, s));
    if ((rc = sc.take((char**)&tmp, (int64_t)(ta > tb ? ta : tb)))) return rc;
    CRB_CUDA(cub::DeviceRadixSort::SortKeys(tmp, ta, raw, sorted, n, 0, 64, s));
    CRB_CUDA(cub::DeviceRunLengthEncode::Encode(tmp, tb, sorted, uniq, counts, n_dev, n, s));
    prep_rank_kernel<<<prep_grid(h, n), 256, 0, s>>>(raw, n, uniq, n_dev, rank);
    h->launches += 3;
    CRB_CUDA(cudaGetLastError());
    CRB_CUDA(cudaMemcpyAsync(n_uniq, n_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CRB_CUDA(cudaStreamSynchronize(s));
    return CRB_OK;
}

// keeps the rows whose id in `col` occurs at least `minimum` times; the three row arrays are compacted in place (order kept)
static int drop_rare(PrepScratch& sc, int64_t** cu, int64_t** ci, int64_t** crow, int64_t* n_cur, bool by_user, int32_t minimum,
                     int64_t* uniq, int32_t* counts, int32_t* rank) {
    crb_handle* h = sc.h;
    cudaStream_t s = sc.s;
    const int64_t n = *n_cur;
    if (n == 0) return CRB_OK;
    int64_t n_uniq = 0;
    int rc = rank_ids(sc, by_user ? *cu : *ci, n, uniq, counts, rank, &n_uniq);
    if (rc) return rc;
    unsigned char* keep = nullptr;
    int64_t *ou = nullptr, *oi = nullptr, *orow = nullptr, *n_dev = nullptr;
    void* tmp = nullptr;
    size_t tb = 0;
    if ((rc = sc.take(&keep, n)) || (rc = sc.take(&ou, n)) || (rc = sc.take(&oi, n)) || (rc = sc.take(&orow, n)) || (rc = sc.take(&n_dev, 1))) return rc;
    prep_flag_kernel<<<prep_grid(h, n), 256, 0, s>>>(rank, counts, n, minimum, keep);
    CRB_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, *cu, keep, ou, n_dev, n, s));
    if ((rc = sc.take((char**)&tmp, (int64_t)tb))) return rc;
    CRB_CUDA(cub::DeviceSelect::Flagged(tmp, tb, *cu, keep, ou, n_dev, n, s));
    CRB_CUDA(cub::DeviceSelect::Flagged(tmp, tb, *ci, keep, oi, n_dev, n, s));
    CRB_CUDA(cub::DeviceSelect::Flagged(tmp, tb, *crow, keep, orow, n_dev, n, s));
    h->launches += 4;
    CRB_CUDA(cudaMemcpyAsync(n_cur, n_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CRB_CUDA(cudaStreamSynchronize(s));
    *cu = ou; *ci = oi; *crow = orow;
    return CRB_OK;
}

__global__ void __launch_bounds__(256) prep_copy64_kernel(const int64_t* __restrict__ src, int64_t* __restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) dst[k] = src[k];
}

extern "C" int crb_prep_filter_reindex(crb_handle* h, const int64_t* raw_u, const int64_t* raw_i, int64_t n, int32_t user_min, int32_t item_min,
                                       int32_t* out_u, int32_t* out_i, int64_t* out_row, int64_t* n_kept, int64_t* n_users, int64_t* n_items,
                                       int64_t* user_ids, int64_t* item_ids, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && raw_u && raw_i && out_u && out_i && out_row && n_kept && n_users && n_items && user_ids && item_ids, "null argument");
    CRB_CHECK_ARG(n >= 0 && n < 0x7fffffffffLL, "row count");
    CRB_CHECK_ARG(crb_is_device_ptr(raw_u) && crb_is_device_ptr(raw_i) && crb_is_device_ptr(out_u) && crb_is_device_ptr(out_i) &&
                  crb_is_device_ptr(out_row) && crb_is_device_ptr(user_ids) && crb_is_device_ptr(item_ids), "columns must be device pointers");
    CRB_CUDA(cudaSetDevice(h->device));
    *n_kept = *n_users = *n_items = 0;
    if (n == 0) return CRB_OK;
    PrepScratch sc = {h, s, {nullptr}, 0};
    struct Release { PrepScratch* sc; ~Release() { sc->release(); } } rel{&sc};
    int64_t *cu = nullptr, *ci = nullptr, *crow = nullptr, *uniq = nullptr;
    int32_t *counts = nullptr, *rank = nullptr;
    int rc;
    if ((rc = sc.take(&crow, n)) || (rc = sc.take(&uniq, n)) || (rc = sc.take(&counts, n)) || (rc = sc.take(&rank, n))) return rc;
    prep_iota_kernel<<<prep_grid(h, n), 256, 0, s>>>(crow, n);
    cu = const_cast<int64_t*>(raw_u);
    ci = const_cast<int64_t*>(raw_i);   // never written: drop_rare compacts into fresh arrays
    int64_t n_cur = n;
    if (user_min > 0 && (rc = drop_rare(sc, &cu, &ci, &crow, &n_cur, true, user_min, uniq, counts, rank))) return rc;
    if (item_min > 0 && (rc = drop_rare(sc, &cu, &ci, &crow, &n_cur, false, item_min, uniq, counts, rank))) return rc;
    *n_kept = n_cur;
    if (n_cur == 0) return CRB_OK;
    // renumber: new id = rank of the raw id among the surviving ids (ascending)
    if ((rc = rank_ids(sc, cu, n_cur, user_ids, counts, out_u, n_users))) return rc;
    if ((rc = rank_ids(sc, ci, n_cur, item_ids, counts, out_i, n_items))) return rc;
    prep_copy64_kernel<<<prep_grid(h, n_cur), 256, 0, s>>>(crow, out_row, n_cur);
    h->launches += 2;
    CRB_CUDA(cudaGetLastError());
    CRB_CUDA(cudaStreamSynchronize(s));
    return CRB_OK;
}

// ------------------------------------------------------------------------------------------------ leave-one-out split
__global__ void __launch_bounds__(256) prep_gather32_kernel(const int32_t* __restrict__ src, const int64_t* __restrict__ idx, int64_t n,
                                                            int32_t* __restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) dst[k] = src[idx[k]];
}

__global__ void __launch_bounds__(256) prep_count_kernel(const int32_t* __restrict__ u, int64_t n, int32_t* __restrict__ count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) atomicAdd(count + u[k], 1);
}

__global__ void __launch_bounds__(256) prep_mark_last_kernel(const int32_t* __restrict__ us, int64_t n, const int32_t* __restrict__ count,
                                                             unsigned char* __restrict__ is_test) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const bool last = p == n - 1 || us[p + 1] != us[p];
        is_test[p] = (last && count[us[p]] > 3) ? 1 : 0;   // "Users with <= 3 interactions are divided into training set" (:101-103)
    }
}

// u [n]: re-indexed user of each row (file order); time [n] or NULL (data.split_by_time).  Outputs, both in the order the
// reference's grouped frame enumerates the rows -- ascending user, then time (ties: file order; pandas' two-key sort is a stable
// lexsort), or file order without `time`: perm [n] = original row of each position, is_test [n] = 1 for the user's last row when the
// user has more than 3 rows.
extern "C" int crb_prep_split_loo(crb_handle* h, const int32_t* u, const int64_t* time, int64_t n, int64_t n_users, int64_t* perm,
                                  unsigned char* is_test, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && u && perm && is_test, "null argument");
    CRB_CHECK_ARG(n >= 0 && n_users > 0 && n_users < 0x7fffffffLL, "sizes");
    CRB_CHECK_ARG(crb_is_device_ptr(u) && crb_is_device_ptr(perm) && crb_is_device_ptr(is_test) && (!time || crb_is_device_ptr(time)), "device pointers");
    CRB_CUDA(cudaSetDevice(h->device));
    if (n == 0) return CRB_OK;
    PrepScratch sc = {h, s, {nullptr}, 0};
    struct Release { PrepScratch* sc; ~Release() { sc->release(); } } rel{&sc};
    int64_t *idx0 = nullptr, *idx1 = nullptr, *tsorted = nullptr;
    int32_t *ug = nullptr, *us = nullptr, *count = nullptr;
    void* tmp = nullptr;
    int rc;
    if ((rc = sc.take(&idx0, n)) || (rc = sc.take(&idx1, n)) || (rc = sc.take(&ug, n)) || (rc = sc.take(&us, n)) || (rc = sc.take(&count, n_users))) return rc;
    prep_iota_kernel<<<prep_grid(h, n), 256, 0, s>>>(idx0, n);
    int ubits = 1;
    while (ubits < 32 && (1LL << ubits) < n_users) ++ubits;
    size_t ta = 0, tb = 0;
    CRB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, ta, time, tsorted, idx0, idx1, n, 0, 64, s));
    CRB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, ug, us, idx1, perm, n, 0, ubits, s));
    if ((rc = sc.take((char**)&tmp, (int64_t)(ta > tb ? ta : tb)))) return rc;
    const int64_t* order = idx0;
    if (time) {
        // least significant key first: stable sort of the row indices by time (signed int64 keys sort correctly as long as times are
        // non-negative, which timestamps are)
        if ((rc = sc.take(&tsorted, n))) return rc;
        CRB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, ta, time, tsorted, idx0, idx1, n, 0, 64, s));
        order = idx1;
    }
    prep_gather32_kernel<<<prep_grid(h, n), 256, 0, s>>>(u, order, n, ug);
    CRB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, ug, us, order, perm, n, 0, ubits, s));   // stable: (time, file) order kept inside a user
    CRB_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * n_users, s));
    prep_count_kernel<<<prep_grid(h, n), 256, 0, s>>>(u, n, count);
    prep_mark_last_kernel<<<prep_grid(h, n), 256, 0, s>>>(us, n, count, is_test);
    h->launches += 6;
    CRB_CUDA(cudaGetLastError());
    CRB_CUDA(cudaStreamSynchronize(s));
    return CRB_OK;
}

// ------------------------------------------------------------------------------------------------ evaluation negatives
// One warp per test user draws neg_samples DISTINCT items outside the user's training items (the history installed with
// crb_set_history / crb_build_history): `np.random.choice(list(item_set - seen_items), size=neg_samples, replace=False)` (:125) as a
// law -- uniform without replacement over the unseen items -- not as NumPy's stream (which permutes the whole candidate list per
// user: O(users x items)).  Candidates come 32 at a time from Philox4x32-10 keyed by (seed, user, round), masked to the next power
// of two like the training sampler; a candidate is accepted in lane order if it is in range, unseen, not yet accepted (a 2048-slot
// hash set in shared memory) and the first of its value in the round -- so the output is a pure function of (seed, user, history).
// Integer-exact CPU twin: oracle/philox.py::sample_eval_negatives.
#define EN_SLOTS 2048
#define EN_MAX 1024
__device__ __forceinline__ bool en_history_contains(const int32_t* __restrict__ cols, int64_t lo, int64_t hi, int32_t v) {
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int32_t c = __ldg(cols + mid);
        if (c < v) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(cols + lo) == v;
}

__global__ void __launch_bounds__(128) prep_eval_negatives_kernel(uint32_t k0, uint32_t k1, const int32_t* __restrict__ users, int64_t n_test,
                                                                  int32_t neg, int32_t n_items, uint32_t item_mask,
                                                                  const int64_t* __restrict__ seen_rowptr, const int32_t* __restrict__ seen_cols,
                                                                  int32_t* __restrict__ out, unsigned int* __restrict__ err) {
    __shared__ int32_t s_set[4][EN_SLOTS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int32_t* set = s_set[w];
    const int64_t gw = (int64_t)blockIdx.x * 4 + w, nw = (int64_t)gridDim.x * 4;
    for (int64_t k = gw; k < n_test; k += nw) {
        for (int q = lane; q < EN_SLOTS; q += 32) set[q] = -1;
        __syncwarp();
        const int32_t u = users[k];
        const int64_t lo = seen_rowptr[u], hi = seen_rowptr[u + 1];
        int32_t got = 0;
        uint32_t round = 0;
        for (; got < neg && round < 65536u; ++round) {
            // lane l of round r takes word (l & 3) of Philox block (u, r * 8 + l / 4)
            uint32_t wds[4];
            philox4x32_10((uint32_t)u, 0xE7A1u, round * 8u + (uint32_t)(lane >> 2), 0xFFFFFFFDu, k0, k1, wds);
            const int32_t v = (int32_t)(wds[lane & 3] & item_mask);
            bool ok = v < n_items && !en_history_contains(seen_cols, lo, hi, v);
            if (ok) {   // already accepted in an earlier round?
                uint32_t slot = ((uint32_t)v * 0x9E3779B1u) >> 21;   // 11 bits
                while (true) {
                    const int32_t e = set[slot];
                    if (e == v) { ok = false; break; }
                    if (e < 0) break;
                    slot = (slot + 1) & (EN_SLOTS - 1);
                }
            }
            const unsigned same = __match_any_sync(0xffffffffu, ok ? v : -1 - lane);
            ok = ok && (__ffs(same) - 1 == lane);          // the first lane holding this value in the round
            const unsigned acc = __ballot_sync(0xffffffffu, ok);
            const int32_t pos = got + __popc(acc & ((1u << lane) - 1u));
            if (ok && pos < neg) {
                out[k * neg + pos] = v;
                uint32_t slot = ((uint32_t)v * 0x9E3779B1u) >> 21;
                while (atomicCAS(set + slot, -1, v) != -1) slot = (slot + 1) & (EN_SLOTS - 1);
            }
            got += __popc(acc);
            __syncwarp();
        }
        if (got < neg && lane == 0) atomicAdd(err, 1u);
        __syncwarp();
    }
}

extern "C" int crb_prep_eval_negatives(crb_handle* h, uint64_t seed, const int32_t* test_users, int64_t n_test, int32_t neg_samples,
                                       int32_t* out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && test_users && out, "null argument");
    CRB_CHECK_ARG(n_test >= 0 && neg_samples >= 1 && neg_samples <= EN_MAX, "neg_samples must be in [1, 1024]");
    CRB_CHECK_ARG(crb_is_device_ptr(test_users) && crb_is_device_ptr(out), "device pointers");
    if (!h->seen_rowptr) { crb_set_error("crb_prep_eval_negatives before crb_set_history"); return CRB_ERR_STATE; }
    if (n_test == 0) return CRB_OK;
    CRB_CUDA(cudaSetDevice(h->device));
    uint32_t m = (uint32_t)(h->n_items > 0 ? h->n_items - 1 : 0);
    m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16;
    unsigned int* err = nullptr;
    CRB_CUDA(cudaMallocAsync(&err, sizeof(unsigned int), s));
    CRB_CUDA(cudaMemsetAsync(err, 0, sizeof(unsigned int), s));
    int64_t grid = (n_test + 3) / 4;
    if (grid > (int64_t)h->sm_count * 8) grid = (int64_t)h->sm_count * 8;
    prep_eval_negatives_kernel<<<(int)grid, 128, 0, s>>>((uint32_t)seed, (uint32_t)(seed >> 32), test_users, n_test, neg_samples, (int32_t)h->n_items, m,
                                                         h->seen_rowptr, h->seen_cols, out, err);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    unsigned int bad = 0;
    CRB_CUDA(cudaMemcpyAsync(&bad, err, sizeof(bad), cudaMemcpyDeviceToHost, s));
    CRB_CUDA(cudaFreeAsync(err, s));
    CRB_CUDA(cudaStreamSynchronize(s));
    if (bad) { crb_set_error("evaluation negatives: %u users have fewer than neg_samples unseen items", bad); return CRB_ERR_SAMPLER; }
    return CRB_OK;
}
