// Native builder of the training history in device form -- what model/RankingPreprocess.py:117 (`ui_train = train_data
// .groupby('u_id').i_id.apply(list).to_dict()`) plus the per-user `set(ui_train[u])` of the samplers / evaluation loops
// (utils/sampler.py:53, model/RankingRecommender.py:222-240) hold as Python dicts, lists and sets.  At 1e9 interactions those
// cannot even be materialised; here the (user, item) rows of the training split go in as two int32 columns and come out as
//   pos_user / pos_item   the positives grouped by user (ascending id, the groupby order), row order kept inside a user
//                         (= the order utils/sampler.py:50-52 enumerates them, and the `items` lists FISM / NAIS read)
//   seen_rowptr / seen_cols  per-user sorted-unique CSR (the seen-item sets; duplicates dropped)
//   list_start / list_len    where each user's list sits inside pos_item (utils/tools.py:90-97 get_ui_sp_mat)
// Two stable radix sorts, one unique, one search per user; everything on the caller's stream.
#include <cub/cub.cuh>

#include "common.cuh"

__global__ void __launch_bounds__(256) hist_pack_keys_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ it, int64_t n,
                                                             unsigned long long* __restrict__ keys) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        keys[k] = ((unsigned long long)(uint32_t)u[k] << 32) | (unsigned long long)(uint32_t)it[k];
}

__global__ void __launch_bounds__(256) hist_range_check_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ it, int64_t n,
                                                               int32_t n_users, int32_t n_items, unsigned int* __restrict__ bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        if (u[k] < 0 || u[k] >= n_users || it[k] < 0 || it[k] >= n_items) atomicAdd(bad, 1u);
}

__global__ void __launch_bounds__(256) hist_unpack_cols_kernel(const unsigned long long* __restrict__ keys, const int64_t* __restrict__ n_dev,
                                                               int32_t* __restrict__ cols) {
    const int64_t n = *n_dev;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) cols[k] = (int32_t)(uint32_t)keys[k];
}

// rowptr[u] = first position whose user id is >= u, in a sorted array.  KEYS64: user id in the high word of 64-bit keys.
template <bool KEYS64>
__global__ void __launch_bounds__(256) hist_rowptr_kernel(const void* __restrict__ sorted, const int64_t* __restrict__ n_dev, int64_t n_host,
                                                          int64_t n_users, int64_t* __restrict__ rowptr, int32_t* __restrict__ len_out) {
    const int64_t n = n_dev ? *n_dev : n_host;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u <= n_users; u += stride) {
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            const int64_t uu = KEYS64 ? (int64_t)(reinterpret_cast<const unsigned long long*>(sorted)[mid] >> 32)
                                      : (int64_t)reinterpret_cast<const int32_t*>(sorted)[mid];
            if (uu < u) lo = mid + 1; else hi = mid;
        }
        rowptr[u] = lo;
    }
    (void)len_out;
}

__global__ void __launch_bounds__(256) hist_len_kernel(const int64_t* __restrict__ start, int64_t n_users, int32_t* __restrict__ len) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_users; u += stride) len[u] = (int32_t)(start[u + 1] - start[u]);
}

static int hist_grid(crb_handle* h, int64_t n) {
    int64_t b = (n + 255) / 256, cap = (int64_t)h->sm_count * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

extern "C" int crb_build_history(crb_handle* h, const int32_t* users, const int32_t* items, int64_t n, int64_t n_users, int64_t n_items,
                                 int32_t* pos_user, int32_t* pos_item, int64_t* seen_rowptr, int32_t* seen_cols, int64_t* n_seen,
                                 int64_t* list_start, int32_t* list_len, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && users && items && pos_user && pos_item && seen_rowptr && seen_cols && n_seen, "null argument");
    CRB_CHECK_ARG(n >= 0 && n < 0x7fffffffffLL && n_users > 0 && n_items > 0 && n_users < 0x7fffffffLL && n_items < 0x7fffffffLL, "sizes");
    CRB_CHECK_ARG(crb_is_device_ptr(pos_user) && crb_is_device_ptr(pos_item) && crb_is_device_ptr(seen_rowptr) && crb_is_device_ptr(seen_cols),
                  "outputs must be device pointers");
    CRB_CHECK_ARG(!list_start || (crb_is_device_ptr(list_start) && list_len && crb_is_device_ptr(list_len)), "list outputs must be device pointers");
    CRB_CUDA(cudaSetDevice(h->device));
    *n_seen = 0;
    if (n == 0) {
        CRB_CUDA(cudaMemsetAsync(seen_rowptr, 0, sizeof(int64_t) * (n_users + 1), s));
        if (list_start) {
            CRB_CUDA(cudaMemsetAsync(list_start, 0, sizeof(int64_t) * (n_users + 1), s));
            CRB_CUDA(cudaMemsetAsync(list_len, 0, sizeof(int32_t) * n_users, s));
        }
        return CRB_OK;
    }
    // inputs may be host columns (the reference hands over a pandas frame): stage them
    int32_t *du = nullptr, *di = nullptr;
    const bool u_dev = crb_is_device_ptr(users), i_dev = crb_is_device_ptr(items);
    if (!u_dev) { CRB_CUDA(cudaMallocAsync(&du, sizeof(int32_t) * n, s)); CRB_CUDA(cudaMemcpyAsync(du, users, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s)); }
    if (!i_dev) { CRB_CUDA(cudaMallocAsync(&di, sizeof(int32_t) * n, s)); CRB_CUDA(cudaMemcpyAsync(di, items, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s)); }
    const int32_t* u_in = u_dev ? users : du;
    const int32_t* i_in = i_dev ? items : di;

    unsigned long long *keys = nullptr, *keys_sorted = nullptr;
    int64_t* n_dev = nullptr;
    unsigned int* bad = nullptr;
    void* tmp = nullptr;
    size_t tmp_a = 0, tmp_b = 0, tmp_c = 0;
    int ubits = 1;
    while (ubits < 32 && (1LL << ubits) < n_users) ++ubits;
    int ibits = 1;
    while (ibits < 32 && (1LL << ibits) < n_items) ++ibits;
    CRB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_a, u_in, pos_user, i_in, pos_item, n, 0, ubits, s));
    CRB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_b, keys, keys_sorted, n, 0, 32 + ubits, s));
    CRB_CUDA(cub::DeviceSelect::Unique(nullptr, tmp_c, keys_sorted, keys, n_dev, n, s));
    size_t tmp_bytes = tmp_a > tmp_b ? tmp_a : tmp_b;
    if (tmp_c > tmp_bytes) tmp_bytes = tmp_c;
    CRB_CUDA(cudaMallocAsync(&keys, sizeof(unsigned long long) * n, s));
    CRB_CUDA(cudaMallocAsync(&keys_sorted, sizeof(unsigned long long) * n, s));
    CRB_CUDA(cudaMallocAsync(&n_dev, sizeof(int64_t) + sizeof(unsigned int) * 2, s));
    CRB_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, s));
    bad = reinterpret_cast<unsigned int*>(n_dev + 1);
    CRB_CUDA(cudaMemsetAsync(n_dev, 0, sizeof(int64_t) + sizeof(unsigned int) * 2, s));
    const int grid = hist_grid(h, n);
    hist_range_check_kernel<<<grid, 256, 0, s>>>(u_in, i_in, n, (int32_t)n_users, (int32_t)n_items, bad);
    // 1. positives grouped by user, row order kept: LSD radix sort is stable
    CRB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_a, u_in, pos_user, i_in, pos_item, n, 0, ubits, s));
    // 2. seen sets: sort (user, item) keys, drop duplicates
    hist_pack_keys_kernel<<<grid, 256, 0, s>>>(u_in, i_in, n, keys);
    CRB_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_b, keys, keys_sorted, n, 0, 32 + ubits, s));
    CRB_CUDA(cub::DeviceSelect::Unique(tmp, tmp_c, keys_sorted, keys, n_dev, n, s));
    hist_unpack_cols_kernel<<<grid, 256, 0, s>>>(keys, n_dev, seen_cols);
    hist_rowptr_kernel<true><<<hist_grid(h, n_users + 1), 256, 0, s>>>(keys, n_dev, 0, n_users, seen_rowptr, nullptr);
    if (list_start) {
        hist_rowptr_kernel<false><<<hist_grid(h, n_users + 1), 256, 0, s>>>(pos_user, nullptr, n, n_users, list_start, nullptr);
        hist_len_kernel<<<hist_grid(h, n_users), 256, 0, s>>>(list_start, n_users, list_len);
    }
    h->launches += 7;
    CRB_CUDA(cudaGetLastError());
    struct { int64_t n; unsigned int bad[2]; } host;
    CRB_CUDA(cudaMemcpyAsync(&host, n_dev, sizeof(host), cudaMemcpyDeviceToHost, s));
    CRB_CUDA(cudaFreeAsync(tmp, s));
    CRB_CUDA(cudaFreeAsync(keys, s));
    CRB_CUDA(cudaFreeAsync(keys_sorted, s));
    CRB_CUDA(cudaFreeAsync(n_dev, s));
    if (du) CRB_CUDA(cudaFreeAsync(du, s));
    if (di) CRB_CUDA(cudaFreeAsync(di, s));
    CRB_CUDA(cudaStreamSynchronize(s));
    (void)ibits;
    if (host.bad[0]) { crb_set_error("crb_build_history: %u rows with a user or item id outside [0, n_users) x [0, n_items)", host.bad[0]); return CRB_ERR_ARG; }
    *n_seen = host.n;
    return CRB_OK;
}
