// `numpy_stream` sampler mode: the reference's samplers (utils/sampler.py:10-99) reproduced BIT FOR BIT on the device, i.e. the
// exact arrays the reference returns after `np.random.seed(s)` (or from any np.random.get_state()).
//
// The reference draws from NumPy's global legacy MT19937 stream: np.random.randint(I) consumes 32-bit outputs with masked
// rejection (v = r & mask, retry while v > I-1), the sampler retries while the value is in the user's history or already drawn
// for the current positive, and np.random.permutation(N) is a top-down Fisher-Yates whose index draws use the same masked
// rejection with a mask that shrinks with i.  Every rejection shifts all later draws, so the process is sequential -- but each
// step is cheap and its outcome only depends on how many values were accepted before it.  Kernels:
//   mt_generate_kernel      one CTA regenerates the 624-word state in three dependency-free phases and tempers outputs in parallel
//   np_negatives_kernel     one CTA walks the raw stream in windows of 1024 values; inside a window every thread decides its
//                           value from the current guess "a_t = accepted before t" and the guess is refined by a block scan until
//                           it is a fixpoint -- which, by induction on t, is exactly the sequential result
//   np_perm_draws_kernel    same scheme for the Fisher-Yates index draws j_i (acceptance and mask depend on i = N-1-a_t)
//   perm_walk_kernel        the swap sequence as a permutation without replaying it: sort (j_i, i), then every position follows
//                           its chain "next step that pulls from where I am" by binary search (expected length O(log N))
#include <cub/cub.cuh>

#include "common.cuh"

#define NP_W 1024

struct NpState {  // device copy of RandomState: key[624], pos
    uint32_t key[624];
    int32_t pos;
};

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t c) {  // new = c ^ ((a&U | b&L) >> 1) ^ mag
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// Generates n outputs from *st (not modified unless `commit`): out[0..n).  One CTA.
__global__ void __launch_bounds__(256) mt_generate_kernel(NpState* st, int64_t n, uint32_t* out, int commit) {
    __shared__ uint32_t mt[624];
    __shared__ int s_pos;
    for (int k = threadIdx.x; k < 624; k += blockDim.x) mt[k] = st->key[k];
    if (threadIdx.x == 0) s_pos = st->pos;
    __syncthreads();
    int64_t produced = 0;
    while (produced < n) {
        int pos = s_pos;
        if (pos >= 624) {
            // k in [0,227): uses old mt[k], old mt[k+1], old mt[k+397]
            uint32_t v = 0;
            const int k0 = threadIdx.x;
            if (k0 < 227) v = mt_twist(mt[k0], mt[k0 + 1], mt[k0 + 397]);
            __syncthreads();
            if (k0 < 227) mt[k0] = v;
            __syncthreads();
            // k in [227,454): uses old mt[k], old mt[k+1], NEW mt[k-227]
            const int k1 = 227 + threadIdx.x;
            if (k1 < 454) v = mt_twist(mt[k1], mt[k1 + 1], mt[k1 - 227]);
            __syncthreads();
            if (k1 < 454) mt[k1] = v;
            __syncthreads();
            // k in [454,624): uses old mt[k], old mt[k+1] (NEW mt[0] for k = 623), NEW mt[k-227]
            const int k2 = 454 + threadIdx.x;
            if (k2 < 624) v = mt_twist(mt[k2], k2 == 623 ? mt[0] : mt[k2 + 1], mt[k2 - 227]);
            __syncthreads();
            if (k2 < 624) mt[k2] = v;
            __syncthreads();
            pos = 0;
        }
        int64_t take = 624 - pos;
        if (take > n - produced) take = n - produced;
        if (out)
            for (int k = threadIdx.x; k < take; k += blockDim.x) out[produced + k] = mt_temper(mt[pos + k]);
        produced += take;
        __syncthreads();
        if (threadIdx.x == 0) s_pos = pos + (int)take;
        __syncthreads();
    }
    if (commit) {
        for (int k = threadIdx.x; k < 624; k += blockDim.x) st->key[k] = mt[k];
        if (threadIdx.x == 0) st->pos = s_pos;
    }
}

struct NpNegArgs {
    const uint32_t* raw;
    int64_t n_raw;
    const int32_t* pos_user;
    const int64_t* seen_rowptr;
    const int32_t* seen_cols;
    int32_t n_items;
    uint32_t mask;
    int32_t R;
    int64_t n_pos;
    int32_t* negs;        // [n_pos, R] in the reference's draw order
    int64_t* result;      // [0] raws consumed, [1] negatives produced
};

__device__ __forceinline__ bool np_seen(const int32_t* __restrict__ cols, int64_t lo, int64_t hi, int32_t v) {
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cols[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo < end && cols[lo] == v;
}

__global__ void __launch_bounds__(NP_W) np_negatives_kernel(NpNegArgs a) {
    typedef cub::BlockScan<int, NP_W> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int32_t acc[2][NP_W];      // accepted values of the window by accepted-index (double buffered across iterations)
    __shared__ int32_t carry[64];         // accepted values of the group that straddles the window start
    __shared__ int s_changed, s_total, s_last;
    const int t = threadIdx.x;
    const int64_t N = a.n_pos * a.R;
    int64_t k = 0, q = 0;
    while (q < N && k < a.n_raw) {
        const int W = (int)((a.n_raw - k) < NP_W ? (a.n_raw - k) : NP_W);
        const int32_t c = t < W ? (int32_t)(a.raw[k + t] & a.mask) : 0x7fffffff;
        const bool in_range = t < W && c < a.n_items;
        const int carry_n = (int)(q % a.R);   // slots of the current group filled before this window
        int a_t, total;
        Scan(tmp).ExclusiveSum(in_range ? 1 : 0, a_t, total);
        int cur = 0;
        acc[0][t] = -1; acc[1][t] = -1;
        __syncthreads();
        if (in_range) acc[0][a_t] = c;
        bool ok = in_range;
        for (int iter = 0; iter < 4 * NP_W; ++iter) {
            if (t == 0) s_changed = 0;
            __syncthreads();
            // decide from the current guess
            bool nok = false;
            if (in_range) {
                const int64_t Q = q + a_t;
                if (Q < N) {
                    const int64_t g = Q / a.R;
                    const int s = (int)(Q % a.R);
                    const int32_t u = a.pos_user[g];
                    nok = !np_seen(a.seen_cols, a.seen_rowptr[u], a.seen_rowptr[u + 1], c);
                    // duplicates inside the group: its earlier accepted values sit at accepted-indices [a_t - s, a_t) (carry if negative)
                    for (int b = 1; b <= s && nok; ++b) {
                        const int idx = a_t - b;
                        const int32_t prev = idx >= 0 ? acc[cur][idx] : carry[carry_n + idx];
                        if (prev == c) nok = false;
                    }
                }
            }
            int na, ntotal;
            Scan(tmp).ExclusiveSum(nok ? 1 : 0, na, ntotal);
            acc[cur ^ 1][t] = -1;
            __syncthreads();
            if (nok) acc[cur ^ 1][na] = c;
            if (nok != ok || (nok && na != a_t)) s_changed = 1;
            ok = nok; a_t = na; total = ntotal; cur ^= 1;
            __syncthreads();
            if (!s_changed) break;
        }
        // commit the window
        if (ok) a.negs[q + a_t] = c;
        if (t == 0) { s_total = total; s_last = -1; }
        __syncthreads();
        const bool finishing = q + total >= N;
        if (finishing && ok && q + a_t == N - 1) s_last = t;   // the last value the epoch needs
        // new carry: accepted values of the last, possibly partial, group
        const int64_t q_new = q + total;
        const int new_carry = (int)(q_new % a.R);
        __syncthreads();
        int32_t cv = 0;
        if (t < new_carry) {
            const int idx = total - new_carry + t;   // accepted-index inside this window (may be negative: still the old carry)
            cv = idx >= 0 ? acc[cur][idx] : carry[carry_n + idx];
        }
        __syncthreads();
        if (t < new_carry) carry[t] = cv;
        __syncthreads();
        if (finishing) { k += s_last + 1; q = N; break; }
        k += W;
        q = q_new;
    }
    if (t == 0) { a.result[0] = k; a.result[1] = q; }
}

__device__ __forceinline__ uint32_t mask_of(uint32_t v) {
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    return v;
}

// Fisher-Yates index draws of np.random.permutation(N): for i = N-1 .. 1: j_i = random_interval(i) (masked rejection)
__global__ void __launch_bounds__(NP_W) np_perm_draws_kernel(const uint32_t* raw, int64_t n_raw, int64_t N, uint32_t* jd, int64_t* result) {
    typedef cub::BlockScan<int, NP_W> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int s_changed, s_last;
    const int t = threadIdx.x;
    int64_t k = 0, i_cur = N - 1;
    while (i_cur >= 1 && k < n_raw) {
        const int W = (int)((n_raw - k) < NP_W ? (n_raw - k) : NP_W);
        const uint32_t r = t < W ? raw[k + t] : 0u;
        int a_t = t, total = W;
        bool ok = t < W;
        uint32_t m = 0;
        for (int iter = 0; iter < 4 * NP_W; ++iter) {
            if (t == 0) s_changed = 0;
            __syncthreads();
            bool nok = false;
            const int64_t it = i_cur - a_t;
            if (t < W && it >= 1) {
                m = r & mask_of((uint32_t)it);
                nok = (int64_t)m <= it;
            }
            int na, ntotal;
            Scan(tmp).ExclusiveSum(nok ? 1 : 0, na, ntotal);
            if (nok != ok || (nok && na != a_t)) s_changed = 1;
            ok = nok; a_t = na; total = ntotal;
            __syncthreads();
            if (!s_changed) break;
        }
        if (ok) jd[i_cur - a_t] = m;
        if (t == 0) s_last = -1;
        __syncthreads();
        const bool finishing = i_cur - total < 1;
        if (finishing && ok && i_cur - a_t == 1) s_last = t;
        __syncthreads();
        if (finishing) { k += s_last + 1; i_cur = 0; break; }
        k += W;
        i_cur -= total;
    }
    if (t == 0) { result[0] = k; result[1] = i_cur; }
}

__global__ void __launch_bounds__(256) perm_keys_kernel(const uint32_t* jd, int64_t N, unsigned long long* keys) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride)
        keys[i] = i == 0 ? ~0ULL : (((unsigned long long)jd[i] << 32) | (unsigned long long)i);   // entry 0 is unused: sorts last
}

// x_final[p] = tau_{N-1}( ... tau_1(p)): start at j_p at time p (0 at time 0 for p = 0); repeatedly jump to the first later step that
// pulls from the current position.
__global__ void __launch_bounds__(256) perm_walk_kernel(const uint32_t* jd, const unsigned long long* keys, int64_t N, int32_t* perm) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += stride) {
        uint32_t x = p >= 1 ? jd[p] : 0u;
        uint32_t time = (uint32_t)p;
        while (true) {
            const unsigned long long want = ((unsigned long long)x << 32) | (unsigned long long)(time + 1u);
            int64_t lo = 0, hi = N - 1;   // keys[0..N-2] are the real entries (the ~0 sentinel is last)
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (keys[mid] < want) lo = mid + 1; else hi = mid;
            }
            if (lo >= N - 1) break;
            const unsigned long long kf = keys[lo];
            if ((uint32_t)(kf >> 32) != x) break;
            x = (uint32_t)kf;     // the step index i
            time = x;
        }
        perm[p] = (int32_t)x;
    }
}

// final layouts (the `arrays[s_idx]` of utils/sampler.py:69)
__global__ void __launch_bounds__(256) np_layout_pairwise_kernel(const int32_t* perm, int64_t N, int R, const int32_t* pos_user, const int32_t* pos_item,
                                                                 const int32_t* negs, const int64_t* seen_rowptr, int32_t* u, int32_t* i, int32_t* j,
                                                                 int32_t* nbr) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += stride) {
        const int64_t s = perm ? perm[k] : k;
        const int64_t p = s / R;
        const int32_t uu = pos_user[p];
        u[k] = uu; i[k] = pos_item[p]; j[k] = negs[s];
        if (nbr) nbr[k] = (int32_t)(seen_rowptr[uu + 1] - seen_rowptr[uu]);
    }
}
__global__ void __launch_bounds__(256) np_layout_pointwise_kernel(const int32_t* perm, int64_t N, int R, const int32_t* pos_user, const int32_t* pos_item,
                                                                  const int32_t* negs, int32_t* u, int32_t* i, float* y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += stride) {
        const int64_t s = perm[k];
        const int64_t p = s / (R + 1);
        const int r = (int)(s % (R + 1));
        u[k] = pos_user[p];
        i[k] = r == 0 ? pos_item[p] : negs[p * R + (r - 1)];
        y[k] = r == 0 ? 1.f : 0.f;
    }
}
__global__ void __launch_bounds__(256) np_layout_cml_kernel(const int32_t* perm, int64_t N, int R, const int32_t* pos_user, const int32_t* pos_item,
                                                            const int32_t* negs, int32_t* u, int32_t* i, int32_t* neg) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += stride) {
        const int64_t p = perm[k];
        u[k] = pos_user[p]; i[k] = pos_item[p];
        for (int r = 0; r < R; ++r) neg[k * R + r] = negs[p * R + r];
    }
}

// ------------------------------------------------------------------------------------------------ host side
static NpState* np_state(crb_handle* h) { return reinterpret_cast<NpState*>(h->np_state); }

static int np_require(crb_handle* h) {
    if (!h->np_state) {
        CRB_CUDA(cudaMalloc(&h->np_state, sizeof(NpState)));
        CRB_CUDA(cudaMemset(h->np_state, 0, sizeof(NpState)));
        h->np_seeded = 0;
    }
    return CRB_OK;
}

extern "C" int crb_np_set_state(crb_handle* h, const uint32_t* key624, int32_t pos) {
    CRB_CHECK_ARG(h && key624 && pos >= 0 && pos <= 624, "bad argument");
    CRB_CUDA(cudaSetDevice(h->device));
    int rc = np_require(h);
    if (rc) return rc;
    NpState st;
    memcpy(st.key, key624, sizeof(st.key));
    st.pos = pos;
    CRB_CUDA(cudaMemcpy(h->np_state, &st, sizeof(st), cudaMemcpyHostToDevice));
    h->np_seeded = 1;
    return CRB_OK;
}

extern "C" int crb_np_get_state(crb_handle* h, uint32_t* key624, int32_t* pos) {
    CRB_CHECK_ARG(h && key624 && pos, "bad argument");
    if (!h->np_state || !h->np_seeded) { crb_set_error("numpy stream not seeded (crb_np_seed / crb_np_set_state)"); return CRB_ERR_STATE; }
    NpState st;
    CRB_CUDA(cudaMemcpy(&st, h->np_state, sizeof(st), cudaMemcpyDeviceToHost));
    memcpy(key624, st.key, sizeof(st.key));
    *pos = st.pos;
    return CRB_OK;
}

// np.random.seed(seed) for an integer seed: init_genrand, pos = 624
extern "C" int crb_np_seed(crb_handle* h, uint32_t seed) {
    uint32_t key[624];
    key[0] = seed;
    for (int k = 1; k < 624; ++k) key[k] = 1812433253u * (key[k - 1] ^ (key[k - 1] >> 30)) + (uint32_t)k;
    return crb_np_set_state(h, key, 624);
}

// Generates n raws from the current state into the scratch buffer (state untouched); -> device pointer
static int np_generate(crb_handle* h, int64_t n, uint32_t** out, cudaStream_t s) {
    if (n > h->np_raw_cap) {
        CRB_CUDA(cudaStreamSynchronize(s));
        cudaFree(h->np_raw);
        h->np_raw = nullptr; h->np_raw_cap = 0;
        CRB_CUDA(cudaMalloc(&h->np_raw, sizeof(uint32_t) * n));
        h->np_raw_cap = n;
    }
    mt_generate_kernel<<<1, 256, 0, s>>>(np_state(h), n, (uint32_t*)h->np_raw, 0);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    *out = (uint32_t*)h->np_raw;
    return CRB_OK;
}
static int np_advance(crb_handle* h, int64_t n, cudaStream_t s) {
    if (n <= 0) return CRB_OK;
    mt_generate_kernel<<<1, 256, 0, s>>>(np_state(h), n, nullptr, 1);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

static int np_scratch(crb_handle* h, int64_t bytes) {
    if (bytes <= h->np_scratch_cap) return CRB_OK;
    CRB_CUDA(cudaDeviceSynchronize());
    cudaFree(h->np_scratch);
    h->np_scratch = nullptr; h->np_scratch_cap = 0;
    CRB_CUDA(cudaMalloc(&h->np_scratch, bytes));
    h->np_scratch_cap = bytes;
    return CRB_OK;
}

// phase 1: the negatives of every positive in the reference's order -> negs [n_pos, R] (device scratch); advances the stream
static int np_negatives(crb_handle* h, int32_t R, int32_t* negs, cudaStream_t s) {
    if (!h->np_state || !h->np_seeded) { crb_set_error("numpy stream not seeded (crb_np_seed / crb_np_set_state)"); return CRB_ERR_STATE; }
    if (!h->pos_user) { crb_set_error("sampler called before crb_set_history"); return CRB_ERR_STATE; }
    CRB_CHECK_ARG(R >= 1 && R <= 64, "neg_ratio must be in [1,64]");
    const int64_t N = h->n_pos * R;
    if (N == 0) return CRB_OK;
    uint32_t mask = (uint32_t)(h->n_items - 1);
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    int64_t n_raw = N + N / 2 + 65536;
    for (int attempt = 0; attempt < 8; ++attempt) {
        uint32_t* raw = nullptr;
        int rc = np_generate(h, n_raw, &raw, s);
        if (rc) return rc;
        NpNegArgs a = {raw, n_raw, h->pos_user, h->seen_rowptr, h->seen_cols, (int32_t)h->n_items, mask, R, h->n_pos, negs, h->np_result};
        np_negatives_kernel<<<1, NP_W, 0, s>>>(a);
        h->launches++;
        CRB_CUDA(cudaGetLastError());
        int64_t res[2];
        CRB_CUDA(cudaMemcpyAsync(res, h->np_result, sizeof(res), cudaMemcpyDeviceToHost, s));
        CRB_CUDA(cudaStreamSynchronize(s));
        if (res[1] == N) return np_advance(h, res[0], s);
        n_raw *= 2;   // ran out of raw values before the epoch was complete: regenerate a longer prefix of the same stream
    }
    crb_set_error("numpy-stream sampler: rejection rate too high (history covers the catalogue?)");
    return CRB_ERR_SAMPLER;
}

// phase 2: np.random.permutation(N) -> perm [N] int32 (device scratch); advances the stream
static int np_permutation(crb_handle* h, int64_t N, int32_t* perm, uint32_t* jd, unsigned long long* keys, unsigned long long* keys_sorted,
                          cudaStream_t s) {
    CRB_CHECK_ARG(N < 0x7fffffffLL, "epoch too long for the numpy-stream mode");
    if (N <= 1) {
        if (N == 1) CRB_CUDA(cudaMemsetAsync(perm, 0, sizeof(int32_t), s));
        return CRB_OK;
    }
    int64_t n_raw = 2 * N + 65536;
    for (int attempt = 0; attempt < 4; ++attempt) {
        uint32_t* raw = nullptr;
        int rc = np_generate(h, n_raw, &raw, s);
        if (rc) return rc;
        np_perm_draws_kernel<<<1, NP_W, 0, s>>>(raw, n_raw, N, jd, h->np_result);
        h->launches++;
        CRB_CUDA(cudaGetLastError());
        int64_t res[2];
        CRB_CUDA(cudaMemcpyAsync(res, h->np_result, sizeof(res), cudaMemcpyDeviceToHost, s));
        CRB_CUDA(cudaStreamSynchronize(s));
        if (res[1] == 0) {
            if ((rc = np_advance(h, res[0], s))) return rc;
            break;
        }
        if (attempt == 3) { crb_set_error("numpy-stream permutation: ran out of raw values"); return CRB_ERR_SAMPLER; }
        n_raw *= 2;
    }
    const int grid = h->sm_count * 8;
    perm_keys_kernel<<<grid, 256, 0, s>>>(jd, N, keys);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, keys_sorted, (int)N, 0, 64, s);
    if ((int64_t)tmp_bytes > h->np_sort_cap) {
        CRB_CUDA(cudaStreamSynchronize(s));
        cudaFree(h->np_sort_tmp);
        h->np_sort_tmp = nullptr; h->np_sort_cap = 0;
        CRB_CUDA(cudaMalloc(&h->np_sort_tmp, tmp_bytes));
        h->np_sort_cap = (int64_t)tmp_bytes;
    }
    CRB_CUDA(cub::DeviceRadixSort::SortKeys(h->np_sort_tmp, tmp_bytes, keys, keys_sorted, (int)N, 0, 64, s));
    perm_walk_kernel<<<grid, 256, 0, s>>>(jd, keys_sorted, N, perm);
    h->launches += 3;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

// kind 0 pairwise (u,i,j[,nbr]); 1 pointwise (u,i,y); 2 cml (u,i,neg[R]); 3 negatives only (out_a = negs [n_pos,R], no permutation)
extern "C" int crb_sample_epoch_numpy(crb_handle* h, int32_t kind, int32_t neg_ratio, int32_t* u, int32_t* i, void* third, int32_t* nbr,
                                      void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h, "null handle");
    CRB_CHECK_ARG(kind >= 0 && kind <= 3, "kind");
    CRB_CUDA(cudaSetDevice(h->device));
    int rc = np_require(h);
    if (rc) return rc;
    if (!h->np_result) CRB_CUDA(cudaMalloc(&h->np_result, 2 * sizeof(int64_t)));
    if (!h->pos_user) { crb_set_error("sampler called before crb_set_history"); return CRB_ERR_STATE; }
    const int64_t n_pos = h->n_pos, R = neg_ratio;
    const int64_t N = kind == 0 ? n_pos * R : (kind == 1 ? n_pos * (R + 1) : n_pos);
    if (n_pos == 0) return CRB_OK;
    CRB_CHECK_ARG(kind == 3 ? third != nullptr : (u && i && third), "null output");
    // scratch: negs | perm | jd | keys | keys_sorted
    const int64_t b_negs = ((n_pos * R * 4 + 255) / 256) * 256, b_perm = ((N * 4 + 255) / 256) * 256, b_keys = ((N * 8 + 255) / 256) * 256;
    if ((rc = np_scratch(h, b_negs + 2 * b_perm + 2 * b_keys + 1024))) return rc;
    char* base = (char*)h->np_scratch;
    int32_t* negs = kind == 3 ? (int32_t*)third : (int32_t*)base;
    int32_t* perm = (int32_t*)(base + b_negs);
    uint32_t* jd = (uint32_t*)(base + b_negs + b_perm);
    unsigned long long* keys = (unsigned long long*)(base + b_negs + 2 * b_perm);
    unsigned long long* keys_sorted = (unsigned long long*)(base + b_negs + 2 * b_perm + b_keys);
    if ((rc = np_negatives(h, neg_ratio, negs, s))) return rc;
    if (kind == 3) return CRB_OK;
    if ((rc = np_permutation(h, N, perm, jd, keys, keys_sorted, s))) return rc;
    const int grid = h->sm_count * 8;
    if (kind == 0)
        np_layout_pairwise_kernel<<<grid, 256, 0, s>>>(perm, N, (int)R, h->pos_user, h->pos_item, negs, h->seen_rowptr, u, i, (int32_t*)third, nbr);
    else if (kind == 1)
        np_layout_pointwise_kernel<<<grid, 256, 0, s>>>(perm, N, (int)R, h->pos_user, h->pos_item, negs, u, i, (float*)third);
    else
        np_layout_cml_kernel<<<grid, 256, 0, s>>>(perm, N, (int)R, h->pos_user, h->pos_item, negs, u, i, (int32_t*)third);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

// ------------------------------------------------------------------------------------------------ SBPR (utils/sampler.py:102-141)
// ranking_sampler_sbpr draws, per (positive, k < neg_ratio) SLOT: s = np.random.randint(len(SPu[u])) -- masked rejection over
// [0, len), and NO raw value at all when len == 1 (NumPy's legacy randint returns `low` for a one-value range without touching the
// stream) -- then neg = np.random.randint(item_nums) until neg is outside the user's own and social items (no distinctness between
// slots); finally one np.random.permutation over all slots.  The accepted draws form one sequence; the q-th of them belongs to the
// slot found by a binary search over prefix[] (accepted draws before each slot: 2 per slot, 1 where len(SPu) == 1) and is the social
// draw iff it is the first of a two-draw slot.  The same fixpoint-per-window scheme as np_negatives_kernel then applies.
struct NpSbprArgs {
    const uint32_t* raw;
    int64_t n_raw;
    const int64_t* prefix;       // [n_slots + 1]
    int64_t n_slots;
    int32_t R;
    const int32_t* sp_pos_user;
    const int64_t* spu_start;
    const int64_t* excl_rowptr;
    const int32_t* excl_cols;
    int32_t n_items;
    uint32_t item_mask;
    int32_t* s_out;              // [n_slots] index into SPu[u] (pre-zeroed: the value of the one-item lists)
    int32_t* neg_out;            // [n_slots]
    int64_t* result;             // [0] raws consumed, [1] accepted draws produced
};

__global__ void __launch_bounds__(256) np_sbpr_count_kernel(const int32_t* __restrict__ sp_pos_user, const int64_t* __restrict__ spu_start, int64_t n_slots,
                                                            int R, int64_t* __restrict__ cons) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_slots; k += stride) {
        const int32_t u = sp_pos_user[k / R];
        cons[k] = (spu_start[u + 1] - spu_start[u]) > 1 ? 2 : 1;
    }
}

__global__ void __launch_bounds__(NP_W) np_sbpr_draws_kernel(NpSbprArgs a) {
    typedef cub::BlockScan<int, NP_W> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int s_changed, s_last;
    const int t = threadIdx.x;
    const int64_t total_draws = a.prefix[a.n_slots];
    int64_t k = 0, q = 0;
    while (q < total_draws && k < a.n_raw) {
        const int W = (int)((a.n_raw - k) < NP_W ? (a.n_raw - k) : NP_W);
        const uint32_t r = t < W ? a.raw[k + t] : 0u;
        int a_t = t, total = W;
        bool ok = t < W;
        int64_t slot = 0;
        int32_t val = 0;
        bool social = false;
        for (int iter = 0; iter < 4 * NP_W; ++iter) {
            if (t == 0) s_changed = 0;
            __syncthreads();
            bool nok = false;
            const int64_t Q = q + a_t;
            if (t < W && Q < total_draws) {
                int64_t lo = 0, hi = a.n_slots;      // last slot with prefix[slot] <= Q
                while (hi - lo > 1) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (a.prefix[mid] <= Q) lo = mid; else hi = mid;
                }
                slot = lo;
                const int32_t u = a.sp_pos_user[slot / a.R];
                const int64_t L = a.spu_start[u + 1] - a.spu_start[u];
                social = L > 1 && Q == a.prefix[slot];
                if (social) {
                    const uint32_t m = r & mask_of((uint32_t)(L - 1));
                    nok = (int64_t)m <= L - 1;
                    val = (int32_t)m;
                } else {
                    const int32_t c = (int32_t)(r & a.item_mask);
                    nok = c < a.n_items && !np_seen(a.excl_cols, a.excl_rowptr[u], a.excl_rowptr[u + 1], c);
                    val = c;
                }
            }
            int na, ntotal;
            Scan(tmp).ExclusiveSum(nok ? 1 : 0, na, ntotal);
            if (nok != ok || (nok && na != a_t)) s_changed = 1;
            ok = nok; a_t = na; total = ntotal;
            __syncthreads();
            if (!s_changed) break;
        }
        if (ok) { if (social) a.s_out[slot] = val; else a.neg_out[slot] = val; }
        if (t == 0) s_last = -1;
        __syncthreads();
        const bool finishing = q + total >= total_draws;
        if (finishing && ok && q + a_t == total_draws - 1) s_last = t;
        __syncthreads();
        if (finishing) { k += s_last + 1; q = total_draws; break; }
        k += W;
        q += total;
    }
    if (t == 0) { a.result[0] = k; a.result[1] = q; }
}

__global__ void __launch_bounds__(256) np_layout_sbpr_kernel(const int32_t* perm, int64_t N, int R, const int32_t* sp_pos_user, const int32_t* sp_pos_item,
                                                             const int64_t* spu_start, const int32_t* spu_items, const float* spu_suk,
                                                             const int32_t* s_out, const int32_t* neg_out, int32_t* u, int32_t* i, int32_t* k_out,
                                                             int32_t* j, float* suk) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += stride) {
        const int64_t slot = perm[k];
        const int64_t p = slot / R;
        const int32_t uu = sp_pos_user[p];
        const int64_t e = spu_start[uu] + s_out[slot];
        u[k] = uu; i[k] = sp_pos_item[p]; k_out[k] = spu_items[e]; j[k] = neg_out[slot];
        if (suk) suk[k] = spu_suk[e];
    }
}

// One epoch exactly as ranking_sampler_sbpr returns it under the current NumPy stream (crb_np_seed / crb_np_set_state; structures of
// crb_set_social): u, i, i_s, i_neg int32 [sp_n_pos * neg_ratio], suk float (or NULL for is_suk=False).  DEVICE outputs.
extern "C" int crb_sample_epoch_numpy_sbpr(crb_handle* h, int32_t neg_ratio, int32_t* u, int32_t* i, int32_t* k_out, int32_t* j, float* suk,
                                           void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && u && i && k_out && j, "null argument");
    CRB_CHECK_ARG(neg_ratio >= 1 && neg_ratio <= 64, "neg_ratio must be in [1,64]");
    CRB_CUDA(cudaSetDevice(h->device));
    int rc = np_require(h);
    if (rc) return rc;
    if (!h->np_seeded) { crb_set_error("numpy stream not seeded (crb_np_seed / crb_np_set_state)"); return CRB_ERR_STATE; }
    if (!h->sp_pos_user) { crb_set_error("SBPR sampler called before crb_set_social"); return CRB_ERR_STATE; }
    if (!h->np_result) CRB_CUDA(cudaMalloc(&h->np_result, 2 * sizeof(int64_t)));
    const int64_t N = h->sp_n_pos * neg_ratio;
    if (N == 0) return CRB_OK;
    // scratch: prefix (+1) | s_out | neg_out | perm | jd | keys | keys_sorted
    auto al = [](int64_t b) { return ((b + 255) / 256) * 256; };
    const int64_t b_pre = al((N + 1) * 8), b_i32 = al(N * 4), b_keys = al(N * 8);
    if ((rc = np_scratch(h, 2 * b_pre + 4 * b_i32 + 2 * b_keys + 1024))) return rc;
    char* base = (char*)h->np_scratch;
    int64_t* cons = (int64_t*)base;
    int64_t* prefix = (int64_t*)(base + b_pre);
    int32_t* s_out = (int32_t*)(base + 2 * b_pre);
    int32_t* neg_out = (int32_t*)(base + 2 * b_pre + b_i32);
    int32_t* perm = (int32_t*)(base + 2 * b_pre + 2 * b_i32);
    uint32_t* jd = (uint32_t*)(base + 2 * b_pre + 3 * b_i32);
    unsigned long long* keys = (unsigned long long*)(base + 2 * b_pre + 4 * b_i32);
    unsigned long long* keys_sorted = (unsigned long long*)(base + 2 * b_pre + 4 * b_i32 + b_keys);
    const int grid = h->sm_count * 8;
    np_sbpr_count_kernel<<<grid, 256, 0, s>>>(h->sp_pos_user, h->spu_start, N, neg_ratio, cons);
    CRB_CUDA(cudaMemsetAsync(cons + N, 0, sizeof(int64_t), s));
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cons, prefix, (int)(N + 1), s);
    if ((int64_t)tmp_bytes > h->np_sort_cap) {
        CRB_CUDA(cudaStreamSynchronize(s));
        cudaFree(h->np_sort_tmp);
        h->np_sort_tmp = nullptr; h->np_sort_cap = 0;
        CRB_CUDA(cudaMalloc(&h->np_sort_tmp, tmp_bytes));
        h->np_sort_cap = (int64_t)tmp_bytes;
    }
    CRB_CUDA(cub::DeviceScan::ExclusiveSum(h->np_sort_tmp, tmp_bytes, cons, prefix, (int)(N + 1), s));
    CRB_CUDA(cudaMemsetAsync(s_out, 0, sizeof(int32_t) * N, s));
    h->launches += 2;
    uint32_t mask = (uint32_t)(h->n_items - 1);
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    int64_t n_raw = 3 * N + 65536;
    bool done = false;
    for (int attempt = 0; attempt < 8 && !done; ++attempt) {
        uint32_t* raw = nullptr;
        if ((rc = np_generate(h, n_raw, &raw, s))) return rc;
        NpSbprArgs a = {raw, n_raw, prefix, N, neg_ratio, h->sp_pos_user, h->spu_start, h->excl_rowptr, h->excl_cols, (int32_t)h->n_items, mask, s_out, neg_out,
                        h->np_result};
        np_sbpr_draws_kernel<<<1, NP_W, 0, s>>>(a);
        h->launches++;
        CRB_CUDA(cudaGetLastError());
        int64_t res[2], want = 0;
        CRB_CUDA(cudaMemcpyAsync(res, h->np_result, sizeof(res), cudaMemcpyDeviceToHost, s));
        CRB_CUDA(cudaMemcpyAsync(&want, prefix + N, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        CRB_CUDA(cudaStreamSynchronize(s));
        if (res[1] == want) {
            if ((rc = np_advance(h, res[0], s))) return rc;
            done = true;
        } else {
            n_raw *= 2;
        }
    }
    if (!done) { crb_set_error("numpy-stream SBPR sampler: rejection rate too high"); return CRB_ERR_SAMPLER; }
    if ((rc = np_permutation(h, N, perm, jd, keys, keys_sorted, s))) return rc;
    np_layout_sbpr_kernel<<<grid, 256, 0, s>>>(perm, N, neg_ratio, h->sp_pos_user, h->sp_pos_item, h->spu_start, h->spu_items, h->spu_suk, s_out, neg_out, u, i, k_out,
                                               j, suk);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}
