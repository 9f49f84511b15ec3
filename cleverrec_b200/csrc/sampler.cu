// On-device triplet sampler: replaces the Python loops of utils/sampler.py:10-99 (reference).
// Spec shared bit-for-bit with oracle/philox.py (see its docstring): Philox4x32-10 candidate words per
// positive, masked rejection (np.random.randint's rule), rejection against the user's sorted history and
// against the positive's already accepted negatives (utils/sampler.py:58-61), epoch shuffle by a keyed
// Feistel bijection (utils/sampler.py:68) so no permutation array is ever materialised.
#include "common.cuh"

#define CRB_MAX_NEG 64

struct SamplerArgs {
    uint32_t keys[6];
    uint32_t half_bits;
    uint32_t half_mask;
    uint64_t n_rows;       // rows in the epoch (N)
    uint32_t group;        // rows per positive
    uint32_t neg_ratio;
    uint32_t k0, k1, epoch;
    uint32_t item_mask;
    int32_t n_items;
    const int32_t* pos_user;
    const int32_t* pos_item;
    const int64_t* seen_rowptr;
    const int32_t* seen_cols;
    const uint32_t* bloom;  // per-user seen-item Bloom filters (NULL: exact search only)
    int bloom_shift;
    int bloom_exact;        // 1: the filter is the exact seen-item bitmap (bit = item id)
    int64_t bloom_stride;   // words per user
    const uint32_t* dyn;    // NULL, or (epoch graph) device words that replace keys[0..5] and epoch when the kernel starts
};

__device__ __forceinline__ uint32_t mix32(uint32_t x, uint32_t k) {
    x ^= k;
    x *= 0x85EBCA6Bu;
    x ^= x >> 13;
    x *= 0xC2B2AE35u;
    x ^= x >> 16;
    return x;
}

__device__ __forceinline__ uint64_t feistel_perm(uint64_t k, const SamplerArgs& a) {
    uint64_t x = k;
    do {
        uint32_t l = (uint32_t)(x >> a.half_bits), r = (uint32_t)x & a.half_mask;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            uint32_t nl = r;
            r = l ^ (mix32(r, a.keys[q]) & a.half_mask);
            l = nl;
        }
        x = ((uint64_t)l << a.half_bits) | r;
    } while (x >= a.n_rows);
    return x;
}

// returns true when v is in cols[lo,hi)
__device__ __forceinline__ bool history_contains(const int32_t* __restrict__ cols, int64_t lo, int64_t hi, int32_t v) {
    int64_t end = hi;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        int32_t c = __ldg(cols + mid);
        if (c < v) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(cols + lo) == v;
}

// Accepts negatives of positive p until `need` are found; acc[] receives them in order.  false on attempt overflow.
// The four candidates of a Philox block are tested against the user's Bloom filter with four independent probes issued together
// (one round trip); only a set bit (or no filter) costs the dependent binary search over the sorted history.  The accept / reject
// decisions, hence the output, are those of the exact test alone.
__device__ __forceinline__ bool draw_negatives(uint64_t p, int32_t u, uint32_t need, int32_t* acc, const SamplerArgs& a) {
    const uint32_t* bl = a.bloom ? a.bloom + (int64_t)u * a.bloom_stride : nullptr;
    int64_t lo = 0, hi = -1;     // the history bounds are fetched only if a candidate needs the exact test
    if (!bl) { lo = a.seen_rowptr[u]; hi = a.seen_rowptr[u + 1]; }
    uint32_t got = 0;
    for (uint32_t blk = 0; blk < CRB_SAMPLER_MAX_BLOCKS; ++blk) {
        uint32_t w[4], bw[4];
        philox4x32_10((uint32_t)p, (uint32_t)(p >> 32), blk, a.epoch, a.k0, a.k1, w);
        // probe for as many candidates as are still needed plus one spare (a random 32-byte sector each)
        const uint32_t probes = need - got + 1u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t v = w[q] & a.item_mask;
            const uint32_t bit = a.bloom_exact ? v : crb_bloom_bit(v, a.bloom_shift);
            bw[q] = (bl && (uint32_t)q < probes && (int32_t)v < a.n_items) ? (__ldg(bl + (bit >> 5)) >> (bit & 31)) & 1u : 2u;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int32_t v = (int32_t)(w[q] & a.item_mask);
            if (v >= a.n_items) continue;
            if (bl && bw[q] == 2u) {   // not probed yet
                const uint32_t bit = a.bloom_exact ? (uint32_t)v : crb_bloom_bit((uint32_t)v, a.bloom_shift);
                bw[q] = (__ldg(bl + (bit >> 5)) >> (bit & 31)) & 1u;
            }
            if (bw[q]) {
                if (bl && a.bloom_exact) continue;   // the bitmap is the history
                if (hi < 0) { lo = a.seen_rowptr[u]; hi = a.seen_rowptr[u + 1]; }
                if (history_contains(a.seen_cols, lo, hi, v)) continue;
            }
            bool dup = false;
            for (uint32_t s = 0; s < got; ++s) dup |= (acc[s] == v);
            if (dup) continue;
            acc[got++] = v;
            if (got == need) return true;
        }
    }
    return false;
}

__device__ __forceinline__ uint32_t bump(unsigned long long* meta, int32_t row) {
    return atomicAdd(reinterpret_cast<unsigned int*>(meta + row), 1u);  // low word = count (little endian)
}

// kind 0: pairwise (u,i,j[,nbr]); kind 1: pointwise (u,i,y[,nbr]); kind 2: cml (u,i,neg[R])
template <int KIND, bool COUNT, bool DYN = false>
__global__ void __launch_bounds__(256) sample_kernel(SamplerArgs a, int64_t first, int64_t count, int32_t* __restrict__ ou,
                                                     int32_t* __restrict__ oi, int32_t* __restrict__ oj, float* __restrict__ oy,
                                                     int32_t* __restrict__ onbr, unsigned long long* metaU,
                                                     unsigned long long* metaI, uint32_t* __restrict__ rk0,
                                                     uint32_t* __restrict__ rk1, uint32_t* __restrict__ rk2,
                                                     crb_step_ctr* ctr) {
    if (DYN) {   // epoch-graph replays only: the keys become registers (16 more than the constant-bank operands of the plain kernel)
#pragma unroll
        for (int q = 0; q < 6; ++q) a.keys[q] = a.dyn[q];
        a.epoch = a.dyn[6];
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const uint64_t s = feistel_perm((uint64_t)(first + t), a);
        const uint64_t p = s / a.group;
        const uint32_t r = (uint32_t)(s % a.group);
        const int32_t u = a.pos_user[p];
        const int32_t it = a.pos_item[p];
        int32_t acc[CRB_MAX_NEG];
        bool ok = true;
        if (KIND == 0) {
            ok = draw_negatives(p, u, r + 1, acc, a);
            const int32_t j = ok ? acc[r] : 0;
            ou[t] = u; oi[t] = it; oj[t] = j;
            if (COUNT) {
                rk0[t] = bump(metaU, u);
                rk1[t] = bump(metaI, it);
                rk2[t] = bump(metaI, j);
            }
        } else if (KIND == 1) {
            int32_t item = it;
            if (r > 0) {
                ok = draw_negatives(p, u, r, acc, a);
                item = ok ? acc[r - 1] : 0;
            }
            ou[t] = u; oi[t] = item; oy[t] = (r == 0) ? 1.f : 0.f;
            if (COUNT) {
                rk0[t] = bump(metaU, u);
                rk1[t] = bump(metaI, item);
            }
        } else {
            ok = draw_negatives(p, u, a.neg_ratio, acc, a);
            ou[t] = u; oi[t] = it;
            for (uint32_t q = 0; q < a.neg_ratio; ++q) oj[t * a.neg_ratio + q] = ok ? acc[q] : 0;
        }
        if (onbr) onbr[t] = (int32_t)(a.seen_rowptr[u + 1] - a.seen_rowptr[u]);
        if (!ok) atomicAdd(&ctr->sampler_err, 1u);
    }
}

// NAIS per-user batch (RankingRecommender.py:64-80): for every positive of the user, in list order: the positive (label 1) then
// its neg_ratio negatives (label 0).  One thread per positive; same negative stream as the other samplers.
__global__ void __launch_bounds__(128) sample_nais_kernel(SamplerArgs a, int64_t pos_first, int n_pos_user, int32_t* __restrict__ targets,
                                                          float* __restrict__ y, crb_step_ctr* ctr) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pos_user) return;
    const uint64_t p = (uint64_t)(pos_first + k);
    const int32_t u = a.pos_user[p];
    int32_t acc[CRB_MAX_NEG];
    const bool ok = draw_negatives(p, u, a.neg_ratio, acc, a);
    const int64_t base = (int64_t)k * (a.neg_ratio + 1);
    targets[base] = a.pos_item[p];
    y[base] = 1.f;
    for (uint32_t q = 0; q < a.neg_ratio; ++q) { targets[base + 1 + q] = ok ? acc[q] : 0; y[base + 1 + q] = 0.f; }
    if (!ok) atomicAdd(&ctr->sampler_err, 1u);
}

void crb_sampler_perm_keys(uint64_t seed, uint32_t epoch, uint32_t keys[6]);
static void host_perm_keys(uint64_t seed, uint32_t epoch, uint32_t keys[6]) { crb_sampler_perm_keys(seed, epoch, keys); }
void crb_sampler_perm_keys(uint64_t seed, uint32_t epoch, uint32_t keys[6]) {
    uint32_t a[4], b[4];
    philox4x32_10(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, epoch, (uint32_t)seed, (uint32_t)(seed >> 32), a);
    philox4x32_10(0xFFFFFFFFu, 0xFFFFFFFFu, 1u, epoch, (uint32_t)seed, (uint32_t)(seed >> 32), b);
    keys[0] = a[0]; keys[1] = a[1]; keys[2] = a[2]; keys[3] = a[3]; keys[4] = b[0]; keys[5] = b[1];
}

static int make_args(crb_handle* h, uint64_t seed, uint32_t epoch, int32_t neg_ratio, int kind, SamplerArgs* a) {
    if (!h || !h->pos_user) {
        crb_set_error("sampler called before crb_set_history");
        return CRB_ERR_STATE;
    }
    CRB_CHECK_ARG(neg_ratio >= 1 && neg_ratio <= CRB_MAX_NEG, "neg_ratio must be in [1,64]");
    host_perm_keys(seed, epoch, a->keys);
    a->group = kind == 0 ? (uint32_t)neg_ratio : (kind == 1 ? (uint32_t)neg_ratio + 1u : 1u);
    a->n_rows = (uint64_t)h->n_pos * a->group;
    uint32_t bits = 2;
    while (((uint64_t)1 << bits) < a->n_rows) bits += 2;
    a->half_bits = bits / 2;
    a->half_mask = (uint32_t)(((uint64_t)1 << a->half_bits) - 1);
    a->neg_ratio = (uint32_t)neg_ratio;
    a->k0 = (uint32_t)seed;
    a->k1 = (uint32_t)(seed >> 32);
    a->epoch = epoch;
    uint32_t m = (uint32_t)(h->n_items > 0 ? h->n_items - 1 : 0);
    m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16;
    a->item_mask = m;
    a->n_items = (int32_t)h->n_items;
    a->pos_user = h->pos_user;
    a->pos_item = h->pos_item;
    a->seen_rowptr = h->seen_rowptr;
    a->seen_cols = h->seen_cols;
    a->bloom = h->bloom;
    a->bloom_shift = h->bloom_shift;
    a->bloom_exact = h->bloom_exact;
    a->bloom_stride = h->bloom_stride;
    a->dyn = h->dyn_mode ? h->dyn_dev : nullptr;
    return CRB_OK;
}

static int sampler_grid(crb_handle* h, int64_t count) {
    int64_t blocks = (count + 255) / 256;
    int64_t cap = (int64_t)h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int crb_launch_sample_pairwise(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count, int32_t neg_ratio,
                               int32_t* u, int32_t* i, int32_t* j, int32_t* nbr, bool count_rows, cudaStream_t s) {
    SamplerArgs a;
    int rc = make_args(h, seed, epoch, neg_ratio, 0, &a);
    if (rc) return rc;
    CRB_CHECK_ARG(first >= 0 && count >= 0 && (uint64_t)(first + count) <= a.n_rows, "rows outside the epoch");
    if (count == 0) return CRB_OK;
    int grid = sampler_grid(h, count);
    if (count_rows && h->dyn_mode)
        sample_kernel<0, true, true><<<grid, 256, 0, s>>>(a, first, count, u, i, j, nullptr, nbr, h->meta[0], h->meta[1], h->rank[0],
                                                          h->rank[1], h->rank[2], h->ctr);
    else if (count_rows)
        sample_kernel<0, true><<<grid, 256, 0, s>>>(a, first, count, u, i, j, nullptr, nbr, h->meta[0], h->meta[1], h->rank[0],
                                                    h->rank[1], h->rank[2], h->ctr);
    else
        sample_kernel<0, false><<<grid, 256, 0, s>>>(a, first, count, u, i, j, nullptr, nbr, nullptr, nullptr, nullptr, nullptr,
                                                     nullptr, h->ctr);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

static int check_outputs_device(const void* a, const void* b, const void* c) {
    if (!crb_is_device_ptr(a) || !crb_is_device_ptr(b) || !crb_is_device_ptr(c)) {
        crb_set_error("sampler outputs must be device pointers");
        return CRB_ERR_ARG;
    }
    return CRB_OK;
}

static int sampler_epilogue(crb_handle* h, cudaStream_t s) {
    // the attempt guard is the only run-time failure: surface it synchronously (cheap: 4 bytes)
    unsigned int err = 0;
    CRB_CUDA(cudaMemcpyAsync(&err, &h->ctr->sampler_err, sizeof(err), cudaMemcpyDeviceToHost, s));
    CRB_CUDA(cudaStreamSynchronize(s));
    if (err) {
        CRB_CUDA(cudaMemsetAsync(&h->ctr->sampler_err, 0, sizeof(unsigned int), s));
        crb_set_error("sampler: %u rows found no admissible negative (history covers the catalogue)", err);
        return CRB_ERR_SAMPLER;
    }
    return CRB_OK;
}

extern "C" int crb_sample_pairwise(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count,
                                   int32_t neg_ratio, int32_t* u, int32_t* i, int32_t* j, int32_t* nbr, void* stream) {
    CRB_CHECK_ARG(h, "null handle");
    if (count == 0) return CRB_OK;
    int rc = check_outputs_device(u, i, j);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    rc = crb_launch_sample_pairwise(h, seed, epoch, first, count, neg_ratio, u, i, j, nbr, false, s);
    if (rc) return rc;
    return sampler_epilogue(h, s);
}

// pointwise rows [first, first+count) into (u, i, y); count_rows also bumps the per-row multiplicities (K1 of a fused epoch loop)
int crb_launch_sample_pointwise(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count, int32_t neg_ratio, int32_t* u,
                                int32_t* i, float* y, bool count_rows, cudaStream_t s) {
    SamplerArgs a;
    int rc = make_args(h, seed, epoch, neg_ratio, 1, &a);
    if (rc) return rc;
    CRB_CHECK_ARG(first >= 0 && count >= 0 && (uint64_t)(first + count) <= a.n_rows, "rows outside the epoch");
    if (count == 0) return CRB_OK;
    if (count_rows)
        sample_kernel<1, true><<<sampler_grid(h, count), 256, 0, s>>>(a, first, count, u, i, nullptr, y, nullptr, h->meta[0], h->meta[1], h->rank[0],
                                                                      h->rank[1], nullptr, h->ctr);
    else
        sample_kernel<1, false><<<sampler_grid(h, count), 256, 0, s>>>(a, first, count, u, i, nullptr, y, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                       nullptr, h->ctr);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

extern "C" int crb_sample_pointwise(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count,
                                    int32_t neg_ratio, int32_t* u, int32_t* i, float* y, int32_t* nbr, void* stream) {
    CRB_CHECK_ARG(h, "null handle");
    if (count == 0) return CRB_OK;
    int rc = check_outputs_device(u, i, y);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    SamplerArgs a;
    rc = make_args(h, seed, epoch, neg_ratio, 1, &a);
    if (rc) return rc;
    CRB_CHECK_ARG(first >= 0 && count >= 0 && (uint64_t)(first + count) <= a.n_rows, "rows outside the epoch");
    if (count == 0) return CRB_OK;
    sample_kernel<1, false><<<sampler_grid(h, count), 256, 0, s>>>(a, first, count, u, i, nullptr, y, nbr, nullptr, nullptr,
                                                                   nullptr, nullptr, nullptr, h->ctr);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return sampler_epilogue(h, s);
}

extern "C" int crb_sample_cml(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count, int32_t neg_ratio,
                              int32_t* u, int32_t* i, int32_t* neg, void* stream) {
    CRB_CHECK_ARG(h, "null handle");
    if (count == 0) return CRB_OK;
    int rc = check_outputs_device(u, i, neg);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    SamplerArgs a;
    rc = make_args(h, seed, epoch, neg_ratio, 2, &a);
    if (rc) return rc;
    CRB_CHECK_ARG(first >= 0 && count >= 0 && (uint64_t)(first + count) <= a.n_rows, "rows outside the epoch");
    if (count == 0) return CRB_OK;
    sample_kernel<2, false><<<sampler_grid(h, count), 256, 0, s>>>(a, first, count, u, i, neg, nullptr, nullptr, nullptr, nullptr,
                                                                   nullptr, nullptr, nullptr, h->ctr);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return sampler_epilogue(h, s);
}

extern "C" int64_t crb_epoch_rows(crb_handle* h, int32_t neg_ratio, int32_t sampler_kind) {
    if (!h) return -1;
    if (sampler_kind == 0) return h->n_pos * neg_ratio;
    if (sampler_kind == 1) return h->n_pos * (neg_ratio + 1);
    if (sampler_kind == 3) return h->sp_n_pos * neg_ratio;   // SBPR: positives of users with a non-empty SPu only
    return h->n_pos;
}

int crb_launch_sample_nais(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t pos_first, int32_t n_pos_user, int32_t neg_ratio,
                           int32_t* targets, float* y, cudaStream_t s) {
    SamplerArgs a;
    int rc = make_args(h, seed, epoch, neg_ratio, 2, &a);
    if (rc) return rc;
    CRB_CHECK_ARG(pos_first >= 0 && n_pos_user >= 1 && pos_first + n_pos_user <= h->n_pos, "positives outside the history");
    sample_nais_kernel<<<(n_pos_user + 127) / 128, 128, 0, s>>>(a, pos_first, n_pos_user, targets, y, h->ctr);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

extern "C" int crb_sample_nais(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t pos_first, int32_t n_pos_user, int32_t neg_ratio,
                               int32_t* targets, float* y, void* stream) {
    CRB_CHECK_ARG(h, "null handle");
    CRB_CHECK_ARG(crb_is_device_ptr(targets) && crb_is_device_ptr(y), "sampler outputs must be device pointers");
    cudaStream_t s = (cudaStream_t)stream;
    int rc = crb_launch_sample_nais(h, seed, epoch, pos_first, n_pos_user, neg_ratio, targets, y, s);
    if (rc) return rc;
    return sampler_epilogue(h, s);
}

// ------------------------------------------------------------------------------------------------ SBPR (utils/sampler.py:102-141)
// Per positive (u, i) of a user with a non-empty SPu, neg_ratio rows: a social item k uniform over SPu[u] and a negative j uniform
// over the items that are neither u's nor in SPu[u]; rows are independent (no distinctness inside a group, :113-121), s_uk = number
// of u's friends that consumed k (:124-131, precomputed per SPu entry).  Spec shared with oracle/philox.py::sample_sbpr: row r of
// the epoch (before the shuffle) draws from the Philox blocks with counter (r_lo, r_hi, 0x80000000 | blk, epoch); words in order;
// the first word w with (w & mask(len(SPu[u]) - 1)) < len(SPu[u]) picks k, every later word is a negative candidate under
// np.random.randint's masked rejection and the membership test.
struct SbprSamplerArgs {
    SamplerArgs base;       // keys / bits / n_rows / group / item mask; pos_* and seen_* replaced by the social structures
    const int64_t* spu_start;
    const int32_t* spu_items;
    const float* spu_suk;
};

__global__ void __launch_bounds__(256) sample_sbpr_kernel(SbprSamplerArgs A, int64_t first, int64_t count, int32_t* __restrict__ ou,
                                                          int32_t* __restrict__ oi, int32_t* __restrict__ ok_, int32_t* __restrict__ oj,
                                                          float* __restrict__ osuk, crb_step_ctr* ctr) {
    const SamplerArgs& a = A.base;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const uint64_t s = feistel_perm((uint64_t)(first + t), a);
        const uint64_t p = s / a.group;
        const int32_t u = a.pos_user[p];
        const int64_t sp0 = A.spu_start[u];
        const uint32_t n_sp = (uint32_t)(A.spu_start[u + 1] - sp0);
        uint32_t sm = n_sp - 1;
        sm |= sm >> 1; sm |= sm >> 2; sm |= sm >> 4; sm |= sm >> 8; sm |= sm >> 16;
        const int64_t lo = a.seen_rowptr[u], hi = a.seen_rowptr[u + 1];
        int64_t pick = -1;
        int32_t neg = -1;
        for (uint32_t blk = 0; blk < CRB_SAMPLER_MAX_BLOCKS && neg < 0; ++blk) {
            uint32_t w[4];
            philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), 0x80000000u | blk, a.epoch, a.k0, a.k1, w);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (neg >= 0) continue;
                if (pick < 0) {
                    const uint32_t c = w[q] & sm;
                    if (c < n_sp) pick = c;
                    continue;
                }
                const int32_t v = (int32_t)(w[q] & a.item_mask);
                if (v >= a.n_items) continue;
                if (history_contains(a.seen_cols, lo, hi, v)) continue;
                neg = v;
            }
        }
        const bool ok = neg >= 0;
        ou[t] = u; oi[t] = a.pos_item[p];
        ok_[t] = ok ? A.spu_items[sp0 + pick] : 0;
        oj[t] = ok ? neg : 0;
        if (osuk) osuk[t] = ok ? A.spu_suk[sp0 + pick] : 1.f;
        if (!ok) atomicAdd(&ctr->sampler_err, 1u);
    }
}

extern "C" int crb_set_social(crb_handle* h, int64_t n_sp_pos, const int32_t* sp_pos_user, const int32_t* sp_pos_item, const int64_t* spu_start,
                              const int32_t* spu_items, const float* spu_suk, const int64_t* excl_rowptr, const int32_t* excl_cols) {
    CRB_CHECK_ARG(h, "null handle");
    if (!h->pos_user) { crb_set_error("crb_set_social before crb_set_history"); return CRB_ERR_STATE; }
    CRB_CHECK_ARG(n_sp_pos >= 0 && spu_start && excl_rowptr, "null argument");
    CRB_CHECK_ARG(n_sp_pos == 0 || (sp_pos_user && sp_pos_item && spu_items && spu_suk && excl_cols), "null argument");
    h->sp_n_pos = n_sp_pos; h->sp_pos_user = sp_pos_user; h->sp_pos_item = sp_pos_item; h->spu_start = spu_start; h->spu_items = spu_items;
    h->spu_suk = spu_suk; h->excl_rowptr = excl_rowptr; h->excl_cols = excl_cols;
    return CRB_OK;
}

extern "C" int crb_sample_sbpr(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count, int32_t neg_ratio, int32_t* u,
                               int32_t* i, int32_t* k, int32_t* j, float* suk, void* stream) {
    CRB_CHECK_ARG(h, "null handle");
    if (!h->spu_start) { crb_set_error("crb_sample_sbpr before crb_set_social"); return CRB_ERR_STATE; }
    if (count == 0) return CRB_OK;
    int rc = check_outputs_device(u, i, k);
    if (rc) return rc;
    if ((rc = check_outputs_device(j, j, suk ? (const void*)suk : (const void*)j))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    SbprSamplerArgs A;
    if ((rc = make_args(h, seed, epoch, neg_ratio, 0, &A.base))) return rc;
    // the epoch runs over the social positives only, rejection is against own + social items
    A.base.n_rows = (uint64_t)h->sp_n_pos * A.base.group;
    uint32_t bits = 2;
    while (((uint64_t)1 << bits) < A.base.n_rows) bits += 2;
    A.base.half_bits = bits / 2;
    A.base.half_mask = (uint32_t)(((uint64_t)1 << A.base.half_bits) - 1);
    A.base.pos_user = h->sp_pos_user; A.base.pos_item = h->sp_pos_item; A.base.seen_rowptr = h->excl_rowptr; A.base.seen_cols = h->excl_cols;
    A.base.bloom = nullptr; A.base.bloom_shift = 0; A.base.bloom_exact = 0; A.base.bloom_stride = 0;   // the filter covers the history, not the own + social exclusion sets
    A.spu_start = h->spu_start; A.spu_items = h->spu_items; A.spu_suk = h->spu_suk;
    CRB_CHECK_ARG(first >= 0 && count >= 0 && (uint64_t)(first + count) <= A.base.n_rows, "rows outside the epoch");
    sample_sbpr_kernel<<<sampler_grid(h, count), 256, 0, s>>>(A, first, count, u, i, k, j, suk, h->ctr);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return sampler_epilogue(h, s);
}
