// Canonical score arithmetic and ranking keys shared by the evaluation kernels.
#pragma once
#include "common.cuh"

// Canonical score of one (user vector, item row) pair: ONE sequential fp32 fma chain over k = 0..dim-1.
// oracle/crb_oracle.c restates exactly this, which is what makes top-K ids and ranks bit-exact.
template <int KIND>
__device__ __forceinline__ float canonical_score(const float* __restrict__ p, const float* __restrict__ q,
                                                 const float* __restrict__ hvec, int32_t item, int dim) {
    float acc = 0.f;
    if ((dim & 3) == 0) {
        for (int k = 0; k < dim; k += 4) {
            const float4 a = *reinterpret_cast<const float4*>(p + k);
            const float4 b = *reinterpret_cast<const float4*>(q + k);
            if (KIND == CRB_SCORE_DOT || KIND == CRB_SCORE_DOT_BIAS) {
                acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
            } else if (KIND == CRB_SCORE_GMF) {
                const float4 hh = *reinterpret_cast<const float4*>(hvec + k);
                acc = fmaf(__fmul_rn(a.x, b.x), hh.x, acc); acc = fmaf(__fmul_rn(a.y, b.y), hh.y, acc);
                acc = fmaf(__fmul_rn(a.z, b.z), hh.z, acc); acc = fmaf(__fmul_rn(a.w, b.w), hh.w, acc);
            } else {
                float d0 = __fsub_rn(a.x, b.x), d1 = __fsub_rn(a.y, b.y), d2 = __fsub_rn(a.z, b.z), d3 = __fsub_rn(a.w, b.w);
                acc = fmaf(d0, d0, acc); acc = fmaf(d1, d1, acc); acc = fmaf(d2, d2, acc); acc = fmaf(d3, d3, acc);
            }
        }
    } else {
        for (int k = 0; k < dim; ++k) {
            if (KIND == CRB_SCORE_DOT || KIND == CRB_SCORE_DOT_BIAS) acc = fmaf(p[k], q[k], acc);
            else if (KIND == CRB_SCORE_GMF) acc = fmaf(__fmul_rn(p[k], q[k]), hvec[k], acc);
            else { float d = __fsub_rn(p[k], q[k]); acc = fmaf(d, d, acc); }
        }
    }
    if (KIND == CRB_SCORE_DOT_BIAS) acc = __fadd_rn(acc, hvec[item]);
    return acc;
}

// 64-bit ranking key, larger = better: (order-preserving score bits, ~index).  Implements the documented tie rule
// (score descending -- ascending for distance models --, index ascending).  0 = "no candidate".
__host__ __device__ __forceinline__ unsigned long long rank_key(float score, uint32_t idx, int ascending) {
    if (score == 0.f) score = 0.f;  // -0 -> +0 so both compare equal like NumPy
    uint32_t f;
    memcpy(&f, &score, 4);
    uint32_t ord = (f & 0x80000000u) ? ~f : (f | 0x80000000u);
    if (ascending) ord = ~ord;
    return ((unsigned long long)ord << 32) | (unsigned long long)(~idx);
}
__host__ __device__ __forceinline__ uint32_t key_index(unsigned long long key) { return ~(uint32_t)key; }
__host__ __device__ __forceinline__ float key_score(unsigned long long key, int ascending) {
    uint32_t ord = (uint32_t)(key >> 32);
    if (ascending) ord = ~ord;
    uint32_t f = (ord & 0x80000000u) ? (ord & 0x7fffffffu) : ~ord;
    float s;
    memcpy(&s, &f, 4);
    return s;
}

__device__ __forceinline__ bool sorted_contains(const int32_t* __restrict__ cols, int64_t lo, int64_t hi, int32_t v) {
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int32_t c = __ldg(cols + mid);
        if (c < v) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(cols + lo) == v;
}

// In-place descending bitonic sort of buf[0..CAP) by a 256-thread CTA; entries >= count are treated as 0.
template <int CAP>
__device__ __forceinline__ void block_sort_desc(unsigned long long* buf, int count) {
    for (int k = threadIdx.x; k < CAP; k += blockDim.x)
        if (k >= count) buf[k] = 0ULL;
    __syncthreads();
    for (int size = 2; size <= CAP; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int k = threadIdx.x; k < CAP / 2; k += blockDim.x) {
                const int lo = 2 * k - (k & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = buf[lo], b = buf[hi];
                if ((a < b) == desc) { buf[lo] = b; buf[hi] = a; }
            }
            __syncthreads();
        }
    }
}

int crb_launch_fullrank_exact(crb_handle* h, int32_t kind, const float* P, const float* Q, const float* hvec, int64_t n_items,
                              int32_t dim, const int32_t* users, const int32_t* hist_users, const int32_t* todo, int64_t n_todo,
                              int32_t K, int32_t* out_items, float* out_scores, cudaStream_t s);
