// Sampled-candidate evaluation in ONE kernel (reference: test_model_loo, model/RankingRecommender.py:250-299):
//   pre_scores = sess.run(self.pre_scores, {u_idx, i_idx})      :257-278   every (user, candidate) pair of a user's segment
//   args_u     = np.argsort(-pre_scores_u)[:topk[-1]]           :281-288   the K best positions inside the segment
// A warp owns a user.  Its candidates are scored 32 at a time (one pair per lane, the canonical sequential fp32 fma chain of
// score_common.cuh, so every score is bit-identical to crb_score_pairs and to oracle/crb_oracle.c) and the warp keeps the K best
// (score, position) keys in registers, one per lane -- the score vector never reaches HBM.  Bound: HBM, the gathered item rows:
// (1 + neg_samples) * 4 * d bytes per user.
//
// Staging: a stage is a 64-column slice of the batch's 32 item rows (8 KB) brought in with 16-byte cp.async (LDGSTS, no
// registers); two stages per warp, the next one requested before the current one is consumed, so a warp always has 8 KB in flight.
// Row r's 16-byte chunk c sits at chunk position c ^ (r & 7): the eight lanes of a quarter-warp that walk eight different rows at
// the same column hit eight different bank groups (conflict-free LDS.128) without padding, which keeps 16-byte alignment.
#include <cuda_pipeline.h>

#include "score_common.cuh"

#define LT_CH 64                 // columns per stage
#define LT_CH4 (LT_CH / 4)       // 16-byte chunks per staged row
#define LT_WARPS 4
#define LT_MAX_DIM 512

struct LooArgs {
    const float* P;
    const float* Q;
    const float* hvec;
    int dim;
    const int32_t* seg_user;   // [n_users] row of P of each segment
    const int32_t* it;         // [total] candidate item ids, segments back to back
    const int64_t* offsets;    // [n_users + 1]
    int64_t n_users;
    int K;
    int ascending;
    int32_t* out_pos;          // [n_users, K] positions inside the segment, best first, -1 padded
    float* out_scores;         // optional [n_users, K]
};

// request stage `st` (batch b = st / n_chunks, columns kc = (st % n_chunks) * LT_CH) of the segment into `dst`
__device__ __forceinline__ void loo_issue(float4* dst, const float* __restrict__ Q, int32_t item_of_lane, int dim, int kc, int lane) {
    const int chunk = lane & (LT_CH4 - 1), half = lane >> 4;
    const bool in = kc + chunk * 4 < dim;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int row = 2 * q + half;
        const int32_t item = __shfl_sync(0xffffffffu, item_of_lane, row);
        if (in) __pipeline_memcpy_async(dst + row * LT_CH4 + (chunk ^ (row & 7)), Q + (int64_t)item * dim + kc + chunk * 4, 16);
    }
    __pipeline_commit();
}

template <int KIND>
__global__ void __launch_bounds__(LT_WARPS * 32) loo_topk_kernel(LooArgs a) {
    extern __shared__ float4 lt_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_f4 = a.dim / 4;                                   // dim % 4 == 0 (checked by the launcher)
    float4* sQ = lt_sm + (size_t)warp * (2 * 32 * LT_CH4 + LT_MAX_DIM / 4);
    float4* sP = sQ + 2 * 32 * LT_CH4;
    const int n_chunks = (a.dim + LT_CH - 1) / LT_CH;
    const int64_t gw = (int64_t)blockIdx.x * LT_WARPS + warp, nw = (int64_t)gridDim.x * LT_WARPS;
    for (int64_t usr = gw; usr < a.n_users; usr += nw) {
        const int64_t lo = a.offsets[usr], hi = a.offsets[usr + 1];
        const int64_t n = hi - lo;
        unsigned long long lst = 0ULL;                              // lane r: the r-th best key so far (0 = none)
        if (n > 0) {
            const float* prow = a.P + (int64_t)a.seg_user[usr] * a.dim;
            for (int c = lane; c < row_f4; c += 32) sP[c] = *reinterpret_cast<const float4*>(prow + 4 * c);
            const int64_t n_batches = (n + 31) / 32;
            const int64_t n_stages = n_batches * n_chunks;
            auto item_at = [&](int64_t b) { const int64_t p = b * 32 + lane; return a.it[lo + (p < n ? p : n - 1)]; };
            int32_t item_cur = item_at(0);
            int32_t item_nxt = n_batches > 1 ? item_at(1) : item_cur;
            loo_issue(sQ, a.Q, item_cur, a.dim, 0, lane);
            float acc = 0.f;
            for (int64_t st = 0; st < n_stages; ++st) {
                const int64_t b = st / n_chunks;
                const int ch = (int)(st - b * n_chunks);
                // request the next stage (same batch, next columns -- or the next batch's first columns) into the other buffer
                if (st + 1 < n_stages) {
                    const bool same = ch + 1 < n_chunks;
                    loo_issue(sQ + ((st + 1) & 1) * 32 * LT_CH4, a.Q, same ? item_cur : item_nxt, a.dim, same ? (ch + 1) * LT_CH : 0, lane);
                    __pipeline_wait_prior(1);
                } else {
                    __pipeline_wait_prior(0);
                }
                __syncwarp();
                const float4* q = sQ + (st & 1) * 32 * LT_CH4 + lane * LT_CH4;
                const int kc = ch * LT_CH;
                const int len4 = (a.dim - kc < LT_CH ? a.dim - kc : LT_CH) / 4;
                const int sw = lane & 7;
#pragma unroll 4
                for (int c = 0; c < len4; ++c) {
                    const float4 pv = sP[kc / 4 + c], qv = q[c ^ sw];
                    if (KIND == CRB_SCORE_DOT || KIND == CRB_SCORE_DOT_BIAS) {
                        acc = fmaf(pv.x, qv.x, acc); acc = fmaf(pv.y, qv.y, acc); acc = fmaf(pv.z, qv.z, acc); acc = fmaf(pv.w, qv.w, acc);
                    } else if (KIND == CRB_SCORE_GMF) {
                        const float4 hh = *reinterpret_cast<const float4*>(a.hvec + kc + 4 * c);
                        acc = fmaf(__fmul_rn(pv.x, qv.x), hh.x, acc); acc = fmaf(__fmul_rn(pv.y, qv.y), hh.y, acc);
                        acc = fmaf(__fmul_rn(pv.z, qv.z), hh.z, acc); acc = fmaf(__fmul_rn(pv.w, qv.w), hh.w, acc);
                    } else {
                        const float d0 = __fsub_rn(pv.x, qv.x), d1 = __fsub_rn(pv.y, qv.y), d2 = __fsub_rn(pv.z, qv.z), d3 = __fsub_rn(pv.w, qv.w);
                        acc = fmaf(d0, d0, acc); acc = fmaf(d1, d1, acc); acc = fmaf(d2, d2, acc); acc = fmaf(d3, d3, acc);
                    }
                }
                __syncwarp();   // the buffer is free for the stage after next
                if (ch + 1 < n_chunks) continue;
                // batch complete: this lane's score -> key -> the warp's running top-K
                if (KIND == CRB_SCORE_DOT_BIAS) acc = __fadd_rn(acc, a.hvec[item_cur]);
                const int64_t p = b * 32 + lane;
                unsigned long long key = p < n ? rank_key(acc, (uint32_t)p, a.ascending) : 0ULL;
                acc = 0.f;
                item_cur = item_nxt;
                if (b + 2 < n_batches) item_nxt = item_at(b + 2);
                unsigned long long thr = __shfl_sync(0xffffffffu, lst, a.K - 1);
                unsigned pass = __ballot_sync(0xffffffffu, key > thr);
                while (pass) {
                    const int src = __ffs(pass) - 1;
                    pass &= pass - 1;
                    const unsigned long long nk = __shfl_sync(0xffffffffu, key, src);
                    if (nk <= thr) continue;                                   // warp-uniform
                    const int pos = __popc(__ballot_sync(0xffffffffu, lane < a.K && lst > nk));
                    const unsigned long long up = __shfl_up_sync(0xffffffffu, lst, 1);
                    if (lane == pos) lst = nk; else if (lane > pos && lane < a.K) lst = up;
                    thr = __shfl_sync(0xffffffffu, lst, a.K - 1);
                }
            }
        }
        if (lane < a.K) {
            a.out_pos[usr * a.K + lane] = lst ? (int32_t)key_index(lst) : -1;
            if (a.out_scores) a.out_scores[usr * a.K + lane] = lst ? key_score(lst, a.ascending) : 0.f;
        }
        __syncwarp();   // sP / sQ are reused by the warp's next user
    }
}

int crb_launch_loo_topk(crb_handle* h, int32_t kind, const LooArgs& a, cudaStream_t s) {
    const size_t smem = sizeof(float4) * LT_WARPS * (2 * 32 * LT_CH4 + LT_MAX_DIM / 4);
    int64_t grid = (a.n_users + LT_WARPS - 1) / LT_WARPS;
    if (grid > (int64_t)h->sm_count * 3) grid = (int64_t)h->sm_count * 3;
#define CRB_LOO(KK)                                                                                                     \
    CRB_CUDA(cudaFuncSetAttribute(loo_topk_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    loo_topk_kernel<KK><<<(int)grid, LT_WARPS * 32, smem, s>>>(a);
    switch (kind) {
        case CRB_SCORE_DOT: { CRB_LOO(CRB_SCORE_DOT) } break;
        case CRB_SCORE_GMF: { CRB_LOO(CRB_SCORE_GMF) } break;
        case CRB_SCORE_SQDIST: { CRB_LOO(CRB_SCORE_SQDIST) } break;
        case CRB_SCORE_DOT_BIAS: { CRB_LOO(CRB_SCORE_DOT_BIAS) } break;
        default: crb_set_error("unknown score kind %d", kind); return CRB_ERR_ARG;
    }
#undef CRB_LOO
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

// One call for test_model_loo's scoring + ranking: seg_user [n_users] (row of P per segment), items [offsets[n_users]] candidate ids,
// offsets [n_users + 1]; -> topk_pos [n_users, K] positions inside each segment (np.argsort(-scores_u)[:K]; ascending for distance
// models), -1 padded.  All buffers DEVICE or HOST.  K <= 32 and dim % 4 == 0, dim <= 512 run the fused kernel; anything else takes the
// two-kernel route (crb_score_pairs + crb_topk_segments) inside the library -- same results.
extern "C" int crb_score_pairs_topk(crb_handle* h, int32_t kind, const float* P, const float* Q, const float* hvec, int32_t dim,
                                    const int32_t* seg_user, const int32_t* items, const int64_t* offsets, int64_t n_users, int32_t K,
                                    int32_t ascending, int32_t* topk_pos, void* stream) {
    CRB_CHECK_ARG(h && P && Q && seg_user && items && offsets && topk_pos, "null argument");
    CRB_CHECK_ARG(dim > 0 && n_users >= 0 && K >= 1, "sizes");
    CRB_CHECK_ARG(kind == CRB_SCORE_DOT || kind == CRB_SCORE_SQDIST || hvec, "this score kind needs hvec");
    if (n_users == 0) return CRB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CUDA(cudaSetDevice(h->device));
    const bool all_dev = crb_is_device_ptr(seg_user) && crb_is_device_ptr(items) && crb_is_device_ptr(offsets) && crb_is_device_ptr(topk_pos);
    const bool fused = K <= 32 && (dim & 3) == 0 && dim <= LT_MAX_DIM && (((uintptr_t)Q | (uintptr_t)P) & 15) == 0;
    if (!fused || !all_dev) {
        // host buffers or an unsupported shape: expand to pairs and use the two entry points (they stage host memory themselves)
        int64_t total = 0;
        if (crb_is_device_ptr(offsets)) {
            CRB_CUDA(cudaMemcpyAsync(&total, offsets + n_users, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
            CRB_CUDA(cudaStreamSynchronize(s));
        } else {
            total = offsets[n_users];
        }
        if (fused) {
            // stage everything on the device, run the fused kernel, copy the positions back
            void* tmp = nullptr;
            const int64_t bytes = 4 * n_users + 4 * total + 8 * (n_users + 1) + 4 * n_users * K + 1024;
            CRB_CUDA(cudaMallocAsync(&tmp, bytes, s));
            char* p = (char*)tmp;
            auto put = [&](const void* src, int64_t nbytes) -> void* {
                void* d = p;
                p += (nbytes + 255) & ~(int64_t)255;
                if (src) cudaMemcpyAsync(d, src, nbytes, crb_is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s);
                return d;
            };
            LooArgs a;
            a.P = P; a.Q = Q; a.hvec = hvec; a.dim = dim; a.n_users = n_users; a.K = K; a.ascending = ascending; a.out_scores = nullptr;
            a.seg_user = (const int32_t*)put(seg_user, 4 * n_users);
            a.it = (const int32_t*)put(items, 4 * total);
            a.offsets = (const int64_t*)put(offsets, 8 * (n_users + 1));
            a.out_pos = (int32_t*)put(nullptr, 4 * n_users * K);
            int rc = crb_launch_loo_topk(h, kind, a, s);
            if (!rc) {
                cudaError_t e = cudaMemcpyAsync(topk_pos, a.out_pos, 4 * n_users * K, crb_is_device_ptr(topk_pos) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s);
                if (e != cudaSuccess) { crb_set_error("copy back: %s", cudaGetErrorString(e)); rc = CRB_ERR_CUDA; }
            }
            cudaFreeAsync(tmp, s);
            if (rc) return rc;
            CRB_CUDA(cudaStreamSynchronize(s));
            return CRB_OK;
        }
        crb_set_error("crb_score_pairs_topk: K <= 32, dim %% 4 == 0, dim <= 512 and 16-byte aligned tables are required (got K=%d, dim=%d); "
                      "use crb_score_pairs + crb_topk_segments", K, dim);
        return CRB_ERR_UNSUPPORTED;
    }
    LooArgs a;
    a.P = P; a.Q = Q; a.hvec = hvec; a.dim = dim; a.seg_user = seg_user; a.it = items; a.offsets = offsets; a.n_users = n_users;
    a.K = K; a.ascending = ascending; a.out_pos = topk_pos; a.out_scores = nullptr;
    return crb_launch_loo_topk(h, kind, a, s);
}
