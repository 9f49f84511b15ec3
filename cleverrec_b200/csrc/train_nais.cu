// NAIS_single (reference: model/ranking/NAIS_single.py:59-97, train_model_nais RankingRecommender.py:64-87): one optimizer
// step per user.  History H (n items, list order), targets T (m = n*(1+neg_ratio)), attention over j_tk = q_t * p_k (atten_type
// 'prod', W [d, A]) or j_tk = [p_k ; q_t] (atten_type 'concat', W [2d, A]; NAIS_single.py:67-71):
//   a_tk = h . relu(W^T j_tk + b);  e = exp(a);  w_tk = e_tk / (sum_k e_tk)^beta;  s_t = sum_k w_tk p_k;
//   x_t = s_t . q_t + bias_t;  loss = sum CE(x, y) + reg * (l2(s) + l2(q_T) + l2(bias_T))          (NAIS_single.py:66-90)
// Three kernels per step (+ dense applies):
//   nais_attn_kernel<false>   a_tk for every (target, history) pair, one warp per pair, W in shared memory
//   nais_target_kernel        per target: smoothed softmax, s_t, logit, loss, d(loss)/da_tk, gradients of P[H], Q[T], bias
//   nais_attn_kernel<true>    attention backward per pair: gradients of W, b, h (per-CTA shared accumulators), P[H], Q[T]
#include "rowopt.cuh"

#define NA_WARPS 8
#define NA_MAXV 16   // d <= 512

struct NaisArgs {
    const float* P;
    const float* Q;
    const float* bias;
    float* gP;
    float* gQ;
    float* gbias;
    const float* dense;    // packed W [d, A] row-major, b [A], h [A]
    float* dense_part;     // [grid, n_dense]
    const int32_t* hist;   // [n]
    const int32_t* tgt;    // [m]
    const float* y;        // [m]
    int n, m, d, A;
    int concat;            // 1: j_tk = [p_k ; q_t], W has 2d rows
    float beta, reg;
    float* abuf;           // [m, n]  a_tk, then d(loss)/da_tk
    float* wbuf;           // [m, n]  d(loss)/dw_tk scratch
    float* scores;         // [m] (scoring mode)
    double* loss_part;
};

__device__ __forceinline__ float wsum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// shared: W padded [d][A+1] | b[A] | h[A] | (BWD) gW padded | gb | gh | per-warp j[d] + dz[32]
template <bool BWD>
__global__ void __launch_bounds__(NA_WARPS * 32) nais_attn_kernel(NaisArgs a) {
    extern __shared__ float sm[];
    const int d = a.d, A = a.A, AP = A + 1;
    const int JD = a.concat ? 2 * d : d;          // rows of W = length of the joint vector
    float* sW = sm;
    float* sb = sW + JD * AP;
    float* sh = sb + A;
    float* gW = sh + A;
    float* gb = gW + (BWD ? JD * AP : 0);
    float* gh = gb + (BWD ? A : 0);
    float* wbase = gh + (BWD ? A : 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* jbuf = wbase + warp * (JD + 32);
    float* dzbuf = jbuf + JD;
    for (int k = threadIdx.x; k < JD * A; k += blockDim.x) sW[(k / A) * AP + (k % A)] = a.dense[k];
    for (int k = threadIdx.x; k < A; k += blockDim.x) { sb[k] = a.dense[JD * A + k]; sh[k] = a.dense[JD * A + A + k]; }
    if (BWD)
        for (int k = threadIdx.x; k < JD * AP + 2 * A; k += blockDim.x) gW[k] = 0.f;
    __syncthreads();
    const int64_t pairs = (int64_t)a.m * a.n;
    for (int64_t idx = (int64_t)blockIdx.x * NA_WARPS + warp; idx < pairs; idx += (int64_t)gridDim.x * NA_WARPS) {
        const int t = (int)(idx / a.n), k = (int)(idx % a.n);
        const int64_t qrow = a.tgt[t], prow = a.hist[k];
        float qv[NA_MAXV], pv[NA_MAXV];
#pragma unroll
        for (int v = 0; v < NA_MAXV; ++v) {
            const int c = lane + 32 * v;
            if (c < d) {
                qv[v] = a.Q[qrow * d + c];
                pv[v] = a.P[prow * d + c];
                if (a.concat) { jbuf[c] = pv[v]; jbuf[d + c] = qv[v]; }   // concat([tile(p), tile(q)], 2)   (:68-69)
                else jbuf[c] = qv[v] * pv[v];                             // einsum('ac,bc->abc', q, p)      (:71)
            }
        }
        __syncwarp();
        float pre = 0.f, z = 0.f;
        if (lane < A) {
            pre = sb[lane];
            for (int c = 0; c < JD; ++c) pre = fmaf(jbuf[c], sW[c * AP + lane], pre);
            z = fmaxf(pre, 0.f);
        }
        if (!BWD) {
            const float att = wsum(lane < A ? sh[lane] * z : 0.f);
            if (lane == 0) a.abuf[idx] = att;
        } else {
            const float da = a.abuf[idx];
            float dz = 0.f;
            if (lane < A) {
                dz = pre > 0.f ? da * sh[lane] : 0.f;
                atomicAdd(gh + lane, da * z);
                atomicAdd(gb + lane, dz);
                if (dz != 0.f)
                    for (int c = 0; c < JD; ++c) atomicAdd(gW + c * AP + lane, jbuf[c] * dz);
            }
            dzbuf[lane] = dz;
            __syncwarp();
#pragma unroll
            for (int v = 0; v < NA_MAXV; ++v) {
                const int c = lane + 32 * v;
                if (c < d) {
                    float dj = 0.f;
                    for (int q = 0; q < A; ++q) dj = fmaf(sW[c * AP + q], dzbuf[q], dj);
                    if (a.concat) {   // the joint vector holds p and q themselves: d/dp = dj[c], d/dq = dj[d + c]
                        float dq = 0.f;
                        for (int q = 0; q < A; ++q) dq = fmaf(sW[(d + c) * AP + q], dzbuf[q], dq);
                        atomicAdd(a.gP + prow * d + c, dj);
                        atomicAdd(a.gQ + qrow * d + c, dq);
                    } else {
                        atomicAdd(a.gQ + qrow * d + c, dj * pv[v]);
                        atomicAdd(a.gP + prow * d + c, dj * qv[v]);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (BWD) {
        __syncthreads();
        float* part = a.dense_part + (int64_t)blockIdx.x * (JD * A + 2 * A);
        for (int k = threadIdx.x; k < JD * A; k += blockDim.x) part[k] = gW[(k / A) * AP + (k % A)];
        for (int k = threadIdx.x; k < A; k += blockDim.x) { part[JD * A + k] = gb[k]; part[JD * A + A + k] = gh[k]; }
    }
}

// one warp per target
template <bool TRAIN>
__global__ void __launch_bounds__(NA_WARPS * 32) nais_target_kernel(NaisArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = a.d, n = a.n;
    double loss = 0.0;
    for (int t = blockIdx.x * NA_WARPS + warp; t < a.m; t += gridDim.x * NA_WARPS) {
        const int64_t qrow = a.tgt[t];
        float* arow = a.abuf + (int64_t)t * n;
        float E = 0.f;
        for (int k = lane; k < n; k += 32) E += expf(arow[k]);
        E = wsum(E);
        const float invD = powf(E, -a.beta);          // 1 / (sum e)^beta   (NAIS_single.py:77-78)
        float s[NA_MAXV], qv[NA_MAXV];
#pragma unroll
        for (int v = 0; v < NA_MAXV; ++v) { s[v] = 0.f; const int c = lane + 32 * v; qv[v] = c < d ? a.Q[qrow * d + c] : 0.f; }
        for (int k = 0; k < n; ++k) {
            const float w = expf(arow[k]) * invD;
            const int64_t prow = a.hist[k];
#pragma unroll
            for (int v = 0; v < NA_MAXV; ++v) { const int c = lane + 32 * v; if (c < d) s[v] = fmaf(w, a.P[prow * d + c], s[v]); }
        }
        float x = 0.f, ssq = 0.f, qsq = 0.f;
#pragma unroll
        for (int v = 0; v < NA_MAXV; ++v) { x = fmaf(s[v], qv[v], x); ssq = fmaf(s[v], s[v], ssq); qsq = fmaf(qv[v], qv[v], qsq); }
        const float bt = a.bias[qrow];
        x = wsum(x) + bt;
        if (!TRAIN) {
            if (lane == 0) a.scores[t] = x;
            continue;
        }
        ssq = wsum(ssq); qsq = wsum(qsq);
        const float y = a.y[t];
        const float g = sigmoid_f(x) - y;
        if (lane == 0) {
            loss += (double)(fmaxf(x, 0.f) - x * y + __logf(1.f + __expf(-fabsf(x))) + a.reg * 0.5f * (ssq + qsq + bt * bt));
            atomicAdd(a.gbias + qrow, g + a.reg * bt);
        }
        float ds[NA_MAXV];
#pragma unroll
        for (int v = 0; v < NA_MAXV; ++v) {
            const int c = lane + 32 * v;
            ds[v] = fmaf(g, qv[v], a.reg * s[v]);
            if (c < d) atomicAdd(a.gQ + qrow * d + c, fmaf(g, s[v], a.reg * qv[v]));
        }
        // dw_tk = ds . p_k ;  M = sum_k dw_tk e_tk ;  da_tk = e_tk (dw_tk E^-beta - beta M E^(-beta-1))
        float* wrow = a.wbuf + (int64_t)t * n;
        float M = 0.f;
        for (int k = 0; k < n; ++k) {
            const int64_t prow = a.hist[k];
            const float e = expf(arow[k]);
            float dw = 0.f;
#pragma unroll
            for (int v = 0; v < NA_MAXV; ++v) {
                const int c = lane + 32 * v;
                if (c < d) {
                    const float p = a.P[prow * d + c];
                    dw = fmaf(ds[v], p, dw);
                    atomicAdd(a.gP + prow * d + c, e * invD * ds[v]);   // through s_t = sum w p
                }
            }
            dw = wsum(dw);
            if (lane == 0) wrow[k] = dw;
            M = fmaf(dw, e, M);
        }
        __syncwarp();
        const float c2 = -a.beta * M * invD / E;
        for (int k = lane; k < n; k += 32) {
            const float e = expf(arow[k]);
            arow[k] = e * fmaf(wrow[k], invD, c2);
        }
    }
    if (TRAIN) {
        __shared__ double sl[NA_WARPS];
        for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
        if (lane == 0) sl[warp] = loss;
        __syncthreads();
        if (threadIdx.x == 0) {
            double ts = 0.0;
            for (int k = 0; k < NA_WARPS; ++k) ts += sl[k];
            a.loss_part[blockIdx.x] = ts;
        }
    }
}

__global__ void __launch_bounds__(256) nais_dense_apply_kernel(float* w, float* s1, float* s2, const float* parts, int n_parts, int n,
                                                              int opt_kind, OptDev o) {
    for (int k0 = blockIdx.x * 32; k0 < n; k0 += gridDim.x * 32) {
        const int k = k0 + (threadIdx.x & 31);
        const float g = block_sum_parts(parts, n_parts, n, k, k < n);
        if (threadIdx.x >= 32 || k >= n) continue;
        float x = w[k];
        if (opt_kind == OPT_SGD) {
            x -= o.lr * g;
        } else if (opt_kind == OPT_ADAGRAD) {
            float acc = s1[k];
            adagrad_elem(x, acc, g, o.lr);
            s1[k] = acc;
        } else {
            float m = s1[k], v = s2[k];
            adam_touch_elem(x, m, v, g, o);
            s1[k] = m; s2[k] = v;
        }
        w[k] = x;
    }
}

int crb_dense_table_apply(crb_handle* h, const crb_table* T, float* grad, int opt_kind, const OptDev& od, float l2, double* loss_part,
                          int* grid_out, cudaStream_t s);
int crb_dense_tables_apply(crb_handle* h, int n, const crb_table* const* tables, float* const* grads, int opt_kind, const OptDev& od, cudaStream_t s);
int crb_opt_to_dev(crb_handle* h, const crb_opt* opt, OptDev* out, int* opt_kind, cudaStream_t s);

static int nais_prepare(crb_handle* h, NaisArgs& a, const float* P, const float* Q, const float* bias, const float* dense, int32_t d, int32_t A,
                        int32_t concat, const int32_t* hist, int32_t n, const int32_t* tgt, int32_t m, float beta, int64_t* buf_floats, cudaStream_t s) {
    CRB_CHECK_ARG(d >= 1 && d <= 512 && A >= 1 && A <= 32, "NAIS needs embed_size <= 512 and atten_size <= 32");
    CRB_CHECK_ARG(n >= 1 && m >= 1, "empty history / target list");
    CRB_CHECK_ARG(crb_is_device_ptr(hist) && crb_is_device_ptr(tgt), "hist/targets must be device pointers");
    a.P = P; a.Q = Q; a.bias = bias; a.dense = dense; a.hist = hist; a.tgt = tgt; a.n = n; a.m = m; a.d = d; a.A = A; a.concat = concat ? 1 : 0; a.beta = beta;
    *buf_floats = 2 * (int64_t)m * n;
    (void)h; (void)s;
    return CRB_OK;
}

static size_t nais_smem(int d, int A, bool bwd, int concat) {
    const int JD = concat ? 2 * d : d;
    return sizeof(float) * ((size_t)(bwd ? 2 : 1) * (JD * (A + 1) + 2 * A) + (size_t)NA_WARPS * (JD + 32));
}

static int nais_ws(crb_handle* h, int64_t floats, cudaStream_t s) {
    h->evq_valid = 0;
    return crb_eval_ws_reserve(h, floats * 4 + 1024) ? CRB_ERR_CUDA : CRB_OK;
}

extern "C" int crb_train_step_nais(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_table* B, float* gradP, float* gradQ,
                                   float* gradB, float* dense, float* dense_s1, float* dense_s2, int32_t atten_size, int32_t atten_concat,
                                   const crb_opt* opt, const int32_t* hist, int32_t n_hist, const int32_t* targets, const float* y,
                                   int32_t n_targets, float beta, float reg, double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && P && Q && B && gradP && gradQ && gradB && dense && hist && targets && y, "null argument");
    CRB_CHECK_ARG(crb_is_device_ptr(y), "y must be a device pointer");
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    const int dk = opt_kind == OPT_ADAM_TF1 ? OPT_ADAM_LAZY : opt_kind;
    CRB_CHECK_ARG(dk == OPT_SGD || dense_s1, "dense slot s1 is NULL");
    CRB_CHECK_ARG(dk != OPT_ADAM_LAZY || dense_s2, "dense slot s2 is NULL");
    CRB_CHECK_ARG(B->dim == 1 && B->rows % 4 == 0, "bias table: dim 1, rows padded to a multiple of 4");
    CRB_CUDA(cudaSetDevice(h->device));
    NaisArgs a;
    int64_t bufs = 0;
    if ((rc = nais_prepare(h, a, P->w, Q->w, B->w, dense, P->dim, atten_size, atten_concat, hist, n_hist, targets, n_targets, beta, &bufs, s))) return rc;
    const int n_dense = (atten_concat ? 2 : 1) * P->dim * atten_size + 2 * atten_size;
    const int64_t pairs = (int64_t)n_targets * n_hist;
    int grid_p = (int)((pairs + NA_WARPS - 1) / NA_WARPS);
    if (grid_p > h->sm_count * 2) grid_p = h->sm_count * 2;
    int grid_t = (n_targets + NA_WARPS - 1) / NA_WARPS;
    if (grid_t > h->sm_count * 4) grid_t = h->sm_count * 4;
    if ((rc = nais_ws(h, bufs + (int64_t)grid_p * n_dense, s))) return rc;
    a.abuf = (float*)h->eval_ws;
    a.wbuf = a.abuf + pairs;
    a.dense_part = a.wbuf + pairs;
    a.gP = gradP; a.gQ = gradQ; a.gbias = gradB; a.y = y; a.reg = reg; a.scores = nullptr; a.loss_part = h->block_loss;
    const size_t sm_f = nais_smem(P->dim, atten_size, false, atten_concat), sm_b = nais_smem(P->dim, atten_size, true, atten_concat);
    if (sm_b > 200 * 1024) { crb_set_error("NAIS attention too large for shared memory"); return CRB_ERR_UNSUPPORTED; }
    CRB_CUDA(cudaFuncSetAttribute(nais_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_f));
    CRB_CUDA(cudaFuncSetAttribute(nais_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_b));
    if ((rc = crb_prof_begin(h, s))) return rc;
    nais_attn_kernel<false><<<grid_p, NA_WARPS * 32, sm_f, s>>>(a);
    nais_target_kernel<true><<<grid_t, NA_WARPS * 32, 0, s>>>(a);
    nais_attn_kernel<true><<<grid_p, NA_WARPS * 32, sm_b, s>>>(a);
    if ((rc = crb_prof_end(h, s))) return rc;
    h->launches += 3;
    CRB_CUDA(cudaGetLastError());
    // sparse TF applies on P, Q, bias == dense applies with zero gradient on untouched rows (see train_neumf.cu)
    crb_table Bt = *B;
    Bt.rows = B->rows / 4; Bt.dim = 4;
    const crb_table* tabs3[3] = {P, Q, &Bt};
    float* grads3[3] = {gradP, gradQ, gradB};
    if ((rc = crb_dense_tables_apply(h, 3, tabs3, grads3, dk, od, s))) return rc;
    nais_dense_apply_kernel<<<(n_dense + 31) / 32, 256, 0, s>>>(dense, dense_s1, dense_s2, a.dense_part, grid_p, n_dense, dk, od);
    h->launches++;
    h->step_grid = grid_t;
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    if ((rc = crb_ws_reserve(h, 1, 4, 4, s))) return rc;
    ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    if ((rc = crb_launch_loss_final(h, ld, s))) return rc;
    if (loss_out && !crb_is_device_ptr(loss_out)) {
        CRB_CUDA(cudaMemcpyAsync(loss_out, h->loss_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
        CRB_CUDA(cudaStreamSynchronize(s));
    }
    return CRB_OK;
}

extern "C" int crb_score_nais(crb_handle* h, const float* P, const float* Q, const float* bias, const float* dense, int32_t dim,
                              int32_t atten_size, int32_t atten_concat, const int32_t* hist, int32_t n_hist, const int32_t* targets, int32_t n_targets, float beta,
                              float* scores, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && P && Q && bias && dense && hist && targets && scores, "null argument");
    CRB_CHECK_ARG(crb_is_device_ptr(scores), "scores must be a device pointer");
    CRB_CUDA(cudaSetDevice(h->device));
    NaisArgs a;
    int64_t bufs = 0;
    int rc = nais_prepare(h, a, P, Q, bias, dense, dim, atten_size, atten_concat, hist, n_hist, targets, n_targets, beta, &bufs, s);
    if (rc) return rc;
    const int64_t pairs = (int64_t)n_targets * n_hist;
    if ((rc = nais_ws(h, pairs, s))) return rc;
    a.abuf = (float*)h->eval_ws; a.wbuf = nullptr; a.dense_part = nullptr; a.gP = a.gQ = a.gbias = nullptr; a.y = nullptr; a.reg = 0.f;
    a.scores = scores; a.loss_part = nullptr;
    int64_t grid_p = (pairs + NA_WARPS - 1) / NA_WARPS;
    if (grid_p > (int64_t)h->sm_count * 4) grid_p = (int64_t)h->sm_count * 4;
    int grid_t = (n_targets + NA_WARPS - 1) / NA_WARPS;
    if (grid_t > h->sm_count * 4) grid_t = h->sm_count * 4;
    const size_t sm_f = nais_smem(dim, atten_size, false, atten_concat);
    if (sm_f > 200 * 1024) { crb_set_error("NAIS attention too large for shared memory"); return CRB_ERR_UNSUPPORTED; }
    CRB_CUDA(cudaFuncSetAttribute(nais_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_f));
    nais_attn_kernel<false><<<(int)grid_p, NA_WARPS * 32, sm_f, s>>>(a);
    nais_target_kernel<false><<<grid_t, NA_WARPS * 32, 0, s>>>(a);
    h->launches += 2;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

int crb_launch_sample_nais(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t pos_first, int32_t n_pos_user, int32_t neg_ratio,
                           int32_t* targets, float* y, cudaStream_t s);

// train_model_nais for `n_users` users in one call: user k has its interaction list at pos_item[list_start[k] .. +list_len[k]) (HOST
// arrays, in the order the reference iterates data.ui_train); one sampler launch + one optimizer step per user; loss_out[k].
extern "C" int crb_train_epoch_nais(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_table* B, float* gradP, float* gradQ,
                                    float* gradB, float* dense, float* dense_s1, float* dense_s2, int32_t atten_size, int32_t atten_concat,
                                    const crb_opt* opt, uint64_t seed, uint32_t epoch, const int64_t* list_start, const int32_t* list_len, int64_t n_users,
                                    int32_t neg_ratio, float beta, float reg, double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && list_start && list_len && loss_out, "null argument");
    CRB_CHECK_ARG(!crb_is_device_ptr(list_start) && !crb_is_device_ptr(list_len), "list_start/list_len are HOST arrays here");
    CRB_CHECK_ARG(crb_is_device_ptr(loss_out), "loss_out must be a device array [n_users]");
    if (!h->pos_item) { crb_set_error("crb_train_epoch_nais before crb_set_history"); return CRB_ERR_STATE; }
    int32_t max_len = 0;
    for (int64_t k = 0; k < n_users; ++k) max_len = list_len[k] > max_len ? list_len[k] : max_len;
    CRB_CHECK_ARG(max_len >= 1, "empty epoch");
    int rc = crb_ws_reserve(h, (int64_t)max_len * (neg_ratio + 1), 4, 4, s);
    if (rc) return rc;
    crb_opt step_opt = *opt;
    for (int64_t k = 0; k < n_users; ++k) {
        const int32_t n = list_len[k];
        if (n < 1) { crb_set_error("user %lld has an empty interaction list", (long long)k); return CRB_ERR_ARG; }
        const int32_t m = n * (neg_ratio + 1);
        if ((rc = crb_launch_sample_nais(h, seed, epoch, list_start[k], n, neg_ratio, h->idx[0], h->yv, s))) return rc;
        step_opt.step = opt->step + k;
        rc = crb_train_step_nais(h, P, Q, B, gradP, gradQ, gradB, dense, dense_s1, dense_s2, atten_size, atten_concat, &step_opt,
                                 h->pos_item + list_start[k], n,
                                 h->idx[0], h->yv, m, beta, reg, loss_out + k, stream);
        if (rc) return rc;
    }
    return CRB_OK;
}
