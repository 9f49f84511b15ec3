// LRML: Latent Relational Metric Learning (reference: model/ranking/LRML.py:42-78).
//   LRAM (:42-51)   x = p_u * q          key = x K   (K: [d, mem])      a = softmax(key)      r = a M   (M: [mem, d])
//   distance (:59)  e = p_u + r - q      dist = |e|^2
//   loss (:62-63)   sum max(dist_ui - dist_uj + margin, 0) + reg * (l2(p_u) + l2(q_i) + l2(q_j))        (hinge: utils/tools.py:73)
// One warp per triplet.  K and M live in shared memory with odd row strides so that both the forward walks (lanes over
// columns) and the backward walks (lanes over rows) are bank-conflict free.  Gradients of K / M are accumulated per CTA in shared
// memory and leave as per-CTA partials that dense_vector_apply_kernel sums in CTA order; the embedding gradients go into the
// tables' dense gradient buffers (TF's sparse Adam decays the moments of every row and moves every row, SURVEY 2.4, which is
// exactly a dense apply with zero gradient on untouched rows; for SGD / Adagrad a zero gradient is a no-op).
// The pair scorer (LRML._predict :70-78) runs the same forward device function, so training and evaluation distances agree bit
// for bit.
#include "rowopt.cuh"

#define LR_WARPS 8
#define LR_MAX_D 256
#define LR_MAX_MEM 128

struct LrmlShape {
    int d, mem;
    int ks;   // row stride of the shared copy of K (odd)
    int ms;   // row stride of the shared copy of M (odd)
    int n_dense;   // packed: K [d, mem] row-major, then M [mem, d] row-major
    int per_warp;  // scratch floats per warp
};

static int lrml_shape(int d, int mem, LrmlShape* s) {
    if (d < 4 || d > LR_MAX_D || (d & 3) || mem < 1 || mem > LR_MAX_MEM) return -1;
    s->d = d; s->mem = mem;
    s->ks = mem | 1; s->ms = d | 1;
    s->n_dense = 2 * d * mem;
    // p, q_i, q_j, e_i, e_j  [d] each;  a_i, a_j, tmp [mem] each
    s->per_warp = 5 * d + 3 * mem;
    return 0;
}

__device__ __forceinline__ float lr_warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float lr_warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Forward of one (user, item) side by one warp.  vp / vq: the rows in shared memory.  Writes a[mem] and e[d]; returns dist.
// Canonical arithmetic: x = p * q rounded; key_m = fma chain over k ascending from 0; softmax with the row maximum subtracted
// (tf.nn.softmax), the sum taken lane-strided then by xor tree; r_k = fma chain over m ascending; e = (p + r) - q.
__device__ __forceinline__ float lrml_forward(const LrmlShape& S, const float* sK, const float* sM, const float* vp, const float* vq,
                                              float* a, float* e, int lane) {
    float mx = -INFINITY;
    for (int m = lane; m < S.mem; m += 32) {
        float acc = 0.f;
        for (int k = 0; k < S.d; ++k) acc = fmaf(__fmul_rn(vp[k], vq[k]), sK[k * S.ks + m], acc);
        a[m] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = lr_warp_max(mx);
    float sum = 0.f;
    for (int m = lane; m < S.mem; m += 32) {
        const float ex = __expf(a[m] - mx);
        a[m] = ex;
        sum += ex;
    }
    sum = lr_warp_sum(sum);
    const float inv = __fdividef(1.f, sum);
    for (int m = lane; m < S.mem; m += 32) a[m] *= inv;
    __syncwarp();
    float dist = 0.f;
    for (int k = lane; k < S.d; k += 32) {
        float r = 0.f;
        for (int m = 0; m < S.mem; ++m) r = fmaf(a[m], sM[m * S.ms + k], r);
        const float ev = (vp[k] + r) - vq[k];
        e[k] = ev;
        dist = fmaf(ev, ev, dist);
    }
    dist = lr_warp_sum(dist);
    __syncwarp();
    return dist;
}

struct LrmlArgs {
    LrmlShape sh;
    const float* P;
    const float* Q;
    float* gP;
    float* gQ;
    const float* dense;
    float* dense_part;   // [gridDim.x, n_dense]
    const int32_t* u;
    const int32_t* i;
    const int32_t* j;
    int64_t batch;
    float margin, reg;
    double* loss_part;
};

// backward of one side: coefficient c = dL/d dist (+1 for the positive, -1 for the negative of an active hinge).
// de = 2 c e;  dp += de, dq -= de, dr = de;  dM[m][k] += a[m] dr[k];  da[m] = M[m] . dr;  dkey = a * (da - a . da);
// dK[k][m] += x[k] dkey[m];  dx[k] = K[k] . dkey;  dp += dx * q;  dq += dx * p.   gp / gq: per-lane accumulators of the caller.
__device__ __forceinline__ void lrml_backward(const LrmlShape& S, const float* sK, const float* sM, float* gK, float* gM, const float* vp,
                                              const float* vq, const float* a, float* e, float* tm, float c, float* gp, float* gq,
                                              int lane) {
    const float c2 = 2.f * c;
    for (int k = lane; k < S.d; k += 32) e[k] *= c2;     // e becomes dr
    __syncwarp();
    float dot = 0.f;
    for (int m = lane; m < S.mem; m += 32) {
        const float am = a[m];
        float da = 0.f;
        for (int k = 0; k < S.d; ++k) {
            const float dr = e[k];
            da = fmaf(sM[m * S.ms + k], dr, da);
            atomicAdd(gM + m * S.ms + k, am * dr);
        }
        tm[m] = da;
        dot = fmaf(am, da, dot);
    }
    dot = lr_warp_sum(dot);
    for (int m = lane; m < S.mem; m += 32) tm[m] = a[m] * (tm[m] - dot);   // dkey
    __syncwarp();
#pragma unroll
    for (int v = 0; v < LR_MAX_D / 32; ++v) {
        const int k = lane + 32 * v;
        if (k < S.d) {
            const float p = vp[k], q = vq[k];
            const float x = __fmul_rn(p, q);
            float dx = 0.f;
            for (int m = 0; m < S.mem; ++m) {
                const float dk = tm[m];
                dx = fmaf(sK[k * S.ks + m], dk, dx);
                atomicAdd(gK + k * S.ks + m, x * dk);
            }
            const float de = e[k];
            gp[v] += fmaf(dx, q, de);
            gq[v] += fmaf(dx, p, -de);
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(LR_WARPS * 32) lrml_step_kernel(LrmlArgs A) {
    extern __shared__ float sm[];
    const LrmlShape& S = A.sh;
    const int nK = S.d * S.ks, nM = S.mem * S.ms;
    float* sK = sm;
    float* sM = sK + nK;
    float* gK = sM + nM;
    float* gM = gK + nK;
    float* wbase = gM + nM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* vp = wbase + warp * S.per_warp;
    float* vqi = vp + S.d;
    float* vqj = vqi + S.d;
    float* ei = vqj + S.d;
    float* ej = ei + S.d;
    float* ai = ej + S.d;
    float* aj = ai + S.mem;
    float* tm = aj + S.mem;
    for (int t = threadIdx.x; t < S.d * S.mem; t += blockDim.x) {
        sK[(t / S.mem) * S.ks + (t % S.mem)] = A.dense[t];
        sM[(t / S.d) * S.ms + (t % S.d)] = A.dense[S.d * S.mem + t];
    }
    for (int t = threadIdx.x; t < nK + nM; t += blockDim.x) gK[t] = 0.f;
    __syncthreads();
    double loss = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * LR_WARPS + warp; t < A.batch; t += (int64_t)gridDim.x * LR_WARPS) {
        const int64_t u = A.u[t], it = A.i[t], jt = A.j[t];
        float sq = 0.f;
        for (int k = lane; k < S.d; k += 32) {
            const float p = A.P[u * S.d + k], qi = A.Q[it * S.d + k], qj = A.Q[jt * S.d + k];
            vp[k] = p; vqi[k] = qi; vqj[k] = qj;
            sq = fmaf(p, p, fmaf(qi, qi, fmaf(qj, qj, sq)));
        }
        __syncwarp();
        const float di = lrml_forward(S, sK, sM, vp, vqi, ai, ei, lane);
        const float dj = lrml_forward(S, sK, sM, vp, vqj, aj, ej, lane);
        sq = lr_warp_sum(sq);
        const float hz = (di - dj) + A.margin;
        if (lane == 0) loss += (double)(fmaxf(hz, 0.f) + A.reg * 0.5f * sq);
        float gp[LR_MAX_D / 32], gqi[LR_MAX_D / 32], gqj[LR_MAX_D / 32];
#pragma unroll
        for (int v = 0; v < LR_MAX_D / 32; ++v) gp[v] = gqi[v] = gqj[v] = 0.f;
        if (hz >= 0.f) {   // tf.maximum(x, 0): MaximumGrad routes the gradient to x where x >= 0
            lrml_backward(S, sK, sM, gK, gM, vp, vqi, ai, ei, tm, 1.f, gp, gqi, lane);
            lrml_backward(S, sK, sM, gK, gM, vp, vqj, aj, ej, tm, -1.f, gp, gqj, lane);
        }
#pragma unroll
        for (int v = 0; v < LR_MAX_D / 32; ++v) {
            const int k = lane + 32 * v;
            if (k < S.d) {
                atomicAdd(A.gP + u * S.d + k, fmaf(A.reg, vp[k], gp[v]));
                atomicAdd(A.gQ + it * S.d + k, fmaf(A.reg, vqi[k], gqi[v]));
                atomicAdd(A.gQ + jt * S.d + k, fmaf(A.reg, vqj[k], gqj[v]));
            }
        }
        __syncwarp();
    }
    __syncthreads();
    float* part = A.dense_part + (int64_t)blockIdx.x * S.n_dense;
    for (int t = threadIdx.x; t < S.d * S.mem; t += blockDim.x) {
        part[t] = gK[(t / S.mem) * S.ks + (t % S.mem)];
        part[S.d * S.mem + t] = gM[(t / S.d) * S.ms + (t % S.d)];
    }
    __shared__ double sl[LR_WARPS];
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) sl[warp] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0;
        for (int k = 0; k < LR_WARPS; ++k) ts += sl[k];
        A.loss_part[blockIdx.x] = ts;
    }
}

struct LrmlScoreArgs {
    LrmlShape sh;
    const float* P;
    const float* Q;
    const float* dense;
    const int32_t* u;
    const int32_t* i;
    int64_t n;
    float* out;
};

__global__ void __launch_bounds__(LR_WARPS * 32) lrml_score_kernel(LrmlScoreArgs A) {
    extern __shared__ float sm[];
    const LrmlShape& S = A.sh;
    const int nK = S.d * S.ks, nM = S.mem * S.ms;
    float* sK = sm;
    float* sM = sK + nK;
    float* wbase = sM + nM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* vp = wbase + warp * (3 * S.d + S.mem);
    float* vq = vp + S.d;
    float* e = vq + S.d;
    float* a = e + S.d;
    for (int t = threadIdx.x; t < S.d * S.mem; t += blockDim.x) {
        sK[(t / S.mem) * S.ks + (t % S.mem)] = A.dense[t];
        sM[(t / S.d) * S.ms + (t % S.d)] = A.dense[S.d * S.mem + t];
    }
    __syncthreads();
    for (int64_t t = (int64_t)blockIdx.x * LR_WARPS + warp; t < A.n; t += (int64_t)gridDim.x * LR_WARPS) {
        const int64_t u = A.u[t], it = A.i[t];
        for (int k = lane; k < S.d; k += 32) { vp[k] = A.P[u * S.d + k]; vq[k] = A.Q[it * S.d + k]; }
        __syncwarp();
        const float dist = lrml_forward(S, sK, sM, vp, vq, a, e, lane);
        if (lane == 0) A.out[t] = dist;
        __syncwarp();
    }
}

// dense_vector_apply_kernel / crb_dense_table_apply live in train_neumf.cu / train_dense.cu
int crb_dense_vector_apply(crb_handle* h, float* w, float* s1, float* s2, const float* parts, int n_parts, int n, int opt_kind, const OptDev& od,
                           cudaStream_t s);
int crb_dense_tables_apply(crb_handle* h, int n, const crb_table* const* tables, float* const* grads, int opt_kind, const OptDev& od, cudaStream_t s);

extern "C" int crb_train_step_lrml(crb_handle* h, const crb_table* P, const crb_table* Q, float* gradP, float* gradQ, float* dense,
                                   float* dense_s1, float* dense_s2, int32_t mem_size, const crb_opt* opt, const int32_t* u, const int32_t* i,
                                   const int32_t* j, int64_t batch, float margin, float reg, double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && P && Q && gradP && gradQ && dense && u && i && j, "null argument");
    CRB_CHECK_ARG(batch > 0, "batch");
    CRB_CHECK_ARG(P->dim == Q->dim, "table dims");
    LrmlArgs a;
    if (lrml_shape(P->dim, mem_size, &a.sh)) {
        crb_set_error("unsupported LRML shape (embed_size=%d must be a multiple of 4 and <= %d, mem_size=%d <= %d)", P->dim, LR_MAX_D, mem_size, LR_MAX_MEM);
        return CRB_ERR_UNSUPPORTED;
    }
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    const int dk = opt_kind == OPT_ADAM_TF1 ? OPT_ADAM_LAZY : opt_kind;
    CRB_CHECK_ARG(dk == OPT_SGD || dense_s1, "dense slot s1 is NULL");
    CRB_CHECK_ARG(dk != OPT_ADAM_LAZY || dense_s2, "dense slot s2 is NULL");
    CRB_CUDA(cudaSetDevice(h->device));
    if ((rc = crb_ws_reserve(h, batch, 4, 4, s))) return rc;
    int grid = (int)((batch + 4 * LR_WARPS - 1) / (4 * LR_WARPS));
    if (grid > h->sm_count) grid = h->sm_count;
    if (grid < 1) grid = 1;
    const int64_t need = (int64_t)grid * a.sh.n_dense;
    if (need > h->cap_dense) {
        CRB_CUDA(cudaStreamSynchronize(s));
        cudaFree(h->dense_grad);
        h->dense_grad = nullptr; h->cap_dense = 0;
        CRB_CUDA(cudaMalloc(&h->dense_grad, sizeof(float) * need));
        h->cap_dense = need;
    }
    const int32_t *du = u, *di = i, *dj = j;
    if (!crb_is_device_ptr(u)) { CRB_CUDA(cudaMemcpyAsync(h->idx[0], u, 4 * batch, cudaMemcpyHostToDevice, s)); du = h->idx[0]; }
    if (!crb_is_device_ptr(i)) { CRB_CUDA(cudaMemcpyAsync(h->idx[1], i, 4 * batch, cudaMemcpyHostToDevice, s)); di = h->idx[1]; }
    if (!crb_is_device_ptr(j)) { CRB_CUDA(cudaMemcpyAsync(h->idx[2], j, 4 * batch, cudaMemcpyHostToDevice, s)); dj = h->idx[2]; }
    a.P = P->w; a.Q = Q->w; a.gP = gradP; a.gQ = gradQ; a.dense = dense; a.dense_part = h->dense_grad;
    a.u = du; a.i = di; a.j = dj; a.batch = batch; a.margin = margin; a.reg = reg; a.loss_part = h->block_loss;
    const size_t smem = sizeof(float) * (2 * (size_t)(a.sh.d * a.sh.ks + a.sh.mem * a.sh.ms) + (size_t)LR_WARPS * a.sh.per_warp);
    if (smem > 220 * 1024) { crb_set_error("LRML memory module too large for shared memory (%zu bytes)", smem); return CRB_ERR_UNSUPPORTED; }
    CRB_CUDA(cudaFuncSetAttribute(lrml_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if ((rc = crb_prof_begin(h, s))) return rc;
    lrml_step_kernel<<<grid, LR_WARPS * 32, smem, s>>>(a);
    if ((rc = crb_prof_end(h, s))) return rc;
    h->launches++;
    const crb_table* tabs2[2] = {P, Q};
    float* grads2[2] = {gradP, gradQ};
    if ((rc = crb_dense_tables_apply(h, 2, tabs2, grads2, dk, od, s))) return rc;
    if ((rc = crb_dense_vector_apply(h, dense, dense_s1, dense_s2, h->dense_grad, grid, a.sh.n_dense, dk, od, s))) return rc;
    h->step_grid = grid;
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    if ((rc = crb_launch_loss_final(h, ld, s))) return rc;
    CRB_CUDA(cudaGetLastError());
    if (loss_out && !crb_is_device_ptr(loss_out)) {
        CRB_CUDA(cudaMemcpyAsync(loss_out, h->loss_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
        CRB_CUDA(cudaStreamSynchronize(s));
    }
    return CRB_OK;
}

extern "C" int crb_score_pairs_lrml(crb_handle* h, const float* P, const float* Q, const float* dense, int32_t dim, int32_t mem_size,
                                    const int32_t* u, const int32_t* i, int64_t n, float* scores, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && P && Q && dense && u && i && scores, "null argument");
    CRB_CHECK_ARG(crb_is_device_ptr(u) && crb_is_device_ptr(i) && crb_is_device_ptr(scores), "u/i/scores must be device pointers");
    LrmlScoreArgs a;
    if (lrml_shape(dim, mem_size, &a.sh)) { crb_set_error("unsupported LRML shape (embed_size=%d, mem_size=%d)", dim, mem_size); return CRB_ERR_UNSUPPORTED; }
    if (n == 0) return CRB_OK;
    a.P = P; a.Q = Q; a.dense = dense; a.u = u; a.i = i; a.n = n; a.out = scores;
    const size_t smem = sizeof(float) * ((size_t)(a.sh.d * a.sh.ks + a.sh.mem * a.sh.ms) + (size_t)LR_WARPS * (3 * a.sh.d + a.sh.mem));
    if (smem > 220 * 1024) { crb_set_error("LRML memory module too large for shared memory (%zu bytes)", smem); return CRB_ERR_UNSUPPORTED; }
    CRB_CUDA(cudaSetDevice(h->device));
    CRB_CUDA(cudaFuncSetAttribute(lrml_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = (n + 8 * LR_WARPS - 1) / (8 * LR_WARPS);
    if (grid > (int64_t)h->sm_count * 2) grid = (int64_t)h->sm_count * 2;
    lrml_score_kernel<<<(int)grid, LR_WARPS * 32, smem, s>>>(a);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}
