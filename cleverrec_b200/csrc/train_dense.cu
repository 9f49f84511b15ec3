// Training steps of the models whose table gradients are DENSE in the reference, so that TF applies the dense form of
// the optimizer to every row every step:
//   CML   model/ranking/CML.py:39-70   hinge on the closest negative x WARP weight + covariance regulariser over [Q;P]
//   FISM  model/ranking/FISM.py:40-63  history mean (utils/tools.py:90-97) + L2 over the whole P, Q, b
// Per step:  [colsum]  ->  sample kernel (sparse part of the gradient, red.global.add.f32 into a dense gradient buffer)
//            ->  dense_table_apply_kernel (adds the closed-form dense term, TF dense optimizer apply, zeroes the buffer)
#include <cstddef>

#include "rowopt.cuh"

struct DenseTable {
    float* w;
    float* s1;
    float* s2;
    float* grad;   // [rows, dim] accumulated sparse part, zero between steps
    int64_t rows;
};

static int dgrid(crb_handle* h, int64_t n, int per_block) {
    int64_t b = (n + per_block - 1) / per_block;
    int64_t cap = (int64_t)h->sm_count * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

__device__ __forceinline__ void block_sum_to(double v, double* out) {
    __shared__ double sm[8];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sm[k];
        out[blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------------ column sums (CML)
// partial[b][c] = sum over the rows handled by block b; fixed assignment -> deterministic
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ w, int64_t rows, int dim, float* partial) {
    // thread c (< dim) sums column c over rows blockIdx.x, +gridDim.x, ...
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        float acc = 0.f;
        for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) acc += w[r * dim + c];
        partial[(int64_t)blockIdx.x * dim + c] = acc;
    }
}
// mean[c] = (sum_a + sum_b) / n_rows ; mean_sum = sum_c mean[c].  A block owns 32 columns; each of its 8 warps adds a slice of the
// partials (double, in order), the eight slice sums are added in warp order; the block that finishes last (ticket) adds the means in
// column order.  One block summing every partial in one chain per thread took 118 us of CML's 283 us step.
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* pa, int na, const float* pb, int nb, int dim, double n_rows, float* mean,
                                                           float* mean_sum, unsigned int* ticket) {
    __shared__ double sd[8][32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    double acc = 0.0;
    if (c < dim) {
        const int pera = (na + 7) / 8, a0 = w * pera, a1 = min(na, a0 + pera);
#pragma unroll 8
        for (int k = a0; k < a1; ++k) acc += pa[(int64_t)k * dim + c];
        const int perb = (nb + 7) / 8, b0 = w * perb, b1 = min(nb, b0 + perb);
#pragma unroll 8
        for (int k = b0; k < b1; ++k) acc += pb[(int64_t)k * dim + c];
    }
    sd[w][lane] = acc;
    __syncthreads();
    if (w == 0 && c < dim) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += sd[q][lane];
        mean[c] = (float)(t / n_rows);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        float t = 0.f;
        for (int k = 0; k < dim; ++k) t += __ldcg(mean + k);
        *mean_sum = t;
        *ticket = 0u;
    }
}

// ------------------------------------------------------------------------------------------------ dense apply
struct DenseApplyArgs {
    DenseTable T;
    int dim;
    int opt_kind;      // OPT_SGD / OPT_ADAGRAD / OPT_ADAM_LAZY (= TF dense ApplyAdam)
    OptDev opt;
    float l2;          // gradient += l2 * w                                  (FISM: reg / batch_size, reg_bias)
    float cov;         // gradient += cov * (S_r - (w - mean_c)), S_r = sum_c (w - mean_c)   (CML: 2*reg/n)
    const float* mean; // [dim] column means of [Q;P]
    const float* mean_sum;
    double* loss_part; // per-block partial of the dense loss term (l2: 0.5*l2'*w^2 handled by caller scale; cov: see below)
    float loss_l2;     // loss += loss_l2 * 0.5 * w^2
    float loss_cov;    // loss += loss_cov * (S_r^2 - sum_c xc^2)                (CML: reg / n)
};

// one warp per row (dim <= 512: up to 4 float4 per lane); `block` / `n_blocks`: this CTA's position among the CTAs working on the table
__device__ __forceinline__ void dense_table_apply_rows(const DenseApplyArgs& a, int block, int n_blocks, double* loss_slot) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)block * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)n_blocks * blockDim.x) >> 5;
    double loss = 0.0;
    const float msum = a.cov != 0.f ? *a.mean_sum : 0.f;
    for (int64_t r = warp; r < a.T.rows; r += n_warps) {
        float4 W[4], G[4];
        float rs = 0.f, sq = 0.f, xsq = 0.f;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int c = (lane + 32 * v) * 4;
            if (c < a.dim) {
                W[v] = ld4(a.T.w + r * a.dim + c);
                G[v] = ld4(a.T.grad + r * a.dim + c);
                rs += W[v].x + W[v].y + W[v].z + W[v].w;
                sq += dot4(W[v], W[v]);
                if (a.cov != 0.f) {
                    const float4 m = ld4(a.mean + c);
                    const float4 x = make_float4(W[v].x - m.x, W[v].y - m.y, W[v].z - m.z, W[v].w - m.w);
                    xsq += dot4(x, x);
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            rs += __shfl_xor_sync(0xffffffffu, rs, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
            xsq += __shfl_xor_sync(0xffffffffu, xsq, o);
        }
        const float S = rs - msum;  // sum_c (w - mean_c)
        if (lane == 0) loss += (double)(a.loss_l2 * 0.5f * sq) + (double)a.loss_cov * ((double)S * (double)S - (double)xsq);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int c = (lane + 32 * v) * 4;
            if (c >= a.dim) continue;
            float4 g = G[v];
            g.x = fmaf(a.l2, W[v].x, g.x); g.y = fmaf(a.l2, W[v].y, g.y); g.z = fmaf(a.l2, W[v].z, g.z); g.w = fmaf(a.l2, W[v].w, g.w);
            if (a.cov != 0.f) {
                const float4 m = ld4(a.mean + c);
                g.x = fmaf(a.cov, S - (W[v].x - m.x), g.x); g.y = fmaf(a.cov, S - (W[v].y - m.y), g.y);
                g.z = fmaf(a.cov, S - (W[v].z - m.z), g.z); g.w = fmaf(a.cov, S - (W[v].w - m.w), g.w);
            }
            float4 w = W[v];
            const int64_t off = r * a.dim + c;
            if (a.opt_kind == OPT_SGD) {
                w.x -= a.opt.lr * g.x; w.y -= a.opt.lr * g.y; w.z -= a.opt.lr * g.z; w.w -= a.opt.lr * g.w;
            } else if (a.opt_kind == OPT_ADAGRAD) {
                float4 acc = ld4(a.T.s1 + off);
                adagrad_elem(w.x, acc.x, g.x, a.opt.lr); adagrad_elem(w.y, acc.y, g.y, a.opt.lr);
                adagrad_elem(w.z, acc.z, g.z, a.opt.lr); adagrad_elem(w.w, acc.w, g.w, a.opt.lr);
                st4(a.T.s1 + off, acc);
            } else {  // ApplyAdam
                float4 m = ld4(a.T.s1 + off), vv = ld4(a.T.s2 + off);
                adam_touch_elem(w.x, m.x, vv.x, g.x, a.opt); adam_touch_elem(w.y, m.y, vv.y, g.y, a.opt);
                adam_touch_elem(w.z, m.z, vv.z, g.z, a.opt); adam_touch_elem(w.w, m.w, vv.w, g.w, a.opt);
                st4(a.T.s1 + off, m); st4(a.T.s2 + off, vv);
            }
            st4(a.T.w + off, w);
            st4(a.T.grad + off, make_float4(0.f, 0.f, 0.f, 0.f));
        }
    }
    // block-level sum of the per-thread loss into *loss_slot
    __shared__ double sm_l[8];
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if ((threadIdx.x & 31) == 0) sm_l[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sm_l[k];
        *loss_slot = t;
    }
}

__global__ void __launch_bounds__(256) dense_table_apply_kernel(DenseApplyArgs a) {
    dense_table_apply_rows(a, blockIdx.x, gridDim.x, a.loss_part + blockIdx.x);
}

// Up to four tables in ONE launch (NeuMF's four embedding tables, NAIS / SBPR's P, Q and bias, LRML's P and Q): at the batch sizes
// the reference ships these steps are a chain of microsecond kernels, so every launch removed is a few per cent of the step.  CTAs
// [first[t], first[t+1]) work on table t.
struct DenseApplyMulti {
    DenseApplyArgs t[4];
    int first[5];
};
__global__ void __launch_bounds__(256) dense_tables_apply_kernel(DenseApplyMulti m) {
    int t = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) t += ((int)blockIdx.x >= m.first[q]) ? 1 : 0;
    dense_table_apply_rows(m.t[t], blockIdx.x - m.first[t], m.first[t + 1] - m.first[t], m.t[t].loss_part + blockIdx.x);
}

__device__ __forceinline__ void atomic_add4(float* p, float4 v) {
    atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
}

// ------------------------------------------------------------------------------------------------ CML sample kernel
struct CmlArgs {
    const float* P;
    const float* Q;
    float* gP;
    float* gQ;
    const int32_t* u;
    const int32_t* i;
    const int32_t* neg;  // [batch, R]
    int64_t batch;
    int dim, R;
    float margin, item_nums;
    double* loss_part;
};

template <int LANES, int VPL>
__global__ void __launch_bounds__(256) cml_step_kernel(CmlArgs a) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31, gl = lane % LANES, sub = lane / LANES;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double loss = 0.0;
    for (int64_t base = warp * GPW; base < a.batch; base += n_warps * GPW) {
        const int64_t t = base + sub;
        const bool active = t < a.batch;
        const int64_t tt = active ? t : a.batch - 1;
        const int32_t u = a.u[tt], it = a.i[tt];
        float4 p[VPL], qi[VPL], qmin[VPL];
        float dui = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            p[v] = c < a.dim ? ld4(a.P + (int64_t)u * a.dim + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            qi[v] = c < a.dim ? ld4(a.Q + (int64_t)it * a.dim + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 d = make_float4(p[v].x - qi[v].x, p[v].y - qi[v].y, p[v].z - qi[v].z, p[v].w - qi[v].w);
            dui += dot4(d, d);
            qmin[v] = qi[v];
        }
        dui = group_sum<LANES>(dui);
        float dmin = INFINITY;
        int kmin = -1, imposters = 0;
        for (int r = 0; r < a.R; ++r) {
            const int32_t k = a.neg[tt * a.R + r];
            float4 q[VPL];
            float d2 = 0.f;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c = (gl + LANES * v) * 4;
                q[v] = c < a.dim ? ld4(a.Q + (int64_t)k * a.dim + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 d = make_float4(p[v].x - q[v].x, p[v].y - q[v].y, p[v].z - q[v].z, p[v].w - q[v].w);
                d2 += dot4(d, d);
            }
            d2 = group_sum<LANES>(d2);
            imposters += (dui + a.margin - d2 > 0.f) ? 1 : 0;   // CML.py:50
            if (d2 < dmin) {                                     // reduce_min, first minimum
                dmin = d2; kmin = k;
#pragma unroll
                for (int v = 0; v < VPL; ++v) qmin[v] = q[v];
            }
        }
        const float hinge = dui + a.margin - dmin;               // CML.py:46
        if (active && hinge > 0.f) {
            const float rank = ((float)imposters / (float)a.R) * a.item_nums / (float)a.R;   // CML.py:52
            const float wgt = logf(rank + 1.f);
            if (gl == 0) loss += (double)(hinge * wgt);
            const float c2 = 2.f * wgt;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c = (gl + LANES * v) * 4;
                if (c >= a.dim) continue;
                // d(hinge)/dp = 2(p-qi) - 2(p-qmin) = 2(qmin-qi); d/dqi = -2(p-qi); d/dqmin = +2(p-qmin)
                const float4 gp = make_float4(c2 * (qmin[v].x - qi[v].x), c2 * (qmin[v].y - qi[v].y), c2 * (qmin[v].z - qi[v].z), c2 * (qmin[v].w - qi[v].w));
                const float4 gi = make_float4(-c2 * (p[v].x - qi[v].x), -c2 * (p[v].y - qi[v].y), -c2 * (p[v].z - qi[v].z), -c2 * (p[v].w - qi[v].w));
                const float4 gk = make_float4(c2 * (p[v].x - qmin[v].x), c2 * (p[v].y - qmin[v].y), c2 * (p[v].z - qmin[v].z), c2 * (p[v].w - qmin[v].w));
                atomic_add4(a.gP + (int64_t)u * a.dim + c, gp);
                atomic_add4(a.gQ + (int64_t)it * a.dim + c, gi);
                atomic_add4(a.gQ + (int64_t)kmin * a.dim + c, gk);
            }
        }
    }
    block_sum_to(loss, a.loss_part);
}

// ------------------------------------------------------------------------------------------------ FISM
struct FismArgs {
    const float* P;      // [(I+1), d]   history-side item embeddings
    const float* Q;      // [(I+1), d]
    const float* b;      // [(I+1)]
    float* gP;
    float* gQ;
    float* gb;
    const int32_t* u;
    const int32_t* i;
    const int32_t* j;
    const int32_t* nbr;  // u_neighbors_num fed by the sampler (len(set(items)), utils/sampler.py:51,64)
    const int64_t* list_start;  // [U] offset of user u's interaction list in pos_item (reference order, duplicates kept)
    const int32_t* list_len;    // [U] len(items)
    const int32_t* pos_item;
    int64_t batch;
    int dim;
    float alpha;
    double* loss_part;
};

// s_u = n_u^-alpha * (1/len_u) * sum_{k in list(u)} P[k]   (FISM.py:42,51 + utils/tools.py:95: values 1/len(items))
template <int LANES, int VPL>
__device__ __forceinline__ float fism_user_vector(float4* s, const float* __restrict__ P, const int32_t* __restrict__ items, int n,
                                                  int nbr, float alpha, int dim, int gl) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) s[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n <= 0 || nbr <= 0) return 0.f;
    const float inv_len = 1.f / (float)n;
    for (int k = 0; k < n; ++k) {
        const int64_t row = items[k];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            if (c < dim) {
                const float4 x = ld4(P + row * dim + c);
                s[v].x = fmaf(x.x, inv_len, s[v].x); s[v].y = fmaf(x.y, inv_len, s[v].y);
                s[v].z = fmaf(x.z, inv_len, s[v].z); s[v].w = fmaf(x.w, inv_len, s[v].w);
            }
        }
    }
    const float coeff = powf((float)nbr, -alpha);
#pragma unroll
    for (int v = 0; v < VPL; ++v) { s[v].x *= coeff; s[v].y *= coeff; s[v].z *= coeff; s[v].w *= coeff; }
    return coeff * inv_len;
}

template <int LANES, int VPL>
__global__ void __launch_bounds__(256) fism_step_kernel(FismArgs a) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31, gl = lane % LANES, sub = lane / LANES;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double loss = 0.0;
    for (int64_t base = warp * GPW; base < a.batch; base += n_warps * GPW) {
        const int64_t t = base + sub;
        const bool active = t < a.batch;
        const int64_t tt = active ? t : a.batch - 1;
        const int32_t u = a.u[tt], it = a.i[tt], jt = a.j[tt];
        const int32_t* items = a.pos_item + a.list_start[u];
        const int n = a.list_len[u];
        float4 s[VPL];
        const float scale = fism_user_vector<LANES, VPL>(s, a.P, items, n, a.nbr[tt], a.alpha, a.dim, gl);
        float xi = 0.f, xj = 0.f;
        float4 qi[VPL], qj[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            qi[v] = c < a.dim ? ld4(a.Q + (int64_t)it * a.dim + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            qj[v] = c < a.dim ? ld4(a.Q + (int64_t)jt * a.dim + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            xi += dot4(qi[v], s[v]);
            xj += dot4(qj[v], s[v]);
        }
        xi = group_sum<LANES>(xi) + a.b[it];
        xj = group_sum<LANES>(xj) + a.b[jt];
        const float x = xi - xj;
        const float g = -sigmoid_f(-x);
        if (active) {
            if (gl == 0) {
                loss += (double)softplus_neg(x);
                atomicAdd(a.gb + it, g);
                atomicAdd(a.gb + jt, -g);
            }
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c = (gl + LANES * v) * 4;
                if (c >= a.dim) continue;
                atomic_add4(a.gQ + (int64_t)it * a.dim + c, make_float4(g * s[v].x, g * s[v].y, g * s[v].z, g * s[v].w));
                atomic_add4(a.gQ + (int64_t)jt * a.dim + c, make_float4(-g * s[v].x, -g * s[v].y, -g * s[v].z, -g * s[v].w));
            }
            // d/dP[k] = g * coeff/len * (q_i - q_j) for every k of the history list (with multiplicity)
            const float gs = g * scale;
            for (int k = 0; k < n; ++k) {
                const int64_t row = items[k];
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    const int c = (gl + LANES * v) * 4;
                    if (c >= a.dim) continue;
                    atomic_add4(a.gP + row * a.dim + c, make_float4(gs * (qi[v].x - qj[v].x), gs * (qi[v].y - qj[v].y), gs * (qi[v].z - qj[v].z),
                                                                    gs * (qi[v].w - qj[v].w)));
                }
            }
        }
    }
    block_sum_to(loss, a.loss_part);
}

// evaluation: user vectors of `users` (row k of out) -- FISM.py:70 `coeff * u_neighbors_embed`
template <int LANES, int VPL>
__global__ void __launch_bounds__(256) fism_user_vectors_kernel(const float* P, int dim, const int32_t* users, const int32_t* nbr, int64_t n,
                                                                const int64_t* list_start, const int32_t* list_len, const int32_t* pos_item,
                                                                float alpha, float* out) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31, gl = lane % LANES, sub = lane / LANES;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * GPW; base < n; base += n_warps * GPW) {
        const int64_t t = base + sub;
        if (t >= n) continue;
        const int32_t u = users[t];
        float4 s[VPL];
        fism_user_vector<LANES, VPL>(s, P, pos_item + list_start[u], list_len[u], nbr[t], alpha, dim, gl);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            if (c < dim) st4(out + t * dim + c, s[v]);
        }
    }
}

// K11: tf.clip_by_norm(rows, max_norm, axes=[1])  (CML.py:72-78; used only for the user rows of CML's full-rank _predict)
__global__ void __launch_bounds__(256) clip_rows_kernel(const float* src, float* dst, int64_t rows, int dim, float max_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += n_warps) {
        float sq = 0.f;
        for (int c = lane; c < dim; c += 32) { const float x = src[r * dim + c]; sq = fmaf(x, x, sq); }
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float nrm = sqrtf(sq);
        const float sc = nrm > max_norm ? max_norm / nrm : 1.f;   // t * clip_norm / max(l2norm, clip_norm)
        for (int c = lane; c < dim; c += 32) dst[r * dim + c] = src[r * dim + c] * sc;
    }
}

__global__ void sum_parts_kernel(const double* a, int na, const double* b, int nb, const double* c, int nc, double* out) {
    double v = 0.0;   // same order of adds as ever; the unrolls only put 8 of a lane's independent loads in flight at a time
#pragma unroll 8
    for (int k = threadIdx.x; k < na; k += 32) v += a[k];
#pragma unroll 8
    for (int k = threadIdx.x; k < nb; k += 32) v += b[k];
#pragma unroll 8
    for (int k = threadIdx.x; k < nc; k += 32) v += c[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) *out = v;
}

// ------------------------------------------------------------------------------------------------ host side
static int dense_opt_kind(int opt_kind) { return opt_kind == OPT_ADAM_TF1 ? OPT_ADAM_LAZY : opt_kind; }

static int dense_ws(crb_handle* h, int64_t floats, cudaStream_t s) {
    if (floats <= h->cap_dense) return CRB_OK;
    CRB_CUDA(cudaStreamSynchronize(s));
    cudaFree(h->dense_grad);
    h->dense_grad = nullptr;
    h->cap_dense = 0;
    CRB_CUDA(cudaMalloc(&h->dense_grad, sizeof(float) * floats));
    h->cap_dense = floats;
    return CRB_OK;
}

static int stage_dev_i32(crb_handle* h, const int32_t* src, int64_t n, int32_t* scratch, const int32_t** out, cudaStream_t s) {
    if (crb_is_device_ptr(src)) { *out = src; return CRB_OK; }
    CRB_CUDA(cudaMemcpyAsync(scratch, src, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    *out = scratch;
    return CRB_OK;
}

static int check_dense_table(const crb_table* T, const float* grad, int opt_kind, const char* name) {
    int rc = crb_table_check(T, dense_opt_kind(opt_kind), name);
    if (rc) return rc;
    if (!grad || !crb_is_device_ptr(grad)) { crb_set_error("table %s: dense gradient buffer must be a zeroed device array", name); return CRB_ERR_ARG; }
    return CRB_OK;
}

static int finish_loss_host(crb_handle* h, double* loss_out, cudaStream_t s) {
    if (!loss_out || crb_is_device_ptr(loss_out)) return CRB_OK;
    CRB_CUDA(cudaMemcpyAsync(loss_out, h->loss_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
    CRB_CUDA(cudaStreamSynchronize(s));
    return CRB_OK;
}

template <int LANES, int VPL>
static int launch_cml_t(crb_handle* h, const CmlArgs& a, int grid, cudaStream_t s) {
    cml_step_kernel<LANES, VPL><<<grid, 256, 0, s>>>(a);
    return CRB_OK;
}
template <int LANES, int VPL>
static int launch_fism_t(crb_handle* h, const FismArgs& a, int grid, cudaStream_t s) {
    fism_step_kernel<LANES, VPL><<<grid, 256, 0, s>>>(a);
    return CRB_OK;
}
template <int LANES, int VPL>
static int launch_fismvec_t(crb_handle* h, const float* P, int dim, const int32_t* users, const int32_t* nbr, int64_t n, float alpha, float* out,
                            int grid, cudaStream_t s) {
    fism_user_vectors_kernel<LANES, VPL><<<grid, 256, 0, s>>>(P, dim, users, nbr, n, h->list_start, h->list_len, h->pos_item, alpha, out);
    return CRB_OK;
}

#define CRB_DIM_DISPATCH(dim, FN, ...)                                   \
    ((dim) <= 32 ? FN<8, 1>(__VA_ARGS__)                                 \
     : (dim) <= 64 ? FN<16, 1>(__VA_ARGS__)                              \
     : (dim) <= 128 ? FN<32, 1>(__VA_ARGS__)                             \
     : (dim) <= 256 ? FN<32, 2>(__VA_ARGS__)                             \
                    : FN<32, 4>(__VA_ARGS__))

extern "C" int crb_train_step_cml(crb_handle* h, const crb_table* P, const crb_table* Q, float* gradP, float* gradQ, const crb_opt* opt,
                                  const int32_t* u, const int32_t* i, const int32_t* neg, int64_t batch, int32_t neg_ratio, float margin,
                                  float reg, int64_t item_nums, double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && u && i && neg, "null argument");
    CRB_CHECK_ARG(batch > 0 && neg_ratio >= 1, "sizes");
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    if ((rc = check_dense_table(P, gradP, opt_kind, "P"))) return rc;
    if ((rc = check_dense_table(Q, gradQ, opt_kind, "Q"))) return rc;
    CRB_CHECK_ARG(P->dim == Q->dim, "P.dim != Q.dim");
    CRB_CUDA(cudaSetDevice(h->device));
    const int dim = P->dim;
    const int parts = h->sm_count * 4;                       // column-sum partial blocks per table
    const int grid = dgrid(h, batch, 256 / 32);
    // dense workspace: [neg staging as int32][colsum partials P, Q][mean][mean_sum]
    const int64_t neg_floats = (((int64_t)batch * neg_ratio + 3) / 4) * 4;   // keeps the float4 views behind it 16-byte aligned
    const int64_t need = neg_floats + 2 * (int64_t)parts * dim + dim + 8;
    if ((rc = dense_ws(h, need, s))) return rc;
    if ((rc = crb_ws_reserve(h, batch, dim, 4, s))) return rc;
    int32_t* neg_scratch = reinterpret_cast<int32_t*>(h->dense_grad);
    float* part_p = h->dense_grad + neg_floats;
    float* part_q = part_p + (int64_t)parts * dim;
    float* mean = part_q + (int64_t)parts * dim;
    float* mean_sum = mean + dim;
    const int32_t *du, *di, *dn;
    if ((rc = stage_dev_i32(h, u, batch, h->idx[0], &du, s))) return rc;
    if ((rc = stage_dev_i32(h, i, batch, h->idx[1], &di, s))) return rc;
    if ((rc = stage_dev_i32(h, neg, batch * neg_ratio, neg_scratch, &dn, s))) return rc;
    const double n_rows = (double)(P->rows + Q->rows);
    colsum_partial_kernel<<<parts, 256, 0, s>>>(P->w, P->rows, dim, part_p);
    colsum_partial_kernel<<<parts, 256, 0, s>>>(Q->w, Q->rows, dim, part_q);
    colsum_final_kernel<<<(dim + 31) / 32, 256, 0, s>>>(part_p, parts, part_q, parts, dim, n_rows, mean, mean_sum, &h->ctr->pad[0]);
    // loss partials: block_loss[0..grid) hinge, then two dense passes
    double* lp_hinge = h->block_loss;
    CmlArgs a = {P->w, Q->w, gradP, gradQ, du, di, dn, batch, dim, neg_ratio, margin, (float)item_nums, lp_hinge};
    if ((rc = crb_prof_begin(h, s))) return rc;
    CRB_DIM_DISPATCH(dim, launch_cml_t, h, a, grid, s);
    if ((rc = crb_prof_end(h, s))) return rc;
    const int gp = dgrid(h, P->rows, 8), gq = dgrid(h, Q->rows, 8);
    double* dp = h->dense_loss;  // per-block partials of the dense loss term
    DenseApplyArgs da;
    da.dim = dim; da.opt_kind = dense_opt_kind(opt_kind); da.opt = od; da.l2 = 0.f; da.cov = (float)(2.0 * (double)reg / n_rows);
    da.mean = mean; da.mean_sum = mean_sum; da.loss_l2 = 0.f; da.loss_cov = (float)((double)reg / n_rows);
    da.T = {P->w, P->s1, P->s2, gradP, P->rows}; da.loss_part = dp;
    dense_table_apply_kernel<<<gp, 256, 0, s>>>(da);
    da.T = {Q->w, Q->s1, Q->s2, gradQ, Q->rows}; da.loss_part = dp + gp;
    dense_table_apply_kernel<<<gq, 256, 0, s>>>(da);
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    sum_parts_kernel<<<1, 32, 0, s>>>(lp_hinge, grid, dp, gp, dp + gp, gq, ld);
    h->launches += 7;
    CRB_CUDA(cudaGetLastError());
    return finish_loss_host(h, loss_out, s);
}

extern "C" int crb_set_history_lists(crb_handle* h, const int64_t* list_start, const int32_t* list_len) {
    CRB_CHECK_ARG(h, "null handle");
    CRB_CHECK_ARG(crb_is_device_ptr(list_start) && crb_is_device_ptr(list_len), "list_start/list_len must be device pointers");
    h->list_start = list_start;
    h->list_len = list_len;
    return CRB_OK;
}

extern "C" int crb_train_step_fism(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_table* B, float* gradP, float* gradQ,
                                   float* gradB, const crb_opt* opt, const int32_t* u, const int32_t* i, const int32_t* j, const int32_t* nbr,
                                   int64_t batch, float alpha, float reg, float reg_bias, int64_t conf_batch_size, double* loss_out,
                                   void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && u && i && j && nbr, "null argument");
    CRB_CHECK_ARG(batch > 0 && conf_batch_size > 0, "sizes");
    if (!h->list_start || !h->pos_item) { crb_set_error("crb_train_step_fism before crb_set_history / crb_set_history_lists"); return CRB_ERR_STATE; }
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    if ((rc = check_dense_table(P, gradP, opt_kind, "P"))) return rc;
    if ((rc = check_dense_table(Q, gradQ, opt_kind, "Q"))) return rc;
    CRB_CHECK_ARG(B && B->w && B->dim == 1 && gradB, "bias table must have dim 1");
    CRB_CHECK_ARG(P->dim == Q->dim, "P.dim != Q.dim");
    CRB_CUDA(cudaSetDevice(h->device));
    const int dim = P->dim;
    if ((rc = crb_ws_reserve(h, batch, dim, 4, s))) return rc;
    const int32_t *du, *di, *dj, *dn;
    if ((rc = stage_dev_i32(h, u, batch, h->idx[0], &du, s))) return rc;
    if ((rc = stage_dev_i32(h, i, batch, h->idx[1], &di, s))) return rc;
    if ((rc = stage_dev_i32(h, j, batch, h->idx[2], &dj, s))) return rc;
    if ((rc = stage_dev_i32(h, nbr, batch, h->idx[3], &dn, s))) return rc;
    const int grid = dgrid(h, batch, 256 / 32);
    FismArgs a = {P->w, Q->w, B->w, gradP, gradQ, gradB, du, di, dj, dn, h->list_start, h->list_len, h->pos_item, batch, dim, alpha, h->block_loss};
    if ((rc = crb_prof_begin(h, s))) return rc;
    CRB_DIM_DISPATCH(dim, launch_fism_t, h, a, grid, s);
    if ((rc = crb_prof_end(h, s))) return rc;
    // dense part: reg*(l2(P)+l2(Q))/batch_size + reg_bias*l2(b)   (FISM.py:57)
    double* dp = h->dense_loss;
    const int gp = dgrid(h, P->rows, 8), gq = dgrid(h, Q->rows, 8), gb = dgrid(h, (B->rows + 3) / 4, 8);
    DenseApplyArgs da;
    da.dim = dim; da.opt_kind = dense_opt_kind(opt_kind); da.opt = od; da.cov = 0.f; da.mean = nullptr; da.mean_sum = nullptr; da.loss_cov = 0.f;
    da.l2 = reg / (float)conf_batch_size; da.loss_l2 = da.l2;
    da.T = {P->w, P->s1, P->s2, gradP, P->rows}; da.loss_part = dp;
    dense_table_apply_kernel<<<gp, 256, 0, s>>>(da);
    da.T = {Q->w, Q->s1, Q->s2, gradQ, Q->rows}; da.loss_part = dp + gp;
    dense_table_apply_kernel<<<gq, 256, 0, s>>>(da);
    // the bias vector is applied as a [(I+1)/4, 4] table (padded by the caller to a multiple of 4)
    CRB_CHECK_ARG(B->rows % 4 == 0, "bias length must be padded to a multiple of 4");
    da.dim = 4; da.l2 = reg_bias; da.loss_l2 = reg_bias;
    da.T = {B->w, B->s1, B->s2, gradB, B->rows / 4}; da.loss_part = dp + gp + gq;
    dense_table_apply_kernel<<<gb, 256, 0, s>>>(da);
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    sum_parts_kernel<<<1, 32, 0, s>>>(h->block_loss, grid, dp, gp + gq + gb, nullptr, 0, ld);
    h->launches += 5;
    CRB_CUDA(cudaGetLastError());
    return finish_loss_host(h, loss_out, s);
}

extern "C" int crb_fism_user_vectors(crb_handle* h, const float* P, int32_t dim, const int32_t* users, const int32_t* nbr, int64_t n, float alpha,
                                     float* out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && P && users && nbr && out, "null argument");
    CRB_CHECK_ARG(dim % 4 == 0 && dim <= 512, "dim % 4 == 0, dim <= 512");
    CRB_CHECK_ARG(crb_is_device_ptr(users) && crb_is_device_ptr(nbr) && crb_is_device_ptr(out), "users/nbr/out must be device pointers");
    if (!h->list_start || !h->pos_item) { crb_set_error("crb_fism_user_vectors before crb_set_history_lists"); return CRB_ERR_STATE; }
    if (n == 0) return CRB_OK;
    CRB_DIM_DISPATCH(dim, launch_fismvec_t, h, P, dim, users, nbr, n, alpha, out, dgrid(h, n, 8), s);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

extern "C" int crb_clip_rows(crb_handle* h, const float* src, float* dst, int64_t rows, int32_t dim, float max_norm, void* stream) {
    CRB_CHECK_ARG(h && src && dst && rows >= 0 && dim > 0, "bad argument");
    if (rows == 0) return CRB_OK;
    clip_rows_kernel<<<dgrid(h, rows, 8), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, dim, max_norm);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

int crb_dense_table_apply(crb_handle* h, const crb_table* T, float* grad, int opt_kind, const OptDev& od, float l2, double* loss_part,
                          int* grid_out, cudaStream_t s);
int crb_dense_tables_apply(crb_handle* h, int n, const crb_table* const* tables, float* const* grads, int opt_kind, const OptDev& od, cudaStream_t s);

// ------------------------------------------------------------------------------------------------ TransCF
// model/ranking/TransCF.py:38-71.  alpha_u = mean of Q over the user's training items (ui_sp_mat, utils/tools.py:100-113: values
// 1/len(items), duplicates counted), beta_i = mean of P over the item's users (iu_sp_mat: values 1/iu_nums[i]);
//   e_i = p_u + alpha_u * beta_i - q_i,  d_ui = |e_i|^2  (same for j);  loss = sum max(d_ui - d_uj + margin, 0)   (tools.py:73)
//        + reg1 * (|p_u - alpha_u|^2 + |q_i - beta_i|^2) + reg2 * (d_ui + margin - d_uj)^2                            (:65-71)
// The reference recomputes both SpMMs over ALL interactions every step; only the rows the batch touches are needed, so here a lane
// group builds alpha_u, beta_i, beta_j for its triplet from the two CSR-style lists and scatters the neighbourhood gradients back
// through them.  The table gradients are dense in TF (they flow through the SpMMs), hence the dense optimizer apply.
struct TcfArgs {
    const float* P;
    const float* Q;
    float* gP;
    float* gQ;
    const int32_t* u;
    const int32_t* i;
    const int32_t* j;
    const int64_t* ul_start;   // user -> items list inside upos
    const int32_t* ul_len;
    const int32_t* upos;
    const int64_t* il_start;   // item -> users list inside ipos
    const int32_t* il_len;
    const int32_t* ipos;
    int64_t batch;
    int dim;
    float margin, reg1, reg2;
    double* loss_part;
};

// mean of `table` rows members[0..n) in list order: acc = fma(x, 1/n, acc)  (the canonical order; restated in the tests' oracle)
template <int LANES, int VPL>
__device__ __forceinline__ void seg_mean(float4* s, const float* __restrict__ table, const int32_t* __restrict__ members, int n, int dim, int gl) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) s[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n <= 0) return;
    const float inv = 1.f / (float)n;
    for (int k = 0; k < n; ++k) {
        const int64_t row = members[k];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            if (c < dim) {
                const float4 x = ld4(table + row * dim + c);
                s[v].x = fmaf(x.x, inv, s[v].x); s[v].y = fmaf(x.y, inv, s[v].y); s[v].z = fmaf(x.z, inv, s[v].z); s[v].w = fmaf(x.w, inv, s[v].w);
            }
        }
    }
}

template <int LANES, int VPL>
__device__ __forceinline__ void seg_scatter(float* __restrict__ grad, const int32_t* __restrict__ members, int n, const float4* g, int dim, int gl) {
    if (n <= 0) return;
    const float inv = 1.f / (float)n;
    for (int k = 0; k < n; ++k) {
        const int64_t row = members[k];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            if (c < dim) atomic_add4(grad + row * dim + c, make_float4(g[v].x * inv, g[v].y * inv, g[v].z * inv, g[v].w * inv));
        }
    }
}

__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float k) { return make_float4(a.x * k, a.y * k, a.z * k, a.w * k); }

template <int LANES, int VPL>
__global__ void __launch_bounds__(256) transcf_step_kernel(TcfArgs a) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31, gl = lane % LANES, sub = lane / LANES;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double loss = 0.0;
    for (int64_t base = warp * GPW; base < a.batch; base += n_warps * GPW) {
        const int64_t t = base + sub;
        const bool active = t < a.batch;
        const int64_t tt = active ? t : a.batch - 1;
        const int32_t u = a.u[tt], it = a.i[tt], jt = a.j[tt];
        const int32_t* u_items = a.upos + a.ul_start[u];
        const int32_t* i_users = a.ipos + a.il_start[it];
        const int32_t* j_users = a.ipos + a.il_start[jt];
        const int nu = a.ul_len[u], ni = a.il_len[it], nj = a.il_len[jt];
        float4 al[VPL], bi[VPL], bj[VPL], p[VPL], qi[VPL], qj[VPL], ei[VPL], ej[VPL];
        seg_mean<LANES, VPL>(al, a.Q, u_items, nu, a.dim, gl);
        seg_mean<LANES, VPL>(bi, a.P, i_users, ni, a.dim, gl);
        seg_mean<LANES, VPL>(bj, a.P, j_users, nj, a.dim, gl);
        float dui = 0.f, duj = 0.f, rn = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            p[v] = c < a.dim ? ld4(a.P + (int64_t)u * a.dim + c) : z;
            qi[v] = c < a.dim ? ld4(a.Q + (int64_t)it * a.dim + c) : z;
            qj[v] = c < a.dim ? ld4(a.Q + (int64_t)jt * a.dim + c) : z;
            ei[v] = f4_sub(f4_add(p[v], f4_mul(al[v], bi[v])), qi[v]);
            ej[v] = f4_sub(f4_add(p[v], f4_mul(al[v], bj[v])), qj[v]);
            dui += dot4(ei[v], ei[v]);
            duj += dot4(ej[v], ej[v]);
            const float4 pa = f4_sub(p[v], al[v]), qb = f4_sub(qi[v], bi[v]);
            rn += dot4(pa, pa) + dot4(qb, qb);
        }
        dui = group_sum<LANES>(dui);
        duj = group_sum<LANES>(duj);
        rn = group_sum<LANES>(rn);
        const float x = dui - duj + a.margin;   // hinge argument and the residual of the distance regulariser
        const float cgrad = (x > 0.f ? 1.f : 0.f) + 2.f * a.reg2 * x;   // dL/dd_ui = -dL/dd_uj
        if (active) {
            if (gl == 0) loss += (double)fmaxf(x, 0.f) + (double)a.reg1 * (double)rn + (double)a.reg2 * (double)x * (double)x;
            float4 g_al[VPL], g_bi[VPL], g_bj[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c = (gl + LANES * v) * 4;
                if (c >= a.dim) continue;
                const float4 ci = f4_scale(ei[v], 2.f * cgrad), cj = f4_scale(ej[v], 2.f * cgrad);   // c * d d_ui/d e_i, c * d d_uj/d e_j
                const float4 pa = f4_scale(f4_sub(p[v], al[v]), 2.f * a.reg1), qb = f4_scale(f4_sub(qi[v], bi[v]), 2.f * a.reg1);
                atomic_add4(a.gP + (int64_t)u * a.dim + c, f4_add(f4_sub(ci, cj), pa));
                atomic_add4(a.gQ + (int64_t)it * a.dim + c, f4_sub(qb, ci));
                atomic_add4(a.gQ + (int64_t)jt * a.dim + c, cj);
                g_al[v] = f4_sub(f4_sub(f4_mul(ci, bi[v]), f4_mul(cj, bj[v])), pa);
                g_bi[v] = f4_sub(f4_mul(ci, al[v]), qb);
                g_bj[v] = f4_scale(f4_mul(cj, al[v]), -1.f);
            }
            seg_scatter<LANES, VPL>(a.gQ, u_items, nu, g_al, a.dim, gl);
            seg_scatter<LANES, VPL>(a.gP, i_users, ni, g_bi, a.dim, gl);
            seg_scatter<LANES, VPL>(a.gP, j_users, nj, g_bj, a.dim, gl);
        }
    }
    block_sum_to(loss, a.loss_part);
}

// out[r] = mean of `table` over the members of row rows[r] (or row r when rows == NULL): alpha for users, beta for items (evaluation)
template <int LANES, int VPL>
__global__ void __launch_bounds__(256) seg_mean_rows_kernel(const float* table, int dim, const int32_t* rows, int64_t n, const int64_t* start,
                                                            const int32_t* len, const int32_t* members, float* out) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31, gl = lane % LANES, sub = lane / LANES;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * GPW; base < n; base += n_warps * GPW) {
        const int64_t r = base + sub;
        if (r >= n) continue;
        const int64_t row = rows ? rows[r] : r;
        float4 s[VPL];
        seg_mean<LANES, VPL>(s, table, members + start[row], len[row], dim, gl);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            if (c < dim) st4(out + r * dim + c, s[v]);
        }
    }
}

// TransCF._predict (TransCF.py:79-85): dist(u, i) = sum_k (p_uk + A_uk * B_ik - q_ik)^2 as ONE sequential fp32 chain over k
// (e = fma(A, B, p) - q; acc = fma(e, e, acc)); A = alpha of all users, B = beta of all items (seg_mean_rows_kernel).
__global__ void __launch_bounds__(256) transcf_pairs_kernel(const float* __restrict__ P, const float* __restrict__ Q, const float* __restrict__ A,
                                                            const float* __restrict__ B, int dim, const int32_t* __restrict__ u,
                                                            const int32_t* __restrict__ it, int64_t n, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const float* p = P + (int64_t)u[k] * dim;
        const float* al = A + (int64_t)u[k] * dim;
        const float* q = Q + (int64_t)it[k] * dim;
        const float* be = B + (int64_t)it[k] * dim;
        float acc = 0.f;
        for (int c = 0; c < dim; ++c) {
            const float e = __fsub_rn(fmaf(al[c], be[c], p[c]), q[c]);
            acc = fmaf(e, e, acc);
        }
        out[k] = acc;
    }
}

template <int LANES, int VPL>
static int launch_tcf_t(crb_handle* h, const TcfArgs& a, int grid, cudaStream_t s) {
    transcf_step_kernel<LANES, VPL><<<grid, 256, 0, s>>>(a);
    return CRB_OK;
}
template <int LANES, int VPL>
static int launch_segmean_t(crb_handle* h, const float* table, int dim, const int32_t* rows, int64_t n, const int64_t* start, const int32_t* len,
                            const int32_t* members, float* out, int grid, cudaStream_t s) {
    seg_mean_rows_kernel<LANES, VPL><<<grid, 256, 0, s>>>(table, dim, rows, n, start, len, members, out);
    return CRB_OK;
}

extern "C" int crb_set_item_lists(crb_handle* h, const int64_t* item_start, const int32_t* item_len, const int32_t* item_users) {
    CRB_CHECK_ARG(h, "null handle");
    CRB_CHECK_ARG(crb_is_device_ptr(item_start) && crb_is_device_ptr(item_len) && crb_is_device_ptr(item_users), "item lists must be device pointers");
    h->ilist_start = item_start;
    h->ilist_len = item_len;
    h->ipos_user = item_users;
    return CRB_OK;
}

extern "C" int crb_train_step_transcf(crb_handle* h, const crb_table* P, const crb_table* Q, float* gradP, float* gradQ, const crb_opt* opt,
                                      const int32_t* u, const int32_t* i, const int32_t* j, int64_t batch, float margin, float reg1, float reg2,
                                      double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && u && i && j, "null argument");
    CRB_CHECK_ARG(batch > 0, "batch");
    if (!h->list_start || !h->pos_item || !h->ilist_start) {
        crb_set_error("crb_train_step_transcf before crb_set_history / crb_set_history_lists / crb_set_item_lists");
        return CRB_ERR_STATE;
    }
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    if ((rc = check_dense_table(P, gradP, opt_kind, "P"))) return rc;
    if ((rc = check_dense_table(Q, gradQ, opt_kind, "Q"))) return rc;
    CRB_CHECK_ARG(P->dim == Q->dim, "P.dim != Q.dim");
    CRB_CUDA(cudaSetDevice(h->device));
    const int dim = P->dim;
    if ((rc = crb_ws_reserve(h, batch, dim, 4, s))) return rc;
    const int32_t *du, *di, *dj;
    if ((rc = stage_dev_i32(h, u, batch, h->idx[0], &du, s))) return rc;
    if ((rc = stage_dev_i32(h, i, batch, h->idx[1], &di, s))) return rc;
    if ((rc = stage_dev_i32(h, j, batch, h->idx[2], &dj, s))) return rc;
    const int grid = dgrid(h, batch, 256 / 32);
    TcfArgs a = {P->w, Q->w, gradP, gradQ, du, di, dj, h->list_start, h->list_len, h->pos_item, h->ilist_start, h->ilist_len, h->ipos_user,
                 batch, dim, margin, reg1, reg2, h->block_loss};
    if ((rc = crb_prof_begin(h, s))) return rc;
    CRB_DIM_DISPATCH(dim, launch_tcf_t, h, a, grid, s);
    if ((rc = crb_prof_end(h, s))) return rc;
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    double* dp = h->dense_loss;
    int gp = 0, gq = 0;
    if ((rc = crb_dense_table_apply(h, P, gradP, opt_kind, od, 0.f, dp, &gp, s))) return rc;
    if ((rc = crb_dense_table_apply(h, Q, gradQ, opt_kind, od, 0.f, dp + gp, &gq, s))) return rc;
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    sum_parts_kernel<<<1, 32, 0, s>>>(h->block_loss, grid, dp, gp + gq, nullptr, 0, ld);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return finish_loss_host(h, loss_out, s);
}

// which = 0: alpha of `rows` users (mean of Q over their items);  which = 1: beta of `rows` items (mean of P over their users)
extern "C" int crb_transcf_neighbourhood(crb_handle* h, int32_t which, const float* table, int32_t dim, const int32_t* rows, int64_t n, float* out,
                                         void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && table && out && (which == 0 || which == 1), "bad argument");
    CRB_CHECK_ARG(dim % 4 == 0 && dim <= 512, "dim % 4 == 0, dim <= 512");
    CRB_CHECK_ARG((!rows || crb_is_device_ptr(rows)) && crb_is_device_ptr(out), "rows/out must be device pointers");
    if (which == 0 && (!h->list_start || !h->pos_item)) { crb_set_error("crb_transcf_neighbourhood before crb_set_history_lists"); return CRB_ERR_STATE; }
    if (which == 1 && !h->ilist_start) { crb_set_error("crb_transcf_neighbourhood before crb_set_item_lists"); return CRB_ERR_STATE; }
    if (n == 0) return CRB_OK;
    const int64_t* st = which == 0 ? h->list_start : h->ilist_start;
    const int32_t* ln = which == 0 ? h->list_len : h->ilist_len;
    const int32_t* mem = which == 0 ? h->pos_item : h->ipos_user;
    CRB_DIM_DISPATCH(dim, launch_segmean_t, h, table, dim, rows, n, st, ln, mem, out, dgrid(h, n, 8), s);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

extern "C" int crb_score_pairs_transcf(crb_handle* h, const float* P, const float* Q, const float* A, const float* B, int32_t dim, const int32_t* u,
                                       const int32_t* i, int64_t n, float* scores, void* stream) {
    CRB_CHECK_ARG(h && P && Q && A && B && u && i && scores, "null argument");
    CRB_CHECK_ARG(crb_is_device_ptr(u) && crb_is_device_ptr(i) && crb_is_device_ptr(scores), "u/i/scores must be device pointers");
    if (n == 0) return CRB_OK;
    transcf_pairs_kernel<<<dgrid(h, n, 256), 256, 0, (cudaStream_t)stream>>>(P, Q, A, B, dim, u, i, n, scores);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

// shared with train_neumf.cu / train_nais.cu: TF dense optimizer apply of one table from its dense gradient buffer
int crb_dense_table_apply(crb_handle* h, const crb_table* T, float* grad, int opt_kind, const OptDev& od, float l2, double* loss_part,
                          int* grid_out, cudaStream_t s) {
    int rc = check_dense_table(T, grad, opt_kind, "dense table");
    if (rc) return rc;
    DenseApplyArgs da;
    da.dim = T->dim; da.opt_kind = dense_opt_kind(opt_kind); da.opt = od; da.cov = 0.f; da.mean = nullptr; da.mean_sum = nullptr;
    da.loss_cov = 0.f; da.l2 = l2; da.loss_l2 = l2;
    da.T = {T->w, T->s1, T->s2, grad, T->rows}; da.loss_part = loss_part;
    const int g = dgrid(h, T->rows, 8);
    dense_table_apply_kernel<<<g, 256, 0, s>>>(da);
    h->launches++;
    if (grid_out) *grid_out = g;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

// several plain (no L2 / covariance term) dense applies in one launch; tables[k] == NULL entries are skipped
int crb_dense_tables_apply(crb_handle* h, int n, const crb_table* const* tables, float* const* grads, int opt_kind, const OptDev& od, cudaStream_t s) {
    DenseApplyMulti m;
    int used = 0, blocks = 0;
    for (int k = 0; k < n; ++k) {
        if (!tables[k]) continue;
        if (used == 4) { crb_set_error("crb_dense_tables_apply: at most four tables"); return CRB_ERR_ARG; }
        int rc = check_dense_table(tables[k], grads[k], opt_kind, "dense table");
        if (rc) return rc;
        DenseApplyArgs& da = m.t[used];
        da.dim = tables[k]->dim; da.opt_kind = dense_opt_kind(opt_kind); da.opt = od; da.cov = 0.f; da.mean = nullptr; da.mean_sum = nullptr;
        da.loss_cov = 0.f; da.l2 = 0.f; da.loss_l2 = 0.f;
        da.T = {tables[k]->w, tables[k]->s1, tables[k]->s2, grads[k], tables[k]->rows}; da.loss_part = h->dense_loss;
        m.first[used] = blocks;
        blocks += dgrid(h, tables[k]->rows, 8);
        ++used;
    }
    if (!used) return CRB_OK;
    for (int k = used; k <= 4; ++k) m.first[k] = blocks;
    for (int k = used; k < 4; ++k) m.t[k] = m.t[used - 1];
    if (blocks > 4 * h->loss_blocks) { crb_set_error("crb_dense_tables_apply: grid too large"); return CRB_ERR_ARG; }
    dense_tables_apply_kernel<<<blocks, 256, 0, s>>>(m);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

// ------------------------------------------------------------------------------------------------ SBPR
// model/ranking/SBPR.py:38-57.  x_ui = p.q_i + b_i, x_uk (a friend's item), x_uj (unobserved);
//   loss = sum softplus(-(x_ui - x_uk) / s_uk) + softplus(-(x_uk - x_uj))                                  (:54, utils/tools.py:71)
//        + reg * (l2(p) + l2(q_i) + l2(q_k) + l2(q_j) + l2(b_i) + l2(b_k) + l2(b_j))                       (:55-56)
// All four gathers are IndexedSlices in TF; the tables take the dense apply (zero gradient on untouched rows: exactly
// tf.train.AdamOptimizer's sparse apply, a no-op for SGD / Adagrad), the gradients are accumulated in the dense buffers.
struct SbprArgs {
    const float* P;
    const float* Q;
    const float* b;
    float* gP;
    float* gQ;
    float* gb;
    const int32_t* u;
    const int32_t* i;
    const int32_t* k;
    const int32_t* j;
    const float* suk;
    int64_t batch;
    int dim;
    float reg;
    double* loss_part;
};

template <int LANES, int VPL>
__global__ void __launch_bounds__(256) sbpr_step_kernel(SbprArgs a) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31, gl = lane % LANES, sub = lane / LANES;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double loss = 0.0;
    for (int64_t base = warp * GPW; base < a.batch; base += n_warps * GPW) {
        const int64_t t = base + sub;
        const bool active = t < a.batch;
        const int64_t tt = active ? t : a.batch - 1;
        const int64_t u = a.u[tt], it = a.i[tt], kt = a.k[tt], jt = a.j[tt];
        const float inv_s = __fdividef(1.f, a.suk[tt]);
        float4 p[VPL], qi[VPL], qk[VPL], qj[VPL];
        float xi = 0.f, xk = 0.f, xj = 0.f, sq = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            p[v] = c < a.dim ? ld4(a.P + u * a.dim + c) : z;
            qi[v] = c < a.dim ? ld4(a.Q + it * a.dim + c) : z;
            qk[v] = c < a.dim ? ld4(a.Q + kt * a.dim + c) : z;
            qj[v] = c < a.dim ? ld4(a.Q + jt * a.dim + c) : z;
            xi += dot4(p[v], qi[v]); xk += dot4(p[v], qk[v]); xj += dot4(p[v], qj[v]);
            sq += dot4(p[v], p[v]) + dot4(qi[v], qi[v]) + dot4(qk[v], qk[v]) + dot4(qj[v], qj[v]);
        }
        const float bi = a.b[it], bk = a.b[kt], bj = a.b[jt];
        xi = group_sum<LANES>(xi) + bi;
        xk = group_sum<LANES>(xk) + bk;
        xj = group_sum<LANES>(xj) + bj;
        sq = group_sum<LANES>(sq) + bi * bi + bk * bk + bj * bj;
        const float x1 = (xi - xk) * inv_s, x2 = xk - xj;
        const float g1 = -sigmoid_f(-x1) * inv_s, g2 = -sigmoid_f(-x2);   // dL/dx_ui = g1, dL/dx_uk = g2 - g1, dL/dx_uj = -g2
        const float gk = g2 - g1;
        if (active) {
            if (gl == 0) {
                loss += (double)(softplus_neg(x1) + softplus_neg(x2) + a.reg * 0.5f * sq);
                atomicAdd(a.gb + it, fmaf(a.reg, bi, g1));
                atomicAdd(a.gb + kt, fmaf(a.reg, bk, gk));
                atomicAdd(a.gb + jt, fmaf(a.reg, bj, -g2));
            }
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c = (gl + LANES * v) * 4;
                if (c >= a.dim) continue;
                const float4 P_ = p[v], I_ = qi[v], K_ = qk[v], J_ = qj[v];
                atomic_add4(a.gP + u * a.dim + c, make_float4(fmaf(a.reg, P_.x, g1 * I_.x + gk * K_.x - g2 * J_.x), fmaf(a.reg, P_.y, g1 * I_.y + gk * K_.y - g2 * J_.y),
                                                               fmaf(a.reg, P_.z, g1 * I_.z + gk * K_.z - g2 * J_.z), fmaf(a.reg, P_.w, g1 * I_.w + gk * K_.w - g2 * J_.w)));
                atomic_add4(a.gQ + it * a.dim + c, make_float4(fmaf(a.reg, I_.x, g1 * P_.x), fmaf(a.reg, I_.y, g1 * P_.y), fmaf(a.reg, I_.z, g1 * P_.z), fmaf(a.reg, I_.w, g1 * P_.w)));
                atomic_add4(a.gQ + kt * a.dim + c, make_float4(fmaf(a.reg, K_.x, gk * P_.x), fmaf(a.reg, K_.y, gk * P_.y), fmaf(a.reg, K_.z, gk * P_.z), fmaf(a.reg, K_.w, gk * P_.w)));
                atomic_add4(a.gQ + jt * a.dim + c, make_float4(fmaf(a.reg, J_.x, -g2 * P_.x), fmaf(a.reg, J_.y, -g2 * P_.y), fmaf(a.reg, J_.z, -g2 * P_.z), fmaf(a.reg, J_.w, -g2 * P_.w)));
            }
        }
    }
    block_sum_to(loss, a.loss_part);
}

template <int LANES, int VPL>
static int launch_sbpr_t(crb_handle* h, const SbprArgs& a, int grid, cudaStream_t s) {
    sbpr_step_kernel<LANES, VPL><<<grid, 256, 0, s>>>(a);
    h->launches++;
    return CRB_OK;
}

extern "C" int crb_train_step_sbpr(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_table* B, float* gradP, float* gradQ, float* gradB,
                                   const crb_opt* opt, const int32_t* u, const int32_t* i, const int32_t* k, const int32_t* j, const float* suk,
                                   int64_t batch, float reg, double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && u && i && k && j && suk, "null argument");
    CRB_CHECK_ARG(batch > 0, "batch");
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    if ((rc = check_dense_table(P, gradP, opt_kind, "P"))) return rc;
    if ((rc = check_dense_table(Q, gradQ, opt_kind, "Q"))) return rc;
    CRB_CHECK_ARG(B && B->w && B->dim == 1 && gradB && B->rows % 4 == 0, "bias table must have dim 1 and a length padded to a multiple of 4");
    CRB_CHECK_ARG(P->dim == Q->dim, "P.dim != Q.dim");
    CRB_CUDA(cudaSetDevice(h->device));
    const int dim = P->dim;
    if ((rc = crb_ws_reserve(h, batch, dim, 4, s))) return rc;
    const int32_t *du, *di, *dk, *dj;
    if ((rc = stage_dev_i32(h, u, batch, h->idx[0], &du, s))) return rc;
    if ((rc = stage_dev_i32(h, i, batch, h->idx[1], &di, s))) return rc;
    if ((rc = stage_dev_i32(h, k, batch, h->idx[2], &dk, s))) return rc;
    if ((rc = stage_dev_i32(h, j, batch, h->idx[3], &dj, s))) return rc;
    const float* ds = suk;
    if (!crb_is_device_ptr(suk)) { CRB_CUDA(cudaMemcpyAsync(h->yv, suk, sizeof(float) * batch, cudaMemcpyHostToDevice, s)); ds = h->yv; }
    const int grid = dgrid(h, batch, 256 / 32);
    SbprArgs a = {P->w, Q->w, B->w, gradP, gradQ, gradB, du, di, dk, dj, ds, batch, dim, reg, h->block_loss};
    if ((rc = crb_prof_begin(h, s))) return rc;
    CRB_DIM_DISPATCH(dim, launch_sbpr_t, h, a, grid, s);
    if ((rc = crb_prof_end(h, s))) return rc;
    double* dp = h->dense_loss;
    const int dk_ = dense_opt_kind(opt_kind);
    crb_table B4 = *B;      // the bias vector as a [n/4, 4] table
    B4.rows = B->rows / 4; B4.dim = 4;
    const crb_table* tabs3[3] = {P, Q, &B4};
    float* grads3[3] = {gradP, gradQ, gradB};
    (void)dp;
    if ((rc = crb_dense_tables_apply(h, 3, tabs3, grads3, dk_, od, s))) return rc;
    h->step_grid = grid;
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    if ((rc = crb_launch_loss_final(h, ld, s))) return rc;
    CRB_CUDA(cudaGetLastError());
    return finish_loss_host(h, loss_out, s);
}
