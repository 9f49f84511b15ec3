// NeuMF (reference: model/ranking/NeuMF.py:58-105, MLP tower MLP.py:44-53): GMF branch (p_g * q_g) + MLP tower over
// [p_m || q_m] with layers [L0, L0/2, ...] (W_k: [layers[k], layers[k]/2], ReLU), fused by h_neumf; pointwise loss.
// One warp per sample; tower weights live in shared memory (rows padded by one float so that both the forward column
// walk and the backward row walk are bank-conflict free); dense-parameter gradients are accumulated per CTA in shared
// memory and written as per-CTA partials (summed in CTA order by dense_vector_apply_kernel); embedding gradients are
// added into dense gradient buffers that the dense table apply consumes (TF's sparse Adam is dense in the moments and
// moves every row, SURVEY 2.4, so a dense apply with zero gradient on untouched rows is its exact semantics).
#include "rowopt.cuh"

#define NM_MAX_LAYERS 4
#define NM_WARPS 8

struct NeumfShape {
    int E;           // GMF embedding size
    int L0;          // MLP input width (2 * MLP embedding size)
    int n_layers;
    int n_in[NM_MAX_LAYERS], n_out[NM_MAX_LAYERS];
    int w_off[NM_MAX_LAYERS], b_off[NM_MAX_LAYERS];   // offsets in the packed dense vector
    int sw_off[NM_MAX_LAYERS];                        // offsets in the padded shared copy
    int h_off;       // h_neumf offset (packed)
    int n_dense;     // packed length
    int n_smem_w;    // padded weight floats
    int act_off[NM_MAX_LAYERS + 1];                   // per-warp activation offsets
    int act_total;
};

static int make_shape(int E, int L0, int n_layers, NeumfShape* s) {
    if (n_layers < 1 || n_layers > NM_MAX_LAYERS || L0 < 2 || (L0 % (1 << n_layers)) != 0 || L0 > 512 || E < 0 || E > 256) return -1;   // E == 0: the MLP model (MLP.py), no GMF branch
    s->E = E; s->L0 = L0; s->n_layers = n_layers;
    int off = 0, soff = 0, n = L0, aoff = 0;
    s->act_off[0] = 0; aoff = L0;
    for (int l = 0; l < n_layers; ++l) {
        s->n_in[l] = n; s->n_out[l] = n / 2;
        s->w_off[l] = off; off += n * (n / 2);
        s->b_off[l] = off; off += n / 2;
        s->sw_off[l] = soff; soff += n * (n / 2 + 1);
        n /= 2;
        s->act_off[l + 1] = aoff; aoff += n;
    }
    s->h_off = off; off += E + n;
    s->n_dense = off; s->n_smem_w = soff; s->act_total = aoff;
    return 0;
}

struct NeumfArgs {
    NeumfShape sh;
    const float *Pg, *Qg, *Pm, *Qm;
    float *gPg, *gQg, *gPm, *gQm;
    const float* dense;     // packed W_k, b_k, h_neumf
    float* dense_part;      // [gridDim.x, n_dense]
    const int32_t* u;
    const int32_t* i;
    const float* y;
    int64_t batch;
    int loss_kind;
    float reg1, reg2;
    double* loss_part;
};

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// shared layout: [W padded][biases + h (n_dense - sum W)] [gW padded][g biases + h] [per-warp: acts | deltas(2 x L0) | gmf u,i (2E)]
__global__ void __launch_bounds__(NM_WARPS * 32) neumf_step_kernel(NeumfArgs a) {
    extern __shared__ float sm[];
    const NeumfShape& S = a.sh;
    const int n_small = S.n_dense - S.h_off + 0;  // h
    int n_bias = 0;
    for (int l = 0; l < S.n_layers; ++l) n_bias += S.n_out[l];
    float* sW = sm;
    float* sB = sW + S.n_smem_w;                 // biases of all layers, contiguous in layer order
    float* sH = sB + n_bias;                     // h_neumf
    float* gW = sH + n_small;
    float* gB = gW + S.n_smem_w;
    float* gH = gB + n_bias;
    float* warp_base = gH + n_small;
    const int per_warp = S.act_total + 2 * S.L0 + 2 * S.E;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* act = warp_base + warp * per_warp;
    float* dA = act + S.act_total;               // delta ping
    float* dB = dA + S.L0;                       // delta pong
    float* eg = dB + S.L0;                       // ug | ig
    // ---- load parameters, zero gradient accumulators
    int boff = 0;
    for (int l = 0; l < S.n_layers; ++l) {
        const int ni = S.n_in[l], no = S.n_out[l];
        for (int k = threadIdx.x; k < ni * no; k += blockDim.x) sW[S.sw_off[l] + (k / no) * (no + 1) + (k % no)] = a.dense[S.w_off[l] + k];
        for (int k = threadIdx.x; k < no; k += blockDim.x) sB[boff + k] = a.dense[S.b_off[l] + k];
        boff += no;
    }
    for (int k = threadIdx.x; k < n_small; k += blockDim.x) sH[k] = a.dense[S.h_off + k];
    for (int k = threadIdx.x; k < S.n_smem_w + n_bias + n_small; k += blockDim.x) gW[k] = 0.f;
    __syncthreads();
    const int Em = S.L0 / 2;
    double loss = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * NM_WARPS + warp; t < a.batch; t += (int64_t)gridDim.x * NM_WARPS) {
        const int64_t u = a.u[t], it = a.i[t];
        const float y = a.y[t];
        float sq1 = 0.f, sq2 = 0.f, lg = 0.f;
        for (int k = lane; k < S.E; k += 32) {
            const float p = a.Pg[u * S.E + k], q = a.Qg[it * S.E + k];
            eg[k] = p; eg[S.E + k] = q;
            sq1 = fmaf(p, p, fmaf(q, q, sq1));
            lg = fmaf(p * q, sH[k], lg);
        }
        for (int k = lane; k < S.L0; k += 32) {
            const float x = k < Em ? a.Pm[u * Em + k] : a.Qm[it * Em + (k - Em)];
            act[k] = x;
            sq2 = fmaf(x, x, sq2);
        }
        __syncwarp();
        // ---- forward tower
        int bo = 0;
        for (int l = 0; l < S.n_layers; ++l) {
            const int ni = S.n_in[l], no = S.n_out[l];
            const float* x = act + S.act_off[l];
            const float* W = sW + S.sw_off[l];
            for (int o = lane; o < no; o += 32) {
                float acc = sB[bo + o];
                for (int k = 0; k < ni; ++k) acc = fmaf(x[k], W[k * (no + 1) + o], acc);
                act[S.act_off[l + 1] + o] = fmaxf(acc, 0.f);
            }
            bo += no;
            __syncwarp();
        }
        const int nl = S.n_out[S.n_layers - 1];
        const float* alast = act + S.act_off[S.n_layers];
        for (int k = lane; k < nl; k += 32) lg = fmaf(alast[k], sH[S.E + k], lg);
        const float logit = warp_sum(lg);
        sq1 = warp_sum(sq1);
        sq2 = warp_sum(sq2);
        float g, lv;
        if (a.loss_kind == CRB_LOSS_CROSS_ENTROPY) {
            lv = fmaxf(logit, 0.f) - logit * y + __logf(1.f + __expf(-fabsf(logit)));
            g = sigmoid_f(logit) - y;
        } else {
            lv = (y - logit) * (y - logit);
            g = 2.f * (logit - y);
        }
        if (lane == 0) loss += (double)(lv + a.reg1 * 0.5f * sq1 + a.reg2 * 0.5f * sq2);
        // ---- backward: fusion vector + GMF branch
        for (int k = lane; k < S.E; k += 32) {
            const float p = eg[k], q = eg[S.E + k], hk = sH[k];
            atomicAdd(gH + k, g * p * q);
            atomicAdd(a.gPg + u * S.E + k, fmaf(g * hk, q, a.reg1 * p));
            atomicAdd(a.gQg + it * S.E + k, fmaf(g * hk, p, a.reg1 * q));
        }
        float* delta = dA;
        float* dnext = dB;
        for (int k = lane; k < nl; k += 32) {
            atomicAdd(gH + S.E + k, g * alast[k]);
            delta[k] = alast[k] > 0.f ? g * sH[S.E + k] : 0.f;
        }
        __syncwarp();
        // ---- backward tower
        for (int l = S.n_layers - 1; l >= 0; --l) {
            const int ni = S.n_in[l], no = S.n_out[l];
            const float* x = act + S.act_off[l];
            const float* W = sW + S.sw_off[l];
            float* GW = gW + S.sw_off[l];
            bo -= no;
            for (int o = lane; o < no; o += 32) {
                const float dl = delta[o];
                atomicAdd(gB + bo + o, dl);
                if (dl != 0.f)
                    for (int k = 0; k < ni; ++k) atomicAdd(GW + k * (no + 1) + o, x[k] * dl);
            }
            for (int k = lane; k < ni; k += 32) {
                float acc = 0.f;
                for (int o = 0; o < no; ++o) acc = fmaf(W[k * (no + 1) + o], delta[o], acc);
                dnext[k] = (l > 0) ? (x[k] > 0.f ? acc : 0.f) : acc;   // x is a ReLU output for l > 0
            }
            __syncwarp();
            float* tmp = delta; delta = dnext; dnext = tmp;
        }
        for (int k = lane; k < S.L0; k += 32) {
            const float x = act[k];
            const float gx = fmaf(a.reg2, x, delta[k]);
            if (k < Em) atomicAdd(a.gPm + u * Em + k, gx); else atomicAdd(a.gQm + it * Em + (k - Em), gx);
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- per-CTA partial of the dense gradient, in packed layout
    float* part = a.dense_part + (int64_t)blockIdx.x * S.n_dense;
    boff = 0;
    for (int l = 0; l < S.n_layers; ++l) {
        const int ni = S.n_in[l], no = S.n_out[l];
        for (int k = threadIdx.x; k < ni * no; k += blockDim.x) part[S.w_off[l] + k] = gW[S.sw_off[l] + (k / no) * (no + 1) + (k % no)];
        for (int k = threadIdx.x; k < no; k += blockDim.x) part[S.b_off[l] + k] = gB[boff + k];
        boff += no;
    }
    for (int k = threadIdx.x; k < n_small; k += blockDim.x) part[S.h_off + k] = gH[k];
    // loss partial
    __shared__ double sl[NM_WARPS];
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) sl[warp] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tsum = 0.0;
        for (int k = 0; k < NM_WARPS; ++k) tsum += sl[k];
        a.loss_part[blockIdx.x] = tsum;
    }
}

// ------------------------------------------------------------------------------------------------ tiled variant (common towers)
// The same step for towers with L0 in {32, 64, 128} (conf/NeuMF.properties and conf/MLP.properties ship [128,64,32]): the weight
// gradients dW_l = sum_t x_t (x) delta_t are NOT accumulated with shared-memory atomics (10 752 per sample for the shipped tower --
// 85 % of the generic kernel's time) but in REGISTERS: a CTA processes tiles of NM_WARPS samples, every warp runs its sample's forward
// and delta back-propagation as before and leaves x_l and delta_l in shared memory; after a block barrier thread `tid` adds the tile's
// outer products into the dW elements it owns (column o = tid % n_out, rows k0 + j * 256 / n_out: x broadcasts across a warp, delta is
// conflict free).  Fixed tile and warp order -> the weight gradients are deterministic.  Template parameters make every loop static.
template <int L0, int NL> struct TileShape {
    static constexpr int n_in(int l) { return L0 >> l; }
    static constexpr int n_out(int l) { return L0 >> (l + 1); }
    static constexpr int per(int l) { return 256 / n_out(l) < n_in(l) ? 256 / n_out(l) : n_in(l); }   // rows covered by one pass of the CTA
    static constexpr int cnt(int l) { return (n_in(l) + per(l) - 1) / per(l); }
    static constexpr int acc_off(int l) { return l == 0 ? 0 : acc_off(l - 1) + cnt(l - 1); }
    static constexpr int n_acc = acc_off(NL);
};

// layer LIDX's share of a tile: thread `tid` owns dW[k0 + j*per][o]; every index into accW is a compile-time constant
template <int L0, int NL, int LIDX>
__device__ __forceinline__ void tile_accumulate(float (&accW)[TileShape<L0, NL>::n_acc], const NeumfShape& S, const float* warp_base, int per_warp,
                                                int n_act, int tid) {
    using TS = TileShape<L0, NL>;
    constexpr int ni = TS::n_in(LIDX), no = TS::n_out(LIDX), per = TS::per(LIDX), cnt = TS::cnt(LIDX), base = TS::acc_off(LIDX);
    const int o = tid % no, k0 = tid / no;
    if (k0 >= per) return;
    for (int w = 0; w < n_act; ++w) {
        const float* wa = warp_base + w * per_warp;
        const float* x = wa + S.act_off[LIDX];
        const float dv = (wa + S.act_total)[S.act_off[LIDX + 1] - L0 + o];
#pragma unroll
        for (int j = 0; j < cnt; ++j) {
            const int k = k0 + j * per;
            if (k < ni) accW[base + j] = fmaf(x[k], dv, accW[base + j]);
        }
    }
}

template <int L0, int NL, int LIDX>
__device__ __forceinline__ void tile_store(const float (&accW)[TileShape<L0, NL>::n_acc], const NeumfShape& S, float* part, int tid) {
    using TS = TileShape<L0, NL>;
    constexpr int ni = TS::n_in(LIDX), no = TS::n_out(LIDX), per = TS::per(LIDX), cnt = TS::cnt(LIDX), base = TS::acc_off(LIDX);
    const int o = tid % no, k0 = tid / no;
    if (k0 >= per) return;
#pragma unroll
    for (int j = 0; j < cnt; ++j) {
        const int k = k0 + j * per;
        if (k < ni) part[S.w_off[LIDX] + k * no + o] = accW[base + j];
    }
}

template <int L0, int NL>
__global__ void __launch_bounds__(NM_WARPS * 32) neumf_tile_kernel(NeumfArgs a) {
    using TS = TileShape<L0, NL>;
    extern __shared__ float sm[];
    const NeumfShape& S = a.sh;
    const int n_small = S.n_dense - S.h_off;
    int n_bias = 0;
#pragma unroll
    for (int l = 0; l < NL; ++l) n_bias += TS::n_out(l);
    float* sW = sm;
    float* sB = sW + S.n_smem_w;
    float* sH = sB + n_bias;
    float* gB = sH + n_small;
    float* gH = gB + n_bias;
    float* warp_base = gH + n_small;
    // per warp: activations (act_total) | deltas of every layer's output (act_total - L0) | gradient of the tower input (L0) | GMF rows (2E)
    const int per_warp = 2 * S.act_total + 2 * S.E;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    float* act = warp_base + warp * per_warp;
    float* dall = act + S.act_total;              // delta of layer l's output at dall[act_off[l+1] - L0 ...]
    float* din = dall + (S.act_total - L0);
    float* eg = din + L0;
    int boff = 0;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int ni = TS::n_in(l), no = TS::n_out(l);
        for (int k = tid; k < ni * no; k += blockDim.x) sW[S.sw_off[l] + (k / no) * (no + 1) + (k % no)] = a.dense[S.w_off[l] + k];
        for (int k = tid; k < no; k += blockDim.x) sB[boff + k] = a.dense[S.b_off[l] + k];
        boff += no;
    }
    for (int k = tid; k < n_small; k += blockDim.x) sH[k] = a.dense[S.h_off + k];
    for (int k = tid; k < n_bias + n_small; k += blockDim.x) gB[k] = 0.f;
    float accW[TS::n_acc];
#pragma unroll
    for (int q = 0; q < TS::n_acc; ++q) accW[q] = 0.f;
    __syncthreads();
    constexpr int Em = L0 / 2;
    double loss = 0.0;
    const int64_t n_tiles = (a.batch + NM_WARPS - 1) / NM_WARPS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t = tile * NM_WARPS + warp;
        const int64_t left = a.batch - tile * NM_WARPS;
        const int n_act = left < NM_WARPS ? (int)left : NM_WARPS;
        if (warp < n_act) {
            const int64_t u = a.u[t], it = a.i[t];
            const float y = a.y[t];
            float sq1 = 0.f, sq2 = 0.f, lg = 0.f;
            for (int k = lane; k < S.E; k += 32) {
                const float p = a.Pg[u * S.E + k], q = a.Qg[it * S.E + k];
                eg[k] = p; eg[S.E + k] = q;
                sq1 = fmaf(p, p, fmaf(q, q, sq1));
                lg = fmaf(p * q, sH[k], lg);
            }
#pragma unroll
            for (int k = lane; k < L0; k += 32) {
                const float x = k < Em ? a.Pm[u * Em + k] : a.Qm[it * Em + (k - Em)];
                act[k] = x;
                sq2 = fmaf(x, x, sq2);
            }
            __syncwarp();
            int bo = 0;
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                constexpr int dummy = 0; (void)dummy;
                const int ni = TS::n_in(l), no = TS::n_out(l);
                const float* x = act + S.act_off[l];
                const float* W = sW + S.sw_off[l];
                for (int o = lane; o < no; o += 32) {
                    float acc = sB[bo + o];
#pragma unroll 8
                    for (int k = 0; k < ni; ++k) acc = fmaf(x[k], W[k * (no + 1) + o], acc);
                    act[S.act_off[l + 1] + o] = fmaxf(acc, 0.f);
                }
                bo += no;
                __syncwarp();
            }
            constexpr int nl = TS::n_out(NL - 1);
            const float* alast = act + S.act_off[NL];
            for (int k = lane; k < nl; k += 32) lg = fmaf(alast[k], sH[S.E + k], lg);
            const float logit = warp_sum(lg);
            sq1 = warp_sum(sq1);
            sq2 = warp_sum(sq2);
            float g, lv;
            if (a.loss_kind == CRB_LOSS_CROSS_ENTROPY) {
                lv = fmaxf(logit, 0.f) - logit * y + __logf(1.f + __expf(-fabsf(logit)));
                g = sigmoid_f(logit) - y;
            } else {
                lv = (y - logit) * (y - logit);
                g = 2.f * (logit - y);
            }
            if (lane == 0) loss += (double)(lv + a.reg1 * 0.5f * sq1 + a.reg2 * 0.5f * sq2);
            for (int k = lane; k < S.E; k += 32) {
                const float p = eg[k], q = eg[S.E + k], hk = sH[k];
                atomicAdd(gH + k, g * p * q);
                atomicAdd(a.gPg + u * S.E + k, fmaf(g * hk, q, a.reg1 * p));
                atomicAdd(a.gQg + it * S.E + k, fmaf(g * hk, p, a.reg1 * q));
            }
            float* dlast = dall + (S.act_off[NL] - L0);
            for (int k = lane; k < nl; k += 32) {
                atomicAdd(gH + S.E + k, g * alast[k]);
                dlast[k] = alast[k] > 0.f ? g * sH[S.E + k] : 0.f;
            }
            __syncwarp();
            // delta back-propagation (no weight-gradient work here): delta_{l-1}[k] = relu'(x_l[k]) * sum_o W_l[k][o] delta_l[o]
#pragma unroll
            for (int l = NL - 1; l >= 0; --l) {
                const int ni = TS::n_in(l), no = TS::n_out(l);
                const float* x = act + S.act_off[l];
                const float* W = sW + S.sw_off[l];
                const float* delta = dall + (S.act_off[l + 1] - L0);
                float* dnext = l > 0 ? dall + (S.act_off[l] - L0) : din;
                bo -= no;
                for (int o = lane; o < no; o += 32) atomicAdd(gB + bo + o, delta[o]);
                for (int k = lane; k < ni; k += 32) {
                    float acc = 0.f;
#pragma unroll 8
                    for (int o = 0; o < no; ++o) acc = fmaf(W[k * (no + 1) + o], delta[o], acc);
                    dnext[k] = (l > 0) ? (x[k] > 0.f ? acc : 0.f) : acc;
                }
                __syncwarp();
            }
#pragma unroll
            for (int k = lane; k < L0; k += 32) {
                const float gx = fmaf(a.reg2, act[k], din[k]);
                if (k < Em) atomicAdd(a.gPm + u * Em + k, gx); else atomicAdd(a.gQm + it * Em + (k - Em), gx);
            }
        }
        __syncthreads();
        // ---- the tile's outer products into the registers that own dW
        tile_accumulate<L0, NL, 0>(accW, S, warp_base, per_warp, n_act, tid);
        if constexpr (NL > 1) tile_accumulate<L0, NL, 1>(accW, S, warp_base, per_warp, n_act, tid);
        if constexpr (NL > 2) tile_accumulate<L0, NL, 2>(accW, S, warp_base, per_warp, n_act, tid);
        if constexpr (NL > 3) tile_accumulate<L0, NL, 3>(accW, S, warp_base, per_warp, n_act, tid);
        __syncthreads();
    }
    // ---- per-CTA partial of the dense gradient, in packed layout
    float* part = a.dense_part + (int64_t)blockIdx.x * S.n_dense;
    tile_store<L0, NL, 0>(accW, S, part, tid);
    if constexpr (NL > 1) tile_store<L0, NL, 1>(accW, S, part, tid);
    if constexpr (NL > 2) tile_store<L0, NL, 2>(accW, S, part, tid);
    if constexpr (NL > 3) tile_store<L0, NL, 3>(accW, S, part, tid);
    boff = 0;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int no = TS::n_out(l);
        for (int k = tid; k < no; k += blockDim.x) part[S.b_off[l] + k] = gB[boff + k];
        boff += no;
    }
    for (int k = tid; k < n_small; k += blockDim.x) part[S.h_off + k] = gH[k];
    __shared__ double sl[NM_WARPS];
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) sl[warp] = loss;
    __syncthreads();
    if (tid == 0) {
        double tsum = 0.0;
        for (int k = 0; k < NM_WARPS; ++k) tsum += sl[k];
        a.loss_part[blockIdx.x] = tsum;
    }
}

template <int L0, int NL>
static int launch_neumf_tile(const NeumfArgs& a, int grid, cudaStream_t s) {
    const NeumfShape& S = a.sh;
    int n_bias = 0;
    for (int l = 0; l < NL; ++l) n_bias += S.n_out[l];
    const int n_small = S.n_dense - S.h_off;
    const size_t smem = sizeof(float) * ((size_t)S.n_smem_w + 2 * (size_t)(n_bias + n_small) + (size_t)NM_WARPS * (2 * S.act_total + 2 * S.E));
    CRB_CUDA(cudaFuncSetAttribute(neumf_tile_kernel<L0, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    neumf_tile_kernel<L0, NL><<<grid, NM_WARPS * 32, smem, s>>>(a);
    return CRB_OK;
}

// -> true when a tiled instantiation exists for the tower (launches it)
static bool try_launch_neumf_tile(const NeumfArgs& a, int grid, cudaStream_t s, int* rc) {
#define CRB_TILE_CASE(L, N) if (a.sh.L0 == L && a.sh.n_layers == N) { *rc = launch_neumf_tile<L, N>(a, grid, s); return true; }
    CRB_TILE_CASE(128, 3) CRB_TILE_CASE(128, 2) CRB_TILE_CASE(128, 4)
    CRB_TILE_CASE(64, 3) CRB_TILE_CASE(64, 2) CRB_TILE_CASE(64, 4)
    CRB_TILE_CASE(32, 3) CRB_TILE_CASE(32, 2)
#undef CRB_TILE_CASE
    return false;
}

// TF dense apply of a packed dense vector from per-CTA partial gradients (fixed summation order)
__global__ void __launch_bounds__(256) dense_vector_apply_kernel(float* w, float* s1, float* s2, const float* parts, int n_parts, int n,
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

// shared with train_lrml.cu
int crb_dense_vector_apply(crb_handle* h, float* w, float* s1, float* s2, const float* parts, int n_parts, int n, int opt_kind, const OptDev& od,
                           cudaStream_t s) {
    dense_vector_apply_kernel<<<(n + 31) / 32, 256, 0, s>>>(w, s1, s2, parts, n_parts, n, opt_kind, od);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

// ------------------------------------------------------------------------------------------------ scoring
// canonical NeuMF logit of (user, item) pairs: per layer acc = b[o]; acc = fma(x[k], W[k][o], acc) for k ascending; ReLU;
// logit = fma chain over the GMF products (rounded first) then over the tower output.  One thread per pair.
struct NeumfScoreArgs {
    NeumfShape sh;
    const float *Pg, *Qg, *Pm, *Qm;
    const float* dense;
    const int32_t* u;
    const int32_t* i;
    int64_t n;
    float* out;
};

__global__ void __launch_bounds__(128) neumf_score_kernel(NeumfScoreArgs a) {
    extern __shared__ float sm[];
    const NeumfShape& S = a.sh;
    for (int k = threadIdx.x; k < S.n_dense; k += blockDim.x) sm[k] = a.dense[k];
    __syncthreads();
    const int Em = S.L0 / 2;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = a.u[t], it = a.i[t];
        float x[256], z[128];  // L0 <= 256 on this path (checked by the host)
        for (int k = 0; k < S.L0; ++k) x[k] = k < Em ? a.Pm[u * Em + k] : a.Qm[it * Em + (k - Em)];
        for (int l = 0; l < S.n_layers; ++l) {
            const int ni = S.n_in[l], no = S.n_out[l];
            const float* W = sm + S.w_off[l];
            const float* b = sm + S.b_off[l];
            for (int o = 0; o < no; ++o) {
                float acc = b[o];
                for (int k = 0; k < ni; ++k) acc = fmaf(x[k], W[k * no + o], acc);
                z[o] = fmaxf(acc, 0.f);
            }
            for (int o = 0; o < no; ++o) x[o] = z[o];
        }
        const float* h = sm + S.h_off;
        float acc = 0.f;
        for (int k = 0; k < S.E; ++k) acc = fmaf(__fmul_rn(a.Pg[u * S.E + k], a.Qg[it * S.E + k]), h[k], acc);
        const int nl = S.n_out[S.n_layers - 1];
        for (int k = 0; k < nl; ++k) acc = fmaf(x[k], h[S.E + k], acc);
        a.out[t] = acc;
    }
}

// The same canonical logit, one WARP per pair: lane o owns output o of a layer and runs its sequential k-chain out of shared memory
// (weights padded like the training kernel's), so every layer output -- and the final chain over the GMF products and the tower output,
// run by lane 0 alone in the same order -- is bit-identical to neumf_score_kernel.  No per-thread arrays in local memory: ~100x the
// pair rate of the thread-per-pair version, which is kept for validation (CRB_NEUMF_SCORE_SIMPLE).
__global__ void __launch_bounds__(NM_WARPS * 32) neumf_score_warp_kernel(NeumfScoreArgs a) {
    extern __shared__ float sm[];
    const NeumfShape& S = a.sh;
    const int n_small = S.n_dense - S.h_off;
    int n_bias = 0;
    for (int l = 0; l < S.n_layers; ++l) n_bias += S.n_out[l];
    float* sW = sm;
    float* sB = sW + S.n_smem_w;
    float* sH = sB + n_bias;
    float* warp_base = sH + n_small;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* act = warp_base + warp * (S.act_total + S.E);
    float* eg = act + S.act_total;   // rounded GMF products p * q
    int boff = 0;
    for (int l = 0; l < S.n_layers; ++l) {
        const int ni = S.n_in[l], no = S.n_out[l];
        for (int k = threadIdx.x; k < ni * no; k += blockDim.x) sW[S.sw_off[l] + (k / no) * (no + 1) + (k % no)] = a.dense[S.w_off[l] + k];
        for (int k = threadIdx.x; k < no; k += blockDim.x) sB[boff + k] = a.dense[S.b_off[l] + k];
        boff += no;
    }
    for (int k = threadIdx.x; k < n_small; k += blockDim.x) sH[k] = a.dense[S.h_off + k];
    __syncthreads();
    const int Em = S.L0 / 2;
    for (int64_t t = (int64_t)blockIdx.x * NM_WARPS + warp; t < a.n; t += (int64_t)gridDim.x * NM_WARPS) {
        const int64_t u = a.u[t], it = a.i[t];
        for (int k = lane; k < S.E; k += 32) eg[k] = __fmul_rn(a.Pg[u * S.E + k], a.Qg[it * S.E + k]);
        for (int k = lane; k < S.L0; k += 32) act[k] = k < Em ? a.Pm[u * Em + k] : a.Qm[it * Em + (k - Em)];
        __syncwarp();
        int bo = 0;
        for (int l = 0; l < S.n_layers; ++l) {
            const int ni = S.n_in[l], no = S.n_out[l];
            const float* x = act + S.act_off[l];
            const float* W = sW + S.sw_off[l];
            for (int o = lane; o < no; o += 32) {
                float acc = sB[bo + o];
                for (int k = 0; k < ni; ++k) acc = fmaf(x[k], W[k * (no + 1) + o], acc);
                act[S.act_off[l + 1] + o] = fmaxf(acc, 0.f);
            }
            bo += no;
            __syncwarp();
        }
        if (lane == 0) {
            const int nl = S.n_out[S.n_layers - 1];
            const float* alast = act + S.act_off[S.n_layers];
            float acc = 0.f;
            for (int k = 0; k < S.E; ++k) acc = fmaf(eg[k], sH[k], acc);
            for (int k = 0; k < nl; ++k) acc = fmaf(alast[k], sH[S.E + k], acc);
            a.out[t] = acc;
        }
        __syncwarp();
    }
}

// scores[k, item] = value for every item the k-th user has seen (RankingRecommender.py:235-240 skip rule as a mask)
__global__ void __launch_bounds__(256) mask_seen_kernel(float* scores, const int32_t* users, int64_t n_users, int64_t n_items,
                                                       const int64_t* seen_rowptr, const int32_t* seen_cols, float value) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp; k < n_users; k += n_warps) {
        const int32_t u = users[k];
        for (int64_t p = seen_rowptr[u] + lane; p < seen_rowptr[u + 1]; p += 32) {
            const int32_t it = seen_cols[p];
            if (it < n_items) scores[k * n_items + it] = value;
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct DenseApplyArgsFwd;  // dense table apply lives in train_dense.cu
int crb_dense_tables_apply(crb_handle* h, int n, const crb_table* const* tables, float* const* grads, int opt_kind, const OptDev& od, cudaStream_t s);
int crb_dense_table_apply(crb_handle* h, const crb_table* T, float* grad, int opt_kind, const OptDev& od, float l2, double* loss_part,
                          int* grid_out, cudaStream_t s);

extern "C" int crb_train_step_neumf(crb_handle* h, const crb_table* Pg, const crb_table* Qg, const crb_table* Pm, const crb_table* Qm,
                                    float* gPg, float* gQg, float* gPm, float* gQm, float* dense, float* dense_s1, float* dense_s2,
                                    int32_t n_layers, const crb_opt* opt, int32_t loss_kind, const int32_t* u, const int32_t* i,
                                    const float* y, int64_t batch, float reg1, float reg2, double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && Pm && Qm && gPm && gQm && dense && u && i && y, "null argument");
    const bool gmf = Pg != nullptr;   // Pg == Qg == NULL: model/ranking/MLP.py (the tower alone, logit = h_mlp . tower)
    CRB_CHECK_ARG(gmf ? (Qg && gPg && gQg) : (!Qg && !gPg && !gQg), "the GMF branch tables must be all given or all NULL");
    CRB_CHECK_ARG(batch > 0, "batch");
    CRB_CHECK_ARG(loss_kind == CRB_LOSS_CROSS_ENTROPY || loss_kind == CRB_LOSS_SQUARE, "pointwise loss must be cross_entropy or square");
    CRB_CHECK_ARG((!gmf || Pg->dim == Qg->dim) && Pm->dim == Qm->dim, "table dims");
    NeumfArgs a;
    const int E_gmf = gmf ? Pg->dim : 0;
    if (make_shape(E_gmf, 2 * Pm->dim, n_layers, &a.sh)) { crb_set_error("unsupported NeuMF shape (E=%d, L0=%d, layers=%d)", E_gmf, 2 * Pm->dim, n_layers); return CRB_ERR_ARG; }
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    const int dk = opt_kind == OPT_ADAM_TF1 ? OPT_ADAM_LAZY : opt_kind;
    CRB_CHECK_ARG(dk == OPT_SGD || dense_s1, "dense slot s1 is NULL");
    CRB_CHECK_ARG(dk != OPT_ADAM_LAZY || dense_s2, "dense slot s2 is NULL");
    CRB_CUDA(cudaSetDevice(h->device));
    if ((rc = crb_ws_reserve(h, batch, 4, 4, s))) return rc;
    // towers with a tiled instantiation (register-resident weight gradients): up to three CTAs per SM, tiles of NM_WARPS samples
    const bool tiled = (a.sh.L0 == 128 || a.sh.L0 == 64 || a.sh.L0 == 32) && a.sh.n_layers >= 2 && a.sh.n_layers <= (a.sh.L0 == 32 ? 3 : 4) &&
                       !getenv("CRB_NEUMF_GENERIC");
    int grid = (int)((batch + 63) / 64);
    if (grid > h->sm_count * 2) grid = h->sm_count * 2;
    if (tiled) {
        // CTAs per SM: 79 registers x 256 threads and ~60 KB of shared memory leave room for three, and a CTA is a chain of tiles (8 warps,
        // one sample each, dependent dot-product chains) that needs company to hide its latencies: 1 -> 3 CTAs per SM took the NeuMF
        // epoch at the ml-1m shape from 136 to 109 ms (more CTAs = more per-CTA partials for the dense apply, which is why not more)
        int per_sm = 3;
        if (const char* e = getenv("CRB_NEUMF_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v <= 4) per_sm = v; }
        grid = (int)((batch + NM_WARPS - 1) / NM_WARPS);
        if (grid > h->sm_count * per_sm) grid = h->sm_count * per_sm;
    }
    if (grid < 1) grid = 1;
    // dense workspace: per-CTA partials
    const int64_t need = (int64_t)grid * a.sh.n_dense;
    if (need > h->cap_dense) {
        CRB_CUDA(cudaStreamSynchronize(s));
        cudaFree(h->dense_grad);
        h->dense_grad = nullptr; h->cap_dense = 0;
        CRB_CUDA(cudaMalloc(&h->dense_grad, sizeof(float) * need));
        h->cap_dense = need;
    }
    const int32_t *du = u, *di = i;
    const float* dy = y;
    if (!crb_is_device_ptr(u)) { CRB_CUDA(cudaMemcpyAsync(h->idx[0], u, 4 * batch, cudaMemcpyHostToDevice, s)); du = h->idx[0]; }
    if (!crb_is_device_ptr(i)) { CRB_CUDA(cudaMemcpyAsync(h->idx[1], i, 4 * batch, cudaMemcpyHostToDevice, s)); di = h->idx[1]; }
    if (!crb_is_device_ptr(y)) { CRB_CUDA(cudaMemcpyAsync(h->yv, y, 4 * batch, cudaMemcpyHostToDevice, s)); dy = h->yv; }
    a.Pg = gmf ? Pg->w : nullptr; a.Qg = gmf ? Qg->w : nullptr; a.Pm = Pm->w; a.Qm = Qm->w;
    a.gPg = gPg; a.gQg = gQg; a.gPm = gPm; a.gQm = gQm;
    a.dense = dense; a.dense_part = h->dense_grad; a.u = du; a.i = di; a.y = dy; a.batch = batch; a.loss_kind = loss_kind;
    a.reg1 = reg1; a.reg2 = reg2; a.loss_part = h->block_loss;
    int n_bias = 0;
    for (int l = 0; l < a.sh.n_layers; ++l) n_bias += a.sh.n_out[l];
    const int n_small = a.sh.n_dense - a.sh.h_off;
    const size_t smem = sizeof(float) * (2 * (size_t)(a.sh.n_smem_w + n_bias + n_small) + (size_t)NM_WARPS * (a.sh.act_total + 2 * a.sh.L0 + 2 * a.sh.E));
    if (smem > 200 * 1024) { crb_set_error("NeuMF tower too large for shared memory (%zu bytes)", smem); return CRB_ERR_UNSUPPORTED; }
    CRB_CUDA(cudaFuncSetAttribute(neumf_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if ((rc = crb_prof_begin(h, s))) return rc;
    int trc = CRB_OK;
    if (!tiled) neumf_step_kernel<<<grid, NM_WARPS * 32, smem, s>>>(a);
    else if (!try_launch_neumf_tile(a, grid, s, &trc)) { crb_set_error("internal: tiled NeuMF dispatch"); return CRB_ERR_ARG; }
    if (trc) return trc;
    if ((rc = crb_prof_end(h, s))) return rc;
    // embedding tables: dense apply (zero gradient on untouched rows == TF's sparse apply for SGD/Adagrad, and exactly
    // tf.train.AdamOptimizer's dense-in-the-moments sparse apply for Adam)
    const crb_table* tabs[4] = {Pg, Qg, Pm, Qm};      // NULL entries (the MLP model has no GMF branch) are skipped
    float* grads[4] = {gPg, gQg, gPm, gQm};
    if ((rc = crb_dense_tables_apply(h, 4, tabs, grads, dk, od, s))) return rc;
    dense_vector_apply_kernel<<<(a.sh.n_dense + 31) / 32, 256, 0, s>>>(dense, dense_s1, dense_s2, h->dense_grad, grid, a.sh.n_dense, dk, od);
    h->step_grid = grid;
    h->launches += 2;
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    if ((rc = crb_launch_loss_final(h, ld, s))) return rc;
    CRB_CUDA(cudaGetLastError());
    if (loss_out && !crb_is_device_ptr(loss_out)) {
        CRB_CUDA(cudaMemcpyAsync(loss_out, h->loss_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
        CRB_CUDA(cudaStreamSynchronize(s));
    }
    return CRB_OK;
}

extern "C" int crb_score_pairs_neumf(crb_handle* h, const float* Pg, const float* Qg, const float* Pm, const float* Qm, const float* dense,
                                     int32_t E, int32_t Em, int32_t n_layers, const int32_t* u, const int32_t* i, int64_t n, float* scores,
                                     void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && Pm && Qm && dense && u && i && scores, "null argument");
    CRB_CHECK_ARG(E == 0 ? (!Pg && !Qg) : (Pg && Qg), "GMF branch tables: both given (E > 0) or both NULL (E == 0, the MLP model)");
    CRB_CHECK_ARG(crb_is_device_ptr(u) && crb_is_device_ptr(i) && crb_is_device_ptr(scores), "u/i/scores must be device pointers");
    NeumfScoreArgs a;
    if (make_shape(E, 2 * Em, n_layers, &a.sh) || 2 * Em > 256) { crb_set_error("unsupported NeuMF shape"); return CRB_ERR_ARG; }
    if (n == 0) return CRB_OK;
    a.Pg = Pg; a.Qg = Qg; a.Pm = Pm; a.Qm = Qm; a.dense = dense; a.u = u; a.i = i; a.n = n; a.out = scores;
    int n_bias = 0;
    for (int l = 0; l < a.sh.n_layers; ++l) n_bias += a.sh.n_out[l];
    const size_t smem_w = sizeof(float) * ((size_t)a.sh.n_smem_w + n_bias + (a.sh.n_dense - a.sh.h_off) + (size_t)NM_WARPS * (a.sh.act_total + a.sh.E));
    if (smem_w <= 200 * 1024 && !getenv("CRB_NEUMF_SCORE_SIMPLE")) {
        CRB_CUDA(cudaFuncSetAttribute(neumf_score_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
        int64_t grid = (n + NM_WARPS - 1) / NM_WARPS;
        if (grid > (int64_t)h->sm_count * 2) grid = (int64_t)h->sm_count * 2;
        neumf_score_warp_kernel<<<(int)grid, NM_WARPS * 32, smem_w, s>>>(a);
    } else {
        const size_t smem = sizeof(float) * a.sh.n_dense;
        CRB_CUDA(cudaFuncSetAttribute(neumf_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t grid = (n + 127) / 128;
        if (grid > (int64_t)h->sm_count * 8) grid = (int64_t)h->sm_count * 8;
        neumf_score_kernel<<<(int)grid, 128, smem, s>>>(a);
    }
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

extern "C" int crb_mask_seen(crb_handle* h, float* scores, const int32_t* users, int64_t n_users, int64_t n_items, float value, void* stream) {
    CRB_CHECK_ARG(h && scores && users, "null argument");
    CRB_CHECK_ARG(crb_is_device_ptr(scores) && crb_is_device_ptr(users), "scores/users must be device pointers");
    if (!h->seen_rowptr) { crb_set_error("crb_mask_seen before crb_set_history"); return CRB_ERR_STATE; }
    if (n_users == 0) return CRB_OK;
    int64_t grid = (n_users + 7) / 8;
    if (grid > (int64_t)h->sm_count * 8) grid = (int64_t)h->sm_count * 8;
    mask_seen_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(scores, users, n_users, n_items, h->seen_rowptr, h->seen_cols, value);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}
