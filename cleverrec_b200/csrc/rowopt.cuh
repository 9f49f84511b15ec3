// Row-sparse optimizer application (TF-1 semantics, SURVEY.md 2.4 / utils/tools.py:79-87) and the
// "duplicate row" machinery shared by every row-sparse training step.
//
// One training step treats a table row in one of two ways, decided by its multiplicity in the batch
// (counted with one 32-bit atomic per occurrence into crb_handle::meta):
//   * multiplicity 1  -> the sample's own lane group applies the optimizer in place: the row is read once
//                        and written once, which is the algorithmic minimum (24*d / 48*d / 72*d bytes per
//                        triplet for SGD / Adagrad / Adam).
//   * multiplicity >1 -> TF de-duplicates IndexedSlices by summing and applies once per unique row, and all
//                        occurrences must see the pre-step value.  The occurrences write their gradient to
//                        slots [base, base+cnt); a second kernel sums the slots (in triplet order when the row
//                        has <= 32 occurrences, so the result is deterministic) and applies once.
#pragma once
#include "common.cuh"

struct TableDev {
    float* w;
    float* s1;
    float* s2;
    int32_t* last;
};

struct OptDev {
    float lr;      // SGD / Adagrad learning rate
    float lr_t;    // Adam: lr * sqrt(1-b2^t)/(1-b1^t) of this step
    float b1, b2, eps;
    float omb1, omb2;  // 1 - b1, 1 - b2 in fp32
    int32_t step;  // 1-based index of this step
    const float* lrt;  // lr_t table for replay (CRB_ADAM_TF1)
    const uint32_t* step_base;  // NULL, or (epoch graph, train.cu) a device word added to `step` when the kernel starts
};

enum { OPT_SGD = 0, OPT_ADAGRAD = 1, OPT_ADAM_LAZY = 2, OPT_ADAM_TF1 = 3 };

template <int OPT> struct OptTraits {
    static constexpr bool has_s1 = OPT != OPT_SGD;
    static constexpr bool has_s2 = OPT == OPT_ADAM_LAZY || OPT == OPT_ADAM_TF1;
    static constexpr bool replay = OPT == OPT_ADAM_TF1;
};

__device__ __forceinline__ float lrt_at(const OptDev& o, int s) { return s < CRB_LRT_TABLE ? o.lrt[s] : o.lr; }

// Kernels replayed from a captured graph carry a RELATIVE step; the absolute one (and Adam's lr_t, the same table entry the host
// would have passed) is resolved when the kernel starts.  A no-op for ordinary launches.
__device__ __forceinline__ void opt_resolve(OptDev& o) {
    if (o.step_base) {
        o.step += (int32_t)*o.step_base;
        o.lr_t = lrt_at(o, o.step);
    }
}

// Optimizer arithmetic.  Multiplies/adds are plain IEEE ops without contraction (the op-by-op sequence of TF's kernels);
// the square root and the division use the SFU approximations (sqrt.approx / rcp.approx, <= 2 ulp), which keeps the step
// kernel inside the instruction cache (the IEEE expansions made the CRB_ADAM_TF1 variant 4.5k instructions) and is far
// inside the 1e-4 parity tolerance.  The dense flush, the replay and the touched-row update share these functions, so a
// replayed row is bit-identical to a densely decayed one.
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// One decay-only Adam step (a row that was not touched at step s under tf.train.AdamOptimizer's sparse apply):
// m *= b1; v *= b2; w -= lr_s * m / (sqrt(v) + eps).
__device__ __forceinline__ void adam_decay_elem(float& w, float& m, float& v, float lr_s, float b1, float b2, float eps) {
    m = __fmul_rn(m, b1);
    v = __fmul_rn(v, b2);
    w = __fsub_rn(w, __fdividef(__fmul_rn(lr_s, m), __fadd_rn(sqrt_approx(v), eps)));
}

__device__ __forceinline__ void adam_decay4(float4& W, float4& M, float4& V, float lr_s, const OptDev& o) {
    adam_decay_elem(W.x, M.x, V.x, lr_s, o.b1, o.b2, o.eps);
    adam_decay_elem(W.y, M.y, V.y, lr_s, o.b1, o.b2, o.eps);
    adam_decay_elem(W.z, M.z, V.z, lr_s, o.b1, o.b2, o.eps);
    adam_decay_elem(W.w, M.w, V.w, lr_s, o.b1, o.b2, o.eps);
}

// tf.train.AdagradOptimizer sparse apply: acc += g*g; w -= lr * g / sqrt(acc)  (no epsilon, acc starts at 0.1)
__device__ __forceinline__ void adagrad_elem(float& w, float& acc, float g, float lr) {
    acc = __fadd_rn(acc, __fmul_rn(g, g));
    w = __fsub_rn(w, __fdividef(__fmul_rn(lr, g), sqrt_approx(acc)));
}

__device__ __forceinline__ void adam_touch_elem(float& w, float& m, float& v, float g, const OptDev& o) {
    m = __fadd_rn(__fmul_rn(m, o.b1), __fmul_rn(g, o.omb1));
    v = __fadd_rn(__fmul_rn(v, o.b2), __fmul_rn(__fmul_rn(g, g), o.omb2));
    w = __fsub_rn(w, __fdividef(__fmul_rn(o.lr_t, m), __fadd_rn(sqrt_approx(v), o.eps)));
}

// Per lane-group view of one row: VPL float4 chunks per lane, chunk index c = gl + LANES*v, valid iff 4c < dim.
template <int LANES, int VPL> struct RowRegs {
    float4 w[VPL];
    float4 s1[VPL];
    float4 s2[VPL];
    int32_t last;
};

template <int LANES, int VPL>
__device__ __forceinline__ void row_load_w(RowRegs<LANES, VPL>& r, const TableDev& T, int64_t row, int dim, int gl) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        int c = (gl + LANES * v) * 4;
        r.w[v] = c < dim ? ld4(T.w + row * dim + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int LANES, int VPL, int OPT>
__device__ __forceinline__ void row_load_state(RowRegs<LANES, VPL>& r, const TableDev& T, int64_t row, int dim, int gl) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        int c = (gl + LANES * v) * 4;
        // padding lanes (c >= dim): Adam m = 0 so a replayed decay step leaves their w at exactly 0; Adagrad acc = 1
        const float pad1 = OptTraits<OPT>::has_s2 ? 0.f : 1.f;
        if (OptTraits<OPT>::has_s1) r.s1[v] = c < dim ? ld4(T.s1 + row * dim + c) : make_float4(pad1, pad1, pad1, pad1);
        if (OptTraits<OPT>::has_s2) r.s2[v] = c < dim ? ld4(T.s2 + row * dim + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
}

// CRB_ADAM_TF1 bookkeeping.  `last` = step of the row's last update; 0 = never updated, i.e. m = v = 0 (the slots are
// created as zeros), for which every decay-only step is an exact no-op (w -= lr*0/(0+eps)) and is skipped.
template <int OPT>
__device__ __forceinline__ bool replay_pending(int last, const OptDev& o) {
    return OptTraits<OPT>::replay && last != 0 && last < o.step - 1;
}

// CRB_ADAM_TF1: apply the decay-only steps last+1 .. step-1 in registers so that r.w is the value the dense TF
// update would hold before this step.
template <int LANES, int VPL, int OPT>
__device__ __forceinline__ void row_replay(RowRegs<LANES, VPL>& r, const OptDev& o, int upto /*exclusive*/) {
    if (!OptTraits<OPT>::replay || r.last == 0) return;
    for (int s = r.last + 1; s < upto; ++s) {
        const float lr_s = lrt_at(o, s);
#pragma unroll
        for (int v = 0; v < VPL; ++v) adam_decay4(r.w[v], r.s1[v], r.s2[v], lr_s, o);
    }
}

// Apply this step's update with gradient g and store the row (+ slots).
template <int LANES, int VPL, int OPT>
__device__ __forceinline__ void row_apply_store(RowRegs<LANES, VPL>& r, const float4* g, const TableDev& T, int64_t row, int dim,
                                                int gl, const OptDev& o) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c = (gl + LANES * v) * 4;
        if (c >= dim) continue;
        float4 W = r.w[v], G = g[v];
        if (OPT == OPT_SGD) {
            W.x = __fsub_rn(W.x, __fmul_rn(o.lr, G.x));
            W.y = __fsub_rn(W.y, __fmul_rn(o.lr, G.y));
            W.z = __fsub_rn(W.z, __fmul_rn(o.lr, G.z));
            W.w = __fsub_rn(W.w, __fmul_rn(o.lr, G.w));
        } else if (OPT == OPT_ADAGRAD) {
            float4 A = r.s1[v];
            adagrad_elem(W.x, A.x, G.x, o.lr);
            adagrad_elem(W.y, A.y, G.y, o.lr);
            adagrad_elem(W.z, A.z, G.z, o.lr);
            adagrad_elem(W.w, A.w, G.w, o.lr);
            st4(T.s1 + row * dim + c, A);
        } else {
            float4 M = r.s1[v], V = r.s2[v];
            adam_touch_elem(W.x, M.x, V.x, G.x, o);
            adam_touch_elem(W.y, M.y, V.y, G.y, o);
            adam_touch_elem(W.z, M.z, V.z, G.z, o);
            adam_touch_elem(W.w, M.w, V.w, G.w, o);
            st4(T.s1 + row * dim + c, M);
            st4(T.s2 + row * dim + c, V);
        }
        st4(T.w + row * dim + c, W);
    }
    if (OptTraits<OPT>::replay && gl == 0) T.last[row] = o.step;
}

// one row of the triplet after the forward/backward: in place when it is the batch's only occurrence, else a slot
template <int LANES, int VPL, int OPT>
__device__ __forceinline__ void emit_row(RowRegs<LANES, VPL>& r, const float4* g, const TableDev& T, unsigned long long* meta,
                                         int32_t row, unsigned long long m, uint32_t rank, uint32_t t, uint32_t role, int dim,
                                         int gl, const OptDev& o, float* dup_grad, uint32_t* dup_t) {
    const uint32_t cnt = (uint32_t)m;
    if (cnt == 1u) {
        row_apply_store<LANES, VPL, OPT>(r, g, T, row, dim, gl, o);
        if (gl == 0) meta[row] = 0ULL;
    } else {
        const uint32_t slot = (uint32_t)(m >> 32) + rank;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            int c = (gl + LANES * v) * 4;
            if (c < dim) st4(dup_grad + (int64_t)slot * dim + c, g[v]);
        }
        if (gl == 0) dup_t[slot] = (t << 2) | role;  // unique, ordered key of the occurrence
    }
}

// block-level reduction of the per-thread double loss; result to block_loss[blockIdx.x]
__device__ __forceinline__ void block_loss_store(double v, double* block_loss) {
    __shared__ double sm_loss_[8];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) sm_loss_[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sm_loss_[k];
        block_loss[blockIdx.x] = t;
    }
}

// Sum of n_parts partial gradient vectors for the 32 consecutive elements k0 .. k0 + 31 by a 256-thread block (dense variables whose
// gradient arrives as per-CTA partials): warp w adds parts [w * per, (w + 1) * per) for element k0 + lane, in part order, then the
// eight warp sums are added in warp order -- a fixed order (deterministic), but eight chains of n_parts / 8 dependent adds over
// coalesced 128-byte loads instead of one chain of n_parts (up to 444 per element: 22 us for a 4160-element vector on 17 blocks).
// Returns the total in warp 0 (lane = element); contains two block barriers.
__device__ __forceinline__ float block_sum_parts(const float* __restrict__ parts, int n_parts, int64_t stride, int k, bool valid) {
    __shared__ float sm_parts_[8 * 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int per = (n_parts + 7) / 8;
    const int p0 = w * per, p1 = min(n_parts, p0 + per);
    float g = 0.f;
    if (valid) {
#pragma unroll 8
        for (int p = p0; p < p1; ++p) g += parts[(int64_t)p * stride + k];
    }
    sm_parts_[w * 32 + lane] = g;
    __syncthreads();
    float t = 0.f;
    if (w == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) t += sm_parts_[q * 32 + lane];
    }
    __syncthreads();
    return t;
}

// Multi-GPU step (train_sharded.cu): where the gradient of an ITEM row goes.  The item table is row-sharded (owner = item % n_ranks,
// local row = item / n_ranks); every owner holds a DIRECT-MAPPED inbox with one gradient slot per (source rank, local row):
// grad[owner] is [n_ranks][rows_cap][dim], stamp[owner] is [n_ranks][rows_cap] and holds the step whose gradient the slot carries.
// A source rank therefore sends at most ONE gradient per item row and step (duplicates are summed locally first), needs no slot
// reservation, and the inbox can neither overflow nor has to be cleared.  All pointers are valid on THIS device (CUDA IPC).
struct ShardSend {
    int n_ranks, rank;
    int64_t rows_cap;
    uint32_t stamp;                        // = opt.step of the current step (>= 1; slots are created zeroed)
    float* grad[CRB_MAX_RANKS];
    uint32_t* stamps[CRB_MAX_RANKS];
};

template <int LANES, int VPL>
__device__ __forceinline__ void shard_send(const ShardSend& sh, int32_t item, const float4* g, int dim, int gl) {
    const int owner = item % sh.n_ranks;
    const int64_t e = (int64_t)sh.rank * sh.rows_cap + item / sh.n_ranks;
    float* dst = sh.grad[owner] + e * dim;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c = (gl + LANES * v) * 4;
        if (c < dim) st4(dst + c, g[v]);
    }
    if (gl == 0) sh.stamps[owner][e] = sh.stamp;
}

// Arguments of the duplicate-row kernels (shared by all row-sparse steps)
struct DupArgs {
    TableDev tab[2];
    unsigned long long* meta[2];
    int dim;
    OptDev opt;
    const crb_dup_row* dup_rows;
    const crb_work* work;
    const unsigned int* multi;   // indices of multi-chunk duplicate rows
    const float* dup_grad;
    const uint32_t* dup_t;
    float* partial;
    const crb_step_ctr* ctr;
    // Optional indirection (multi-GPU inbox): slot q's gradient is row dup_src[q] of src_grad instead of row q of dup_grad, so
    // gradients that already sit in a buffer are summed from where they are instead of being copied into slots first.
    const uint32_t* dup_src = nullptr;
    const float* src_grad = nullptr;
    ShardSend send;   // SHARD kernels only: rows of table 1 are GLOBAL item ids whose summed gradient is sent to the owner
};

int crb_launch_dup_pipeline(crb_handle* h, const DupArgs& a, int opt_kind, cudaStream_t s);
bool crb_dup_tail_enabled(int64_t batch);
int crb_launch_dup_tail(crb_handle* h, const DupArgs& a, int opt_kind, double* loss_out_dev, cudaStream_t s);   // K4 + K5 + counter reset in one launch
int crb_launch_assign(crb_handle* h, int64_t batch, int n_roles, const int32_t* const* idx, const int* role_table,
                      cudaStream_t s, const unsigned int* n_dev = nullptr, bool every_row = false);
int crb_count_rows(crb_handle* h, int64_t batch, int n_roles, const int32_t* const* idx, const int* role_table, cudaStream_t s,
                   const unsigned int* n_dev = nullptr);
int crb_opt_to_dev(crb_handle* h, const crb_opt* opt, OptDev* out, int* opt_kind, cudaStream_t s);
int crb_table_check(const crb_table* T, int opt_kind, const char* name);
int crb_launch_loss_final(crb_handle* h, double* loss_out_dev, cudaStream_t s);
TableDev crb_to_dev(const crb_table* T);
int crb_zero_step_counters(crb_handle* h, cudaStream_t s);
int crb_launch_dense_apply_strided(crb_handle* h, float* w, float* s1, float* s2, const float* parts, int n_parts, int n, int part_stride,
                                   int opt_kind, const OptDev& od, cudaStream_t s);
int crb_launch_dense_apply(crb_handle* h, float* w, float* s1, float* s2, const float* parts, int n_parts, int n, int opt_kind, const OptDev& od,
                           cudaStream_t s);   // TF dense apply of a small dense variable from n_parts partial gradients summed in order
