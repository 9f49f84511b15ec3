// Fused training steps of the dot-product family (reference: model/ranking/BPR.py:31-44, GMF.py:37-49, MF per
// SURVEY F6).  One step = what `sess.run([self.train, self.loss], feed)` does in the reference:
//   gather rows -> interaction -> loss -> backward -> de-duplicated row-sparse optimizer apply.
// Kernels per step (all on one stream, no host synchronisation):
//   K1 count_rows / sample_kernel<.,true>  one 32-bit atomic per row occurrence -> multiplicity + occurrence rank
//   K2 assign_kernel                        duplicate rows get a slot range, a work list is built
//   K3 *_step_kernel                        the fused gather/forward/backward; rows of multiplicity 1 are updated
//                                           in place (read once, written once); duplicates store their gradient
//   K4 dup_reduce_kernel / dup_final_kernel slot sums (deterministic order for <= 32 occurrences) + one apply
//   K5 loss_final_kernel                    fixed-order sum of the per-block loss partials
#include <cstddef>

#include "rowopt.cuh"

// ------------------------------------------------------------------------------------------------ K1 / K2
struct RoleArgs {
    const int32_t* idx[3];
    uint32_t* rank[3];
    unsigned long long* meta[3];  // meta table of each role
    int table[3];
    int n_roles;
    const unsigned int* n_dev;   // optional device-side element count (inbox): effective batch = min(batch, *n_dev)
};

__global__ void __launch_bounds__(256) count_rows_kernel(RoleArgs a, int64_t batch) {
    if (a.n_dev && (int64_t)*a.n_dev < batch) batch = (int64_t)*a.n_dev;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < batch; t += stride) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
            if (r < a.n_roles) {
                const int32_t row = a.idx[r][t];   // negative = hole (multi-GPU inbox): never counted, never ranked 1
                a.rank[r][t] = row < 0 ? 0xFFFFFFFFu : atomicAdd(reinterpret_cast<unsigned int*>(a.meta[r] + row), 1u);
            }
    }
}

__global__ void __launch_bounds__(256) assign_kernel(RoleArgs a, int64_t batch, crb_step_ctr* ctr, crb_dup_row* dup_rows,
                                                     crb_work* work, unsigned int* multi, uint32_t alloc_rank) {
    // One occurrence of every duplicated row has rank 1: it allocates the row's slot range, work items and partial rows.
    // (alloc_rank == 0 instead lets the FIRST occurrence allocate, i.e. every row gets a slot range -- the multi-GPU inbox.)
    // The five global counters live in one 32-byte sector, so every atomic on them serialises in one L2 slice: they are
    // bumped once per BLOCK and round (a thread first adds up what its up to three roles need, then warp scan -> scan of the 8 warp
    // totals -> one atomic per counter), not once per row, warp or role: at the batch sizes the reference ships the kernel is one round
    // of 24 blocks and its duration is the atomics' round trip -- one now, three when the roles took turns.
    if (a.n_dev && (int64_t)*a.n_dev < batch) batch = (int64_t)*a.n_dev;
    __shared__ uint32_t s_tot[8][5];    // per-warp totals, then per-warp exclusive bases (global)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (batch + stride - 1) / stride;
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t t = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool mine[3];
        int32_t row[3];
        uint32_t c[3], nch[3];
        unsigned int* m[3];
        bool any = false;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            mine[r] = r < a.n_roles && t < batch && a.rank[r][t] == alloc_rank;
            row[r] = 0; c[r] = 0; nch[r] = 0; m[r] = nullptr;
            any |= mine[r];
        }
        if (!__syncthreads_or(any)) continue;   // block-uniform
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (mine[r]) {
                row[r] = a.idx[r][t];
                m[r] = reinterpret_cast<unsigned int*>(a.meta[r] + row[r]);
                c[r] = m[r][0];
                nch[r] = (c[r] + CRB_DUP_CHUNK - 1) / CRB_DUP_CHUNK;
            }
        }
        // what this thread needs of each counter (slots, rows, work items, partial rows, multi-chunk rows), all roles together
        uint32_t tc = 0, t1 = 0, tn = 0, tp = 0, tm = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            tc += c[r]; t1 += mine[r] ? 1u : 0u; tn += nch[r]; tp += nch[r] > 1 ? nch[r] : 0u; tm += nch[r] > 1 ? 1u : 0u;
        }
        // inclusive warp scans
        uint32_t sc = tc, s1 = t1, sn = tn, sp = tp, sm = tm;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t uc = __shfl_up_sync(0xffffffffu, sc, o), u1 = __shfl_up_sync(0xffffffffu, s1, o);
            const uint32_t un = __shfl_up_sync(0xffffffffu, sn, o), up = __shfl_up_sync(0xffffffffu, sp, o);
            const uint32_t um = __shfl_up_sync(0xffffffffu, sm, o);
            if (lane >= o) { sc += uc; s1 += u1; sn += un; sp += up; sm += um; }
        }
        if (lane == 31) { s_tot[wid][0] = sc; s_tot[wid][1] = s1; s_tot[wid][2] = sn; s_tot[wid][3] = sp; s_tot[wid][4] = sm; }
        __syncthreads();
        if (wid == 0 && lane < 5) {
            // lane q owns counter q: exclusive prefix over the 8 warps, one atomic for the block
            uint32_t pre[8], tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { pre[w] = tot; tot += s_tot[w][lane]; }
            unsigned int* cp = lane == 0 ? &ctr->dup_slots : lane == 1 ? &ctr->dup_rows : lane == 2 ? &ctr->work_items
                               : lane == 3 ? &ctr->partial_slots : &ctr->multi_rows;
            const uint32_t gb = tot ? atomicAdd(cp, tot) : 0u;
#pragma unroll
            for (int w = 0; w < 8; ++w) s_tot[w][lane] = gb + pre[w];
        }
        __syncthreads();
        // this thread's first slot / row / work item / partial row / multi-chunk entry; its roles take theirs in role order
        uint32_t base = s_tot[wid][0] + sc - tc, k = s_tot[wid][1] + s1 - t1, wb = s_tot[wid][2] + sn - tn, pb = s_tot[wid][3] + sp - tp,
                 mb = s_tot[wid][4] + sm - tm;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (mine[r]) {
                const uint32_t pch = nch[r] > 1 ? nch[r] : 0u;
                m[r][1] = base;
                if (nch[r] > 1) multi[mb++] = k;
                crb_dup_row d;
                d.row = row[r]; d.table = a.table[r]; d.base = base; d.cnt = c[r]; d.wbase = wb; d.nchunk = nch[r]; d.pbase = pb; d.pad = 0;
                dup_rows[k] = d;
                for (uint32_t q = 0; q < nch[r]; ++q) { crb_work w; w.dup = k; w.chunk = q; work[wb + q] = w; }
                base += c[r]; ++k; wb += nch[r]; pb += pch;
            }
        }
        __syncthreads();   // s_tot is reused by the next round
    }
}

static int flat_grid(crb_handle* h, int64_t n, int per_block) {
    int64_t b = (n + per_block - 1) / per_block;
    int64_t cap = (int64_t)h->sm_count * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

int crb_count_rows(crb_handle* h, int64_t batch, int n_roles, const int32_t* const* idx, const int* role_table, cudaStream_t s,
                   const unsigned int* n_dev) {
    RoleArgs a;
    a.n_roles = n_roles;
    a.n_dev = n_dev;
    for (int r = 0; r < 3; ++r) {
        a.idx[r] = r < n_roles ? idx[r] : nullptr;
        a.rank[r] = h->rank[r];
        a.table[r] = r < n_roles ? role_table[r] : 0;
        a.meta[r] = h->meta[a.table[r]];
    }
    count_rows_kernel<<<flat_grid(h, batch, 256), 256, 0, s>>>(a, batch);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

int crb_launch_assign(crb_handle* h, int64_t batch, int n_roles, const int32_t* const* idx, const int* role_table, cudaStream_t s,
                      const unsigned int* n_dev, bool every_row) {
    RoleArgs a;
    a.n_roles = n_roles;
    a.n_dev = n_dev;
    for (int r = 0; r < 3; ++r) {
        a.idx[r] = r < n_roles ? idx[r] : nullptr;
        a.rank[r] = h->rank[r];
        a.table[r] = r < n_roles ? role_table[r] : 0;
        a.meta[r] = h->meta[a.table[r]];
    }
    h->ctr_zeroed = 0;
    assign_kernel<<<flat_grid(h, batch, 256), 256, 0, s>>>(a, batch, h->ctr, h->dup_rows, h->work, h->multi, every_row ? 0u : 1u);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

#include "dup.cuh"


template <int LANES, int VPL>
static int launch_dup_t(crb_handle* h, const DupArgs& a, int opt_kind, cudaStream_t s) {
    const int grid = h->sm_count * 4;
#define CRB_DUP_CASE(O)                                                \
    case O:                                                            \
        dup_reduce_kernel<LANES, VPL, O><<<grid, 256, 0, s>>>(a);      \
        dup_final_kernel<LANES, VPL, O><<<h->sm_count, 256, 0, s>>>(a); \
        break;
    switch (opt_kind) {
        CRB_DUP_CASE(OPT_SGD)
        CRB_DUP_CASE(OPT_ADAGRAD)
        CRB_DUP_CASE(OPT_ADAM_LAZY)
        CRB_DUP_CASE(OPT_ADAM_TF1)
    }
#undef CRB_DUP_CASE
    h->launches += 2;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

#define CRB_DIM_DISPATCH(dim, FN, ...)                                   \
    ((dim) <= 32 ? FN<8, 1>(__VA_ARGS__)                                 \
     : (dim) <= 64 ? FN<16, 1>(__VA_ARGS__)                              \
     : (dim) <= 128 ? FN<32, 1>(__VA_ARGS__)                             \
     : (dim) <= 256 ? FN<32, 2>(__VA_ARGS__)                             \
                    : FN<32, 4>(__VA_ARGS__))

int crb_launch_dup_pipeline(crb_handle* h, const DupArgs& a, int opt_kind, cudaStream_t s) {
    return CRB_DIM_DISPATCH(a.dim, launch_dup_t, h, a, opt_kind, s);
}

template <int LANES, int VPL>
static int launch_dup_tail_t(crb_handle* h, const DupArgs& a, const DupTail& t, int opt_kind, cudaStream_t s) {
    const int grid = h->sm_count * 4;   // as dup_reduce_kernel: one lane group per duplicate row at the shapes this path serves
    // Programmatic dependent launch behind the step kernel (which signals griddepcontrol.launch_dependents when it starts): the blocks
    // of this kernel become resident as the step kernel's blocks retire and read the counters / work list / row descriptors -- K2's
    // output -- while the step kernel's tail is still running; `griddepcontrol.wait` orders everything else.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    switch (opt_kind) {
        case OPT_SGD: CRB_CUDA(cudaLaunchKernelEx(&cfg, dup_tail_kernel<LANES, VPL, OPT_SGD>, a, t)); break;
        case OPT_ADAGRAD: CRB_CUDA(cudaLaunchKernelEx(&cfg, dup_tail_kernel<LANES, VPL, OPT_ADAGRAD>, a, t)); break;
        case OPT_ADAM_LAZY: CRB_CUDA(cudaLaunchKernelEx(&cfg, dup_tail_kernel<LANES, VPL, OPT_ADAM_LAZY>, a, t)); break;
        case OPT_ADAM_TF1: CRB_CUDA(cudaLaunchKernelEx(&cfg, dup_tail_kernel<LANES, VPL, OPT_ADAM_TF1>, a, t)); break;
    }
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    h->ctr_zeroed = 1;
    return CRB_OK;
}

// batches up to 2^16 rows (see dup_tail_kernel); CRB_DUP_TAIL=0 keeps the three separate launches (A/B and the bit-identity test)
bool crb_dup_tail_enabled(int64_t batch) {
    const char* e = getenv("CRB_DUP_TAIL");   // read per call: the test flips it inside one process
    return !(e && atoi(e) == 0) && batch <= 65536;
}

int crb_launch_dup_tail(crb_handle* h, const DupArgs& a, int opt_kind, double* loss_out_dev, cudaStream_t s) {
    DupTail t;
    t.ctr = h->ctr; t.block_loss = h->block_loss; t.n_block_loss = h->step_grid; t.loss_out = loss_out_dev;
    return CRB_DIM_DISPATCH(a.dim, launch_dup_tail_t, h, a, t, opt_kind, s);
}

// ------------------------------------------------------------------------------------------------ K5
__global__ void loss_final_kernel(const double* __restrict__ block_loss, int n, double* out) {
    // one warp, fixed order: lane l sums entries l, l+32, ...; then a fixed shuffle tree
    double v = 0.0;
    for (int k = threadIdx.x; k < n; k += 32) v += block_loss[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) *out = v;
}

int crb_launch_loss_final(crb_handle* h, double* loss_out_dev, cudaStream_t s) {
    loss_final_kernel<<<1, 32, 0, s>>>(h->block_loss, h->step_grid, loss_out_dev);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

// block-level reduction of the per-thread double loss; result to block_loss[blockIdx.x]
__device__ __forceinline__ void block_loss_store_(double v, double* block_loss) {
    __shared__ double sm[8];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) sm[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sm[k];
        block_loss[blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------------ optimizer plumbing
int crb_opt_to_dev(crb_handle* h, const crb_opt* opt, OptDev* o, int* opt_kind, cudaStream_t s) {
    CRB_CHECK_ARG(opt, "opt is null");
    h->evq_valid = 0;   // every entry point that writes a table comes through here: the cached bf16 item table is stale from now on
    CRB_CHECK_ARG(opt->step >= 1 && opt->step < 0x7fffffffLL, "opt.step must be >= 1");
    o->lr = (float)opt->lr;
    o->b1 = (float)opt->beta1;
    o->b2 = (float)opt->beta2;
    o->eps = (float)opt->eps;
    o->omb1 = 1.f - o->b1;
    o->omb2 = 1.f - o->b2;
    o->step = (int32_t)opt->step;
    o->lr_t = 0.f;
    o->lrt = h->lrt;
    o->step_base = nullptr;
    switch (opt->kind) {
        case CRB_OPT_SGD: *opt_kind = OPT_SGD; break;
        case CRB_OPT_ADAGRAD: *opt_kind = OPT_ADAGRAD; break;
        case CRB_OPT_ADAM: {
            *opt_kind = opt->adam_mode == CRB_ADAM_LAZY ? OPT_ADAM_LAZY : OPT_ADAM_TF1;
            const double t = (double)opt->step;
            o->lr_t = (float)(opt->lr * sqrt(1.0 - pow(opt->beta2, t)) / (1.0 - pow(opt->beta1, t)));
            int rc = crb_lrt_prepare(h, opt, s);
            if (rc) return rc;
            break;
        }
        default:
            crb_set_error("unknown optimizer kind %d", opt->kind);
            return CRB_ERR_ARG;
    }
    return CRB_OK;
}

int crb_table_check(const crb_table* T, int opt_kind, const char* name) {
    if (!T || !T->w || T->rows <= 0 || T->dim <= 0 || (T->dim % 4) != 0 || T->dim > 512) {
        crb_set_error("table %s: need w != NULL, rows > 0, dim %% 4 == 0, dim <= 512", name);
        return CRB_ERR_ARG;
    }
    if (((uintptr_t)T->w & 15) != 0) { crb_set_error("table %s: w must be 16-byte aligned", name); return CRB_ERR_ARG; }
    if (opt_kind != OPT_SGD && !T->s1) { crb_set_error("table %s: optimizer slot s1 is NULL", name); return CRB_ERR_ARG; }
    if ((opt_kind == OPT_ADAM_LAZY || opt_kind == OPT_ADAM_TF1) && !T->s2) { crb_set_error("table %s: optimizer slot s2 is NULL", name); return CRB_ERR_ARG; }
    if (opt_kind == OPT_ADAM_TF1 && !T->last) { crb_set_error("table %s: CRB_ADAM_TF1 needs the `last` step array", name); return CRB_ERR_ARG; }
    return CRB_OK;
}

TableDev crb_to_dev(const crb_table* T) {
    TableDev d;
    d.w = T->w; d.s1 = T->s1; d.s2 = T->s2; d.last = T->last;
    return d;
}

// ------------------------------------------------------------------------------------------------ K3: BPR
#include "bpr_args.cuh"

// CTAs per SM: the Adam variants need ~80 registers (3 CTAs).  SGD moves 3x fewer bytes per triplet and is bound by the bytes in
// flight: it fits 64 registers without spilling and takes a fourth resident CTA (K3 1.03 -> 0.76 ms).  Adagrad spills at 64 registers
// and gets slower (1.23 -> 1.32 ms), so it stays at three.
#ifndef CRB_SGD_CTAS
#define CRB_SGD_CTAS 4
#endif
template <int OPT> struct BprOcc { static constexpr int ctas = OPT == OPT_SGD ? CRB_SGD_CTAS : 3; };

template <int LANES, int VPL, int OPT>
__global__ void __launch_bounds__(256, BprOcc<OPT>::ctas) bpr_step_kernel(BprArgs a) {
    asm volatile("griddepcontrol.launch_dependents;");   // dup_tail_kernel (small batches) may start its prologue; no-op otherwise
    opt_resolve(a.opt);
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LANES;
    const int sub = lane / LANES;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double loss_acc = 0.0;
    // Software pipeline over the index side: the triplet's ids are fetched two iterations ahead and its multiplicity words / `last`
    // steps one iteration ahead, so that when an iteration starts everything that decides WHICH rows and slots to load is already
    // in registers and the weights and optimizer slots go out in one round trip (instead of ids -> meta -> slots, three dependent
    // round trips: the SGD / Adagrad variants were latency bound on that chain).  Out-of-range iterations clamp (loads only).
    // The Adam variants sit at the 85-register limit of three CTAs per SM and are bandwidth bound already: there the extra live
    // registers spill and the pipeline costs 20 % (measured), so they keep the plain order.
    constexpr bool PIPE = !OptTraits<OPT>::has_s2;
    const int64_t stride = n_warps * GPW;
    const int64_t base0 = warp * GPW;
    auto clampt = [&](int64_t b_) { const int64_t t_ = b_ + sub; return t_ < a.batch ? t_ : a.batch - 1; };
    int32_t nu = 0, ni = 0, nj = 0, fu = 0, fi = 0, fj = 0;
    unsigned long long nmu = 0, nmi = 0, nmj = 0;
    int32_t nlu = 0, nli = 0, nlj = 0;
    if (PIPE) {
        { const int64_t t0 = clampt(base0); nu = a.u[t0]; ni = a.i[t0]; nj = a.j[t0]; }
        { const int64_t t1 = clampt(base0 + stride); fu = a.u[t1]; fi = a.i[t1]; fj = a.j[t1]; }
        nmu = a.metaU[nu]; nmi = a.metaI[ni]; nmj = a.metaI[nj];
    }
    for (int64_t base = base0; base < a.batch; base += stride) {
        const int64_t t = base + sub;
        const bool active = t < a.batch;
        if (!PIPE) {
            const int64_t tt = clampt(base);
            nu = a.u[tt]; ni = a.i[tt]; nj = a.j[tt];
            nmu = a.metaU[nu]; nmi = a.metaI[ni]; nmj = a.metaI[nj];
            if (OptTraits<OPT>::replay) { nlu = a.P.last[nu]; nli = a.Q.last[ni]; nlj = a.Q.last[nj]; }
        }
        const int32_t u = nu, i = ni, j = nj;
        const unsigned long long mu = nmu, mi = nmi, mj = nmj;
        RowRegs<LANES, VPL> ru, ri, rj;
        ru.last = nlu; ri.last = nli; rj.last = nlj;
        row_load_w<LANES, VPL>(ru, a.P, u, a.dim, gl);
        row_load_w<LANES, VPL>(ri, a.Q, i, a.dim, gl);
        row_load_w<LANES, VPL>(rj, a.Q, j, a.dim, gl);
        // optimizer slots are needed here only for rows this group will update in place, or -- CRB_ADAM_TF1 -- rows with
        // missed decay steps, because the forward must see the replayed (dense-equivalent) value.
        const bool su = (uint32_t)mu == 1u || replay_pending<OPT>(ru.last, a.opt);
        const bool si = (uint32_t)mi == 1u || replay_pending<OPT>(ri.last, a.opt);
        const bool sj = (uint32_t)mj == 1u || replay_pending<OPT>(rj.last, a.opt);
        if (OptTraits<OPT>::has_s1) {
            if (su) row_load_state<LANES, VPL, OPT>(ru, a.P, u, a.dim, gl);
            if (si) row_load_state<LANES, VPL, OPT>(ri, a.Q, i, a.dim, gl);
            if (sj) row_load_state<LANES, VPL, OPT>(rj, a.Q, j, a.dim, gl);
        }
        // rotate the index pipeline: next iteration's multiplicity words / last steps, and the ids after that
        if (PIPE) {
            nu = fu; ni = fi; nj = fj;
            nmu = a.metaU[nu]; nmi = a.metaI[ni]; nmj = a.metaI[nj];
            const int64_t t2 = clampt(base + 2 * stride);
            fu = a.u[t2]; fi = a.i[t2]; fj = a.j[t2];
        }
        if (replay_pending<OPT>(ru.last, a.opt)) row_replay<LANES, VPL, OPT>(ru, a.opt, a.opt.step);
        if (replay_pending<OPT>(ri.last, a.opt)) row_replay<LANES, VPL, OPT>(ri, a.opt, a.opt.step);
        if (replay_pending<OPT>(rj.last, a.opt)) row_replay<LANES, VPL, OPT>(rj, a.opt, a.opt.step);
        // forward: x = p_u.(q_i - q_j)  (BPR.py:39-41);  l2 = |p_u|^2 + |q_i|^2 + |q_j|^2 (BPR.py:42-43)
        float x = 0.f, sq = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const float4 p = ru.w[v], qi = ri.w[v], qj = rj.w[v];
            const float4 dq = make_float4(qi.x - qj.x, qi.y - qj.y, qi.z - qj.z, qi.w - qj.w);
            x += dot4(p, dq);
            sq += dot4(p, p) + dot4(qi, qi) + dot4(qj, qj);
        }
        x = group_sum<LANES>(x);
        sq = group_sum<LANES>(sq);
        // d/dx softplus(-x) = sigmoid(x) - 1 = -sigmoid(-x)
        const float g = -sigmoid_f(-x);
        if (active && gl == 0) loss_acc += (double)(softplus_neg(x) + a.reg * 0.5f * sq);
        float4 gu[VPL], gi[VPL], gj[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const float4 p = ru.w[v], qi = ri.w[v], qj = rj.w[v];
            gu[v] = make_float4(fmaf(g, qi.x - qj.x, a.reg * p.x), fmaf(g, qi.y - qj.y, a.reg * p.y),
                                fmaf(g, qi.z - qj.z, a.reg * p.z), fmaf(g, qi.w - qj.w, a.reg * p.w));
            gi[v] = make_float4(fmaf(g, p.x, a.reg * qi.x), fmaf(g, p.y, a.reg * qi.y), fmaf(g, p.z, a.reg * qi.z),
                                fmaf(g, p.w, a.reg * qi.w));
            gj[v] = make_float4(fmaf(-g, p.x, a.reg * qj.x), fmaf(-g, p.y, a.reg * qj.y), fmaf(-g, p.z, a.reg * qj.z),
                                fmaf(-g, p.w, a.reg * qj.w));
        }
        if (active) {
            emit_row<LANES, VPL, OPT>(ru, gu, a.P, a.metaU, u, mu, a.rk[0][t], (uint32_t)t, 0u, a.dim, gl, a.opt, a.dup_grad, a.dup_t);
            emit_row<LANES, VPL, OPT>(ri, gi, a.Q, a.metaI, i, mi, a.rk[1][t], (uint32_t)t, 1u, a.dim, gl, a.opt, a.dup_grad, a.dup_t);
            emit_row<LANES, VPL, OPT>(rj, gj, a.Q, a.metaI, j, mj, a.rk[2][t], (uint32_t)t, 2u, a.dim, gl, a.opt, a.dup_grad, a.dup_t);
        }
    }
    block_loss_store(loss_acc, a.block_loss);
}

// Persistent launch: exactly one wave of resident CTAs (SM count x occupancy), each looping over the batch.
template <typename K>
static int persistent_grid(crb_handle* h, K kernel, int64_t work_groups, int groups_per_block) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 2; }
    int64_t grid = (int64_t)h->sm_count * occ;
    const int64_t need = (work_groups + groups_per_block - 1) / groups_per_block;
    if (grid > need) grid = need;
    if (grid > h->loss_blocks) grid = h->loss_blocks;
    if (grid < 1) grid = 1;
    h->step_grid = (int)grid;
    return (int)grid;
}

template <int LANES, int VPL>
static int launch_bpr_t(crb_handle* h, const BprArgs& a, int opt_kind, cudaStream_t s) {
    const int gpb = 256 / LANES;
#define CRB_BPR_CASE(O)                                                                                    \
    case O: {                                                                                              \
        const int grid = persistent_grid(h, bpr_step_kernel<LANES, VPL, O>, a.batch, gpb);                 \
        bpr_step_kernel<LANES, VPL, O><<<grid, 256, 0, s>>>(a);                                            \
        break;                                                                                             \
    }
    switch (opt_kind) {
        CRB_BPR_CASE(OPT_SGD)
        CRB_BPR_CASE(OPT_ADAGRAD)
        CRB_BPR_CASE(OPT_ADAM_LAZY)
        CRB_BPR_CASE(OPT_ADAM_TF1)
    }
#undef CRB_BPR_CASE
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

// Host feeds of the reference's batch sizes are range-checked before they are used as table rows (a feed_dict with a bad id makes
// TF's gather raise; here it would be an out-of-bounds access).  Device-resident feeds and very large host feeds are trusted: they
// come from the library's own sampler / the caller's pipeline, and a scan of 3 x 2^20 ids would cost a third of the step.
#define CRB_FEED_CHECK_MAX (1 << 16)
static int check_feed(const int32_t* src, int64_t n, int64_t rows, const char* name) {
    if (n > CRB_FEED_CHECK_MAX) return CRB_OK;
    for (int64_t k = 0; k < n; ++k)
        if (src[k] < 0 || src[k] >= rows) {
            crb_set_error("feed %s[%lld] = %d is outside [0, %lld)", name, (long long)k, src[k], (long long)rows);
            return CRB_ERR_ARG;
        }
    return CRB_OK;
}

// stage a possibly-host int32 feed into the handle workspace; returns the device pointer to use
static int stage_i32(crb_handle* h, const int32_t* src, int slot, int64_t n, const int32_t** out, cudaStream_t s, int64_t rows = -1,
                     const char* name = "") {
    if (crb_is_device_ptr(src)) { *out = src; return CRB_OK; }
    if (rows >= 0) { int rc = check_feed(src, n, rows, name); if (rc) return rc; }
    CRB_CUDA(cudaMemcpyAsync(h->idx[slot], src, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    *out = h->idx[slot];
    return CRB_OK;
}

static int finish_loss(crb_handle* h, double* loss_out, int64_t n, cudaStream_t s) {
    // device buffers were written in place by loss_final_kernel; host buffers get one copy + sync
    if (!loss_out || crb_is_device_ptr(loss_out)) return CRB_OK;
    CRB_CUDA(cudaMemcpyAsync(loss_out, h->loss_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    CRB_CUDA(cudaStreamSynchronize(s));
    return CRB_OK;
}

// K2 (and K1's counting half when the sampler did not do it) of one BPR step: everything that depends only on the indices
static int bpr_step_prepare(crb_handle* h, const int32_t* u, const int32_t* i, const int32_t* j, int64_t batch, bool counted,
                            cudaStream_t s) {
    const int32_t* idx[3] = {u, i, j};
    const int role_table[3] = {0, 1, 1};
    int rc;
    if (!counted) {
        rc = crb_count_rows(h, batch, 3, idx, role_table, s);
        if (rc) return rc;
    }
    return crb_launch_assign(h, batch, 3, idx, role_table, s);
}

// K3, K4, K5 of one BPR step (reads the state bpr_step_prepare left in the handle's current step set)
static int bpr_step_compute(crb_handle* h, const crb_table* P, const crb_table* Q, const OptDev& od, int opt_kind,
                            const int32_t* u, const int32_t* i, const int32_t* j, int64_t batch, float reg, double* loss_dev,
                            cudaStream_t s) {
    int rc;
    BprArgs a;
    a.P = crb_to_dev(P); a.Q = crb_to_dev(Q);
    a.metaU = h->meta[0]; a.metaI = h->meta[1];
    a.u = u; a.i = i; a.j = j;
    a.rk[0] = h->rank[0]; a.rk[1] = h->rank[1]; a.rk[2] = h->rank[2];
    a.batch = batch; a.dim = P->dim; a.reg = reg; a.opt = od;
    a.dup_grad = h->dup_grad; a.dup_t = h->dup_t; a.block_loss = h->block_loss;
    if ((rc = crb_prof_begin(h, s))) return rc;
    rc = crb_bpr_ring_enabled(a.dim) ? crb_launch_bpr_ring(h, a, opt_kind, s) : CRB_DIM_DISPATCH(a.dim, launch_bpr_t, h, a, opt_kind, s);
    if (rc) return rc;
    if ((rc = crb_prof_end(h, s))) return rc;
    DupArgs d;
    d.tab[0] = a.P; d.tab[1] = a.Q;
    d.meta[0] = h->meta[0]; d.meta[1] = h->meta[1];
    d.dim = a.dim; d.opt = od;
    d.dup_rows = h->dup_rows; d.work = h->work; d.multi = h->multi;
    d.dup_grad = h->dup_grad; d.dup_t = h->dup_t; d.partial = h->partial; d.ctr = h->ctr;
    if ((rc = crb_prof_begin(h, s, 2))) return rc;
    const bool tail = crb_dup_tail_enabled(batch);
    rc = tail ? crb_launch_dup_tail(h, d, opt_kind, loss_dev, s) : crb_launch_dup_pipeline(h, d, opt_kind, s);
    if (rc) return rc;
    if ((rc = crb_prof_end(h, s, 2))) return rc;
    return tail ? CRB_OK : crb_launch_loss_final(h, loss_dev, s);
}

// the part of one BPR step after the indices are on the device and (optionally) already counted
static int bpr_step_device(crb_handle* h, const crb_table* P, const crb_table* Q, const OptDev& od, int opt_kind,
                           const int32_t* u, const int32_t* i, const int32_t* j, int64_t batch, float reg, bool counted,
                           double* loss_dev, cudaStream_t s) {
    int rc = bpr_step_prepare(h, u, i, j, batch, counted, s);
    if (rc) return rc;
    return bpr_step_compute(h, P, Q, od, opt_kind, u, i, j, batch, reg, loss_dev, s);
}

__global__ void merge_sampler_err_kernel(crb_step_ctr* into, crb_step_ctr* from) {
    into->sampler_err += from->sampler_err;
    from->sampler_err = 0;
}

int crb_zero_step_counters(crb_handle* h, cudaStream_t s) {
    // everything except sampler_err (sticky until reported); nothing to do when the last step on this copy ended in dup_tail_kernel
    if (h->ctr_zeroed) return CRB_OK;
    CRB_CUDA(cudaMemsetAsync(h->ctr, 0, offsetof(crb_step_ctr, sampler_err), s));
    return CRB_OK;
}

static int bpr_common_checks(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_opt* opt, OptDev* od, int* opt_kind,
                             int64_t batch, int64_t steps, cudaStream_t s) {
    CRB_CHECK_ARG(h, "null handle");
    int rc = crb_opt_to_dev(h, opt, od, opt_kind, s);
    if (rc) return rc;
    if ((rc = crb_table_check(P, *opt_kind, "P"))) return rc;
    if ((rc = crb_table_check(Q, *opt_kind, "Q"))) return rc;
    CRB_CHECK_ARG(P->dim == Q->dim, "P.dim != Q.dim");
    CRB_CHECK_ARG(batch > 0 && batch < 0x40000000LL, "batch must be in [1, 2^30)");
    CRB_CUDA(cudaSetDevice(h->device));
    if ((rc = crb_ws_reserve(h, batch, P->dim, steps, s))) return rc;
    if ((rc = crb_meta_reserve(h, 0, P->rows, s))) return rc;
    if ((rc = crb_meta_reserve(h, 1, Q->rows, s))) return rc;
    return CRB_OK;
}

extern "C" int crb_train_step_bpr(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_opt* opt, const int32_t* u,
                                  const int32_t* i, const int32_t* j, int64_t batch, float reg, double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    OptDev od;
    int opt_kind = 0;
    int rc = bpr_common_checks(h, P, Q, opt, &od, &opt_kind, batch, 1, s);
    if (rc) return rc;
    CRB_CHECK_ARG(u && i && j, "null index feed");
    const int32_t *du, *di, *dj;
    if ((rc = stage_i32(h, u, 0, batch, &du, s, P->rows, "u_idx"))) return rc;
    if ((rc = stage_i32(h, i, 1, batch, &di, s, Q->rows, "i_idx"))) return rc;
    if ((rc = stage_i32(h, j, 2, batch, &dj, s, Q->rows, "j_idx"))) return rc;
    if ((rc = crb_zero_step_counters(h, s))) return rc;
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    rc = bpr_step_device(h, P, Q, od, opt_kind, du, di, dj, batch, reg, false, ld, s);
    if (rc) return rc;
    return finish_loss(h, loss_out, 1, s);
}

// ------------------------------------------------------------------------------------------------ whole-epoch graph
// At the shapes the reference ships (ml-1m: 6144-row batches, tables of a few MB) a step is a handful of microsecond kernels and
// the epoch costs launch count x launch latency: 649 steps x 8 launches from the host.  For such shapes the single-stream epoch loop
// is captured ONCE into a CUDA graph and replayed with one launch per epoch.  What changes from epoch to epoch -- the sampler's
// permutation keys and epoch word, the optimizer's absolute step -- is read by the kernels from handle->dyn_dev (SamplerArgs::dyn,
// OptDev::step_base), which is rewritten before every replay; everything else (row offsets of the steps, batch sizes, loss slots,
// workspace pointers) is identical every epoch and baked into the nodes.  Same kernels, same order, same arguments: the tables are
// bit-identical to the step-by-step path (tests/test_gpu_train_bpr.py::test_epoch_graph_is_bit_identical).
void crb_sampler_perm_keys(uint64_t seed, uint32_t epoch, uint32_t keys[6]);

struct EpochGraphKey {
    crb_table P, Q;
    crb_opt opt;            // step zeroed
    uint64_t seed;
    int64_t first, batch, n_steps, rows, ws_generation, n_pos;
    int32_t neg_ratio;
    float reg;
    const void *pos_user, *seen_rowptr, *bloom, *stream;
};
static_assert(sizeof(EpochGraphKey) <= 256, "epoch graph key");

// MEASURED (B200, ml-1m shape: 649 steps of 6144 rows, d=64, Adam): graph replay 24.4 ms per epoch, ordinary launch loop 24.3 ms.
// The epoch is NOT bound by host launches: a step is six kernels whose cost is their chains of dependent L2 / DRAM round trips
// (sampler 19 us, assign 10, step 8, duplicate reduce up to 40 under ncu), and the ordinary path already overlaps step k+1's
// index kernels with step k's compute on a second stream, which the single-stream graph gives up.  So the graph is opt-in
// (CRB_EPOCH_GRAPH=1), kept because it is bit-identical and removes the host from the loop (one launch per epoch).
static bool epoch_graph_eligible(crb_handle* h, int64_t batch, int64_t n_steps) {
    const char* e = getenv("CRB_EPOCH_GRAPH");
    if (h->prof_on || !e || atoi(e) != 1) return false;
    return n_steps >= 8 && batch <= 32768;
}

// the epoch loop on ONE stream (no second copy of the step state); relative optimizer steps + device-side resolution when dyn != 0
static int bpr_epoch_single_stream(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_opt* opt, uint64_t seed, uint32_t epoch,
                                   int64_t first, int64_t batch, int64_t n_steps, int64_t rows, int32_t neg_ratio, float reg, bool dyn,
                                   double* loss_dev, cudaStream_t s) {
    crb_opt step_opt = *opt;
    OptDev od;
    int opt_kind = 0, rc;
    for (int64_t k = 0; k < n_steps; ++k) {
        const int64_t lo = first + k * batch;
        if (lo >= rows) { crb_set_error("step %lld starts past the end of the epoch", (long long)k); return CRB_ERR_ARG; }
        const int64_t b = (rows - lo) < batch ? (rows - lo) : batch;
        step_opt.step = dyn ? k + 1 : opt->step + k;
        if ((rc = crb_opt_to_dev(h, &step_opt, &od, &opt_kind, s))) return rc;
        if (dyn) { od.step_base = h->dyn_dev + 7; od.lr_t = 0.f; }
        if ((rc = crb_zero_step_counters(h, s))) return rc;
        if ((rc = crb_launch_sample_pairwise(h, seed, epoch, lo, b, neg_ratio, h->idx[0], h->idx[1], h->idx[2], nullptr, true, s))) return rc;
        if ((rc = bpr_step_prepare(h, h->idx[0], h->idx[1], h->idx[2], b, true, s))) return rc;
        if ((rc = bpr_step_compute(h, P, Q, od, opt_kind, h->idx[0], h->idx[1], h->idx[2], b, reg, loss_dev + k, s))) return rc;
    }
    return CRB_OK;
}

// returns CRB_OK with *done = true when the epoch ran from the graph; *done = false means "use the ordinary path"
static int bpr_epoch_graph(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_opt* opt, uint64_t seed, uint32_t epoch,
                           int64_t first, int64_t batch, int64_t n_steps, int64_t rows, int32_t neg_ratio, float reg, cudaStream_t s, bool* done) {
    *done = false;
    EpochGraphKey key;
    memset(&key, 0, sizeof(key));
    key.P = *P; key.Q = *Q; key.opt = *opt; key.opt.step = 0;
    key.seed = seed; key.first = first; key.batch = batch; key.n_steps = n_steps; key.rows = rows; key.ws_generation = h->ws_generation;
    key.n_pos = h->n_pos; key.neg_ratio = neg_ratio; key.reg = reg;
    key.pos_user = h->pos_user; key.seen_rowptr = h->seen_rowptr; key.bloom = h->bloom; key.stream = (const void*)s;
    const bool hit = h->epoch_graph && h->epoch_graph_key_len == (int)sizeof(key) && memcmp(h->epoch_graph_key, &key, sizeof(key)) == 0;
    if (!hit) {
        if (h->epoch_graph) { cudaGraphExecDestroy((cudaGraphExec_t)h->epoch_graph); h->epoch_graph = nullptr; }
        // everything that may allocate or synchronise happens before the capture (lr_t table, workspaces: done by the caller's checks)
        OptDev od;
        int opt_kind = 0;
        int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
        if (rc) return rc;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return CRB_OK; }
        h->dyn_mode = 1;
        const int64_t launches_before = h->launches;
        rc = bpr_epoch_single_stream(h, P, Q, opt, seed, epoch, first, batch, n_steps, rows, neg_ratio, reg, true, h->loss_dev, s);
        h->dyn_mode = 0;
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(s, &graph);
        h->launches = launches_before;
        if (rc || ce != cudaSuccess || !graph) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            return rc;   // an argument error is an error; a capture failure falls back to the ordinary path
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess || !exec) { cudaGetLastError(); return CRB_OK; }
        h->epoch_graph = exec;
        memcpy(h->epoch_graph_key, &key, sizeof(key));
        h->epoch_graph_key_len = (int)sizeof(key);
    }
    uint32_t dyn[8];
    crb_sampler_perm_keys(seed, epoch, dyn);
    dyn[6] = epoch;
    dyn[7] = (uint32_t)(opt->step - 1);
    CRB_CUDA(cudaMemcpyAsync(h->dyn_dev, dyn, sizeof(dyn), cudaMemcpyHostToDevice, s));   // pageable source: staged before the call returns
    CRB_CUDA(cudaGraphLaunch((cudaGraphExec_t)h->epoch_graph, s));
    h->launches += n_steps * 6;   // K1, K2, K3, K4 (two kernels), K5 per step
    *done = true;
    return CRB_OK;
}

extern "C" int crb_train_epoch_bpr(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_opt* opt, uint64_t seed,
                                   uint32_t epoch, int64_t first, int64_t batch, int64_t n_steps, int32_t neg_ratio, float reg,
                                   double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    OptDev od;
    int opt_kind = 0;
    CRB_CHECK_ARG(n_steps >= 1, "n_steps");
    int rc = bpr_common_checks(h, P, Q, opt, &od, &opt_kind, batch, n_steps, s);
    if (rc) return rc;
    const int64_t rows = crb_epoch_rows(h, neg_ratio, 0);
    CRB_CHECK_ARG(first >= 0 && first < rows, "first row outside the epoch");
    const bool host_loss = !(loss_out && crb_is_device_ptr(loss_out));
    if (epoch_graph_eligible(h, batch, n_steps) && !crb_bpr_ring_enabled(P->dim)) {
        if (h->alt_active) crb_alt_swap(h);
        bool done = false;
        if ((rc = bpr_epoch_graph(h, P, Q, opt, seed, epoch, first, batch, n_steps, rows, neg_ratio, reg, s, &done))) return rc;
        if (done) {
            if (loss_out && !host_loss) CRB_CUDA(cudaMemcpyAsync(loss_out, h->loss_dev, sizeof(double) * n_steps, cudaMemcpyDeviceToDevice, s));
            if (loss_out && host_loss) {
                if ((rc = finish_loss(h, loss_out, n_steps, s))) return rc;
                unsigned int err = 0;
                CRB_CUDA(cudaMemcpy(&err, &h->ctr->sampler_err, sizeof(err), cudaMemcpyDeviceToHost));
                if (err) {
                    CRB_CUDA(cudaMemset(&h->ctr->sampler_err, 0, sizeof(err)));
                    crb_set_error("sampler: %u rows found no admissible negative", err);
                    return CRB_ERR_SAMPLER;
                }
            }
            return CRB_OK;
        }
    }
    crb_opt step_opt = *opt;
    // Two copies of the sampling / assignment state: step k's K1 (sampler + row counts) and K2 (slot assignment) run on the
    // auxiliary stream into copy k & 1 while step k-1's K3/K4 (the HBM-bound part) run on the caller's stream from the other
    // copy.  K1/K2 never touch the tables, so the only ordering needed is: prepare(k) after compute(k-2) released the copy,
    // compute(k) after prepare(k).
    const bool overlap = n_steps > 1;
    struct Restore { crb_handle* h; ~Restore() { if (h->alt_active) crb_alt_swap(h); } } restore{h};
    cudaStream_t ps = s;
    if (overlap) {
        if ((rc = crb_alt_reserve(h, s))) return rc;
        ps = h->aux_stream;
        CRB_CUDA(cudaEventRecord(h->ev_entry, s));
        CRB_CUDA(cudaStreamWaitEvent(ps, h->ev_entry, 0));
    }
    for (int64_t k = 0; k < n_steps; ++k) {
        const int64_t lo = first + k * batch;
        if (lo >= rows) { crb_set_error("step %lld starts past the end of the epoch", (long long)k); return CRB_ERR_ARG; }
        const int64_t b = (rows - lo) < batch ? (rows - lo) : batch;
        const int set = (int)(k & 1);
        if (overlap && h->alt_active != set) crb_alt_swap(h);
        step_opt.step = opt->step + k;
        if ((rc = crb_opt_to_dev(h, &step_opt, &od, &opt_kind, s))) return rc;
        if (overlap && k >= 2) CRB_CUDA(cudaStreamWaitEvent(ps, h->ev_done[set], 0));
        if ((rc = crb_zero_step_counters(h, ps))) return rc;
        rc = crb_launch_sample_pairwise(h, seed, epoch, lo, b, neg_ratio, h->idx[0], h->idx[1], h->idx[2], nullptr, true, ps);
        if (rc) return rc;
        if ((rc = bpr_step_prepare(h, h->idx[0], h->idx[1], h->idx[2], b, true, ps))) return rc;
        if (overlap) {
            CRB_CUDA(cudaEventRecord(h->ev_prep[set], ps));
            CRB_CUDA(cudaStreamWaitEvent(s, h->ev_prep[set], 0));
        }
        double* ld = host_loss ? h->loss_dev + k : loss_out + k;
        rc = bpr_step_compute(h, P, Q, od, opt_kind, h->idx[0], h->idx[1], h->idx[2], b, reg, ld, s);
        if (rc) return rc;
        if (overlap) CRB_CUDA(cudaEventRecord(h->ev_done[set], s));
    }
    // the sampler's sticky error word lives in the step counters: fold the alternate copy's into the primary's
    if (overlap) {
        if (h->alt_active) crb_alt_swap(h);
        merge_sampler_err_kernel<<<1, 1, 0, s>>>(h->ctr, h->alt.ctr);
        CRB_CUDA(cudaGetLastError());
    }
    if (loss_out && host_loss) {
        rc = finish_loss(h, loss_out, n_steps, s);
        if (rc) return rc;
    }
    // sampler attempt overflow is sticky in ctr->sampler_err; report it when the caller synchronises on a host loss
    if (loss_out && host_loss) {
        unsigned int err = 0;
        CRB_CUDA(cudaMemcpy(&err, &h->ctr->sampler_err, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) {
            CRB_CUDA(cudaMemset(&h->ctr->sampler_err, 0, sizeof(err)));
            crb_set_error("sampler: %u rows found no admissible negative", err);
            return CRB_ERR_SAMPLER;
        }
    }
    return CRB_OK;
}

// RankingRecommender.train_model's loop (:39-46) over an epoch the CALLER sampled (the reference's own sampler output, host arrays):
//   for id in range(train_batches): sess.run([train, loss], {u_idx: u[id*B:(id+1)*B], ...})
// as one call.  Step k+1's feed is copied host -> device and counted / assigned on the auxiliary stream into the alternate copy of the
// step state while step k computes (the copy engine runs beside the SMs, so the 12*B bytes of a feed cost no step time); every step's
// loss goes back to the host with its own asynchronous copy when loss_out is page-locked memory (one copy at the end otherwise).
extern "C" int crb_train_epoch_bpr_feeds(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_opt* opt, const int32_t* u,
                                         const int32_t* i, const int32_t* j, int64_t n_rows, int64_t batch, float reg, double* loss_out,
                                         void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    OptDev od;
    int opt_kind = 0;
    CRB_CHECK_ARG(u && i && j && n_rows >= 1, "null / empty feed");
    int rc = bpr_common_checks(h, P, Q, opt, &od, &opt_kind, batch, (n_rows + batch - 1) / batch, s);
    if (rc) return rc;
    const int64_t n_steps = (n_rows + batch - 1) / batch;
    const bool dev_feed = crb_is_device_ptr(u);
    CRB_CHECK_ARG(dev_feed == crb_is_device_ptr(i) && dev_feed == crb_is_device_ptr(j), "u / i / j must all be host or all be device arrays");
    const bool host_loss = !(loss_out && crb_is_device_ptr(loss_out));
    bool pinned_loss = false;
    if (loss_out && host_loss) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, loss_out) == cudaSuccess) pinned_loss = at.type == cudaMemoryTypeHost; else cudaGetLastError();
    }
    crb_opt step_opt = *opt;
    const bool overlap = n_steps > 1;
    struct Restore { crb_handle* h; ~Restore() { if (h->alt_active) crb_alt_swap(h); } } restore{h};
    cudaStream_t ps = s;
    if (overlap) {
        if ((rc = crb_alt_reserve(h, s))) return rc;
        ps = h->aux_stream;
        CRB_CUDA(cudaEventRecord(h->ev_entry, s));
        CRB_CUDA(cudaStreamWaitEvent(ps, h->ev_entry, 0));
    }
    for (int64_t k = 0; k < n_steps; ++k) {
        const int64_t lo = k * batch;
        const int64_t b = (n_rows - lo) < batch ? (n_rows - lo) : batch;
        const int set = (int)(k & 1);
        if (overlap && h->alt_active != set) crb_alt_swap(h);
        step_opt.step = opt->step + k;
        if ((rc = crb_opt_to_dev(h, &step_opt, &od, &opt_kind, s))) return rc;
        if (overlap && k >= 2) CRB_CUDA(cudaStreamWaitEvent(ps, h->ev_done[set], 0));
        if ((rc = crb_zero_step_counters(h, ps))) return rc;
        const int32_t *du = u + lo, *di = i + lo, *dj = j + lo;
        if (!dev_feed) {
            CRB_CUDA(cudaMemcpyAsync(h->idx[0], u + lo, sizeof(int32_t) * b, cudaMemcpyHostToDevice, ps));
            CRB_CUDA(cudaMemcpyAsync(h->idx[1], i + lo, sizeof(int32_t) * b, cudaMemcpyHostToDevice, ps));
            CRB_CUDA(cudaMemcpyAsync(h->idx[2], j + lo, sizeof(int32_t) * b, cudaMemcpyHostToDevice, ps));
            du = h->idx[0]; di = h->idx[1]; dj = h->idx[2];
        }
        if ((rc = bpr_step_prepare(h, du, di, dj, b, false, ps))) return rc;
        if (overlap) {
            CRB_CUDA(cudaEventRecord(h->ev_prep[set], ps));
            CRB_CUDA(cudaStreamWaitEvent(s, h->ev_prep[set], 0));
        }
        double* ld = host_loss ? h->loss_dev + k : loss_out + k;
        if ((rc = bpr_step_compute(h, P, Q, od, opt_kind, du, di, dj, b, reg, ld, s))) return rc;
        if (pinned_loss) CRB_CUDA(cudaMemcpyAsync(loss_out + k, ld, sizeof(double), cudaMemcpyDeviceToHost, s));
        if (overlap) CRB_CUDA(cudaEventRecord(h->ev_done[set], s));
    }
    if (overlap && h->alt_active) crb_alt_swap(h);
    if (loss_out && host_loss) {
        if (pinned_loss) CRB_CUDA(cudaStreamSynchronize(s));
        else if ((rc = finish_loss(h, loss_out, n_steps, s))) return rc;
    }
    return CRB_OK;
}

// ------------------------------------------------------------------------------------------------ Adam flush
template <int OPT>
__global__ void __launch_bounds__(256) adam_flush_kernel(TableDev T, int64_t rows, int dim, OptDev o) {
    // thread per float4 chunk; brings the row up to o.step (inclusive) with decay-only steps
    const int64_t chunks_per_row = dim / 4;
    const int64_t n = rows * chunks_per_row;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const int64_t row = k / chunks_per_row;
        const int last = T.last[row];
        if (last == 0 || last >= o.step) continue;  // never updated (m = v = 0: decay steps are exact no-ops) or up to date
        float4 W = ld4(T.w + k * 4), M = ld4(T.s1 + k * 4), V = ld4(T.s2 + k * 4);
        for (int s = last + 1; s <= o.step; ++s) adam_decay4(W, M, V, lrt_at(o, s), o);
        st4(T.w + k * 4, W); st4(T.s1 + k * 4, M); st4(T.s2 + k * 4, V);
    }
}

__global__ void __launch_bounds__(256) set_last_kernel(int32_t* last, int64_t rows, int32_t step) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < rows; k += stride)
        if (last[k] != 0 && last[k] < step) last[k] = step;
}

extern "C" int crb_adam_flush(crb_handle* h, const crb_table* T, const crb_opt* opt, void* stream) {
    CRB_CHECK_ARG(h && T && opt, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (opt->kind != CRB_OPT_ADAM || opt->adam_mode != CRB_ADAM_TF1) return CRB_OK;
    if (opt->step < 1) return CRB_OK;  // nothing applied yet
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    if ((rc = crb_table_check(T, opt_kind, "T"))) return rc;
    const int64_t n = T->rows * (T->dim / 4);
    adam_flush_kernel<OPT_ADAM_TF1><<<flat_grid(h, n, 256), 256, 0, s>>>(crb_to_dev(T), T->rows, T->dim, od);
    set_last_kernel<<<flat_grid(h, T->rows, 256), 256, 0, s>>>(T->last, T->rows, od.step);
    h->launches += 2;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

// ------------------------------------------------------------------------------------------------ K3: pointwise MF / GMF
// sess.run([train, loss], {u_idx, i_idx, y}): logit = sum_k p_uk q_ik (MF) or sum_k (p_uk q_ik) h_k (GMF.py:43);
// loss = get_loss(loss_func, y, logits) + reg*(l2(p_u) + l2(q_i)) (GMF.py:48).  h is a dense variable: its gradient is
// reduced per block in a fixed order, summed over blocks by dense_apply_kernel, which then applies TF's dense optimizer.
struct PwArgs {
    TableDev P, Q;
    unsigned long long* metaU;
    unsigned long long* metaI;
    const int32_t* u;
    const int32_t* i;
    const float* y;
    const uint32_t* rk[2];
    const float* hvec;     // NULL for MF
    float* hpart;          // [gridDim.x, dim] per-block partial gradient of h
    int64_t batch;
    int dim;
    int loss_kind;
    float reg;
    OptDev opt;
    float* dup_grad;
    uint32_t* dup_t;
    double* block_loss;
};

template <int LANES, int VPL, int OPT, bool GMF>
__global__ void __launch_bounds__(256) pointwise_step_kernel(PwArgs a) {
    asm volatile("griddepcontrol.launch_dependents;");   // see bpr_step_kernel
    constexpr int GPW = 32 / LANES;
    __shared__ float4 s_h[GMF ? 256 * VPL : 1];
    const int lane = threadIdx.x & 31;
    const int gl = lane % LANES;
    const int sub = lane / LANES;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float4 hreg[VPL], gh[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c = (gl + LANES * v) * 4;
        hreg[v] = (GMF && c < a.dim) ? ld4(a.hvec + c) : make_float4(1.f, 1.f, 1.f, 1.f);
        gh[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    double loss_acc = 0.0;
    for (int64_t base = warp * GPW; base < a.batch; base += n_warps * GPW) {
        const int64_t t = base + sub;
        const bool active = t < a.batch;
        const int64_t tt = active ? t : a.batch - 1;
        const int32_t u = a.u[tt], i = a.i[tt];
        const float y = a.y[tt];
        const unsigned long long mu = a.metaU[u], mi = a.metaI[i];
        RowRegs<LANES, VPL> ru, ri;
        row_load_w<LANES, VPL>(ru, a.P, u, a.dim, gl);
        row_load_w<LANES, VPL>(ri, a.Q, i, a.dim, gl);
        ru.last = OptTraits<OPT>::replay ? a.P.last[u] : 0;
        ri.last = OptTraits<OPT>::replay ? a.Q.last[i] : 0;
        const bool su = (uint32_t)mu == 1u || replay_pending<OPT>(ru.last, a.opt);
        const bool si = (uint32_t)mi == 1u || replay_pending<OPT>(ri.last, a.opt);
        if (OptTraits<OPT>::has_s1) {
            if (su) row_load_state<LANES, VPL, OPT>(ru, a.P, u, a.dim, gl);
            if (si) row_load_state<LANES, VPL, OPT>(ri, a.Q, i, a.dim, gl);
        }
        if (replay_pending<OPT>(ru.last, a.opt)) row_replay<LANES, VPL, OPT>(ru, a.opt, a.opt.step);
        if (replay_pending<OPT>(ri.last, a.opt)) row_replay<LANES, VPL, OPT>(ri, a.opt, a.opt.step);
        float x = 0.f, sq = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const float4 p = ru.w[v], q = ri.w[v], hh = hreg[v];
            const float4 pq = make_float4(p.x * q.x, p.y * q.y, p.z * q.z, p.w * q.w);
            x += GMF ? dot4(pq, hh) : (pq.x + pq.y + pq.z + pq.w);
            sq += dot4(p, p) + dot4(q, q);
        }
        x = group_sum<LANES>(x);
        sq = group_sum<LANES>(sq);
        float g, l;
        if (a.loss_kind == CRB_LOSS_CROSS_ENTROPY) {  // utils/tools.py:68-69
            l = fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
            g = sigmoid_f(x) - y;
        } else {                                       // 'square', utils/tools.py:74-75
            l = (y - x) * (y - x);
            g = 2.f * (x - y);
        }
        if (active && gl == 0) loss_acc += (double)(l + a.reg * 0.5f * sq);
        float4 gu[VPL], gi[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const float4 p = ru.w[v], q = ri.w[v], hh = hreg[v];
            gu[v] = make_float4(fmaf(g, q.x * hh.x, a.reg * p.x), fmaf(g, q.y * hh.y, a.reg * p.y), fmaf(g, q.z * hh.z, a.reg * p.z),
                                fmaf(g, q.w * hh.w, a.reg * p.w));
            gi[v] = make_float4(fmaf(g, p.x * hh.x, a.reg * q.x), fmaf(g, p.y * hh.y, a.reg * q.y), fmaf(g, p.z * hh.z, a.reg * q.z),
                                fmaf(g, p.w * hh.w, a.reg * q.w));
            if (GMF && active) {
                gh[v].x = fmaf(g, p.x * q.x, gh[v].x); gh[v].y = fmaf(g, p.y * q.y, gh[v].y);
                gh[v].z = fmaf(g, p.z * q.z, gh[v].z); gh[v].w = fmaf(g, p.w * q.w, gh[v].w);
            }
        }
        if (active) {
            emit_row<LANES, VPL, OPT>(ru, gu, a.P, a.metaU, u, mu, a.rk[0][t], (uint32_t)t, 0u, a.dim, gl, a.opt, a.dup_grad, a.dup_t);
            emit_row<LANES, VPL, OPT>(ri, gi, a.Q, a.metaI, i, mi, a.rk[1][t], (uint32_t)t, 1u, a.dim, gl, a.opt, a.dup_grad, a.dup_t);
        }
    }
    if (GMF) {
        // fixed-order block reduction of the h gradient: thread -> smem, then one thread per float4 chunk sums the groups
#pragma unroll
        for (int v = 0; v < VPL; ++v) s_h[threadIdx.x * VPL + v] = gh[v];
        __syncthreads();
        constexpr int GROUPS = 256 / LANES;
        for (int c = threadIdx.x; c < LANES * VPL; c += blockDim.x) {
            const int l = c % LANES, v = c / LANES;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int gI = 0; gI < GROUPS; ++gI) {
                const float4 t4 = s_h[(gI * LANES + l) * VPL + v];
                acc.x += t4.x; acc.y += t4.y; acc.z += t4.z; acc.w += t4.w;
            }
            const int col = (l + LANES * v) * 4;
            if (col < a.dim) st4(a.hpart + (int64_t)blockIdx.x * a.dim + col, acc);
        }
    }
    block_loss_store(loss_acc, a.block_loss);
}

// TF dense apply of a small dense variable (ApplyGradientDescent / ApplyAdagrad / ApplyAdam) whose gradient arrives as
// per-block partials [n_parts, n]; summed in a fixed order (block_sum_parts: deterministic).
__global__ void __launch_bounds__(256) dense_apply_kernel(float* w, float* s1, float* s2, const float* parts, int n_parts, int n,
                                                         int opt_kind, OptDev o, int part_stride = 0) {
    const int64_t ps = part_stride > 0 ? part_stride : n;   // distance between consecutive partial vectors
    for (int k0 = blockIdx.x * 32; k0 < n; k0 += gridDim.x * 32) {
        const int k = k0 + (threadIdx.x & 31);
        const float g = block_sum_parts(parts, n_parts, ps, k, k < n);
        if (threadIdx.x >= 32 || k >= n) continue;
        float x = w[k];
        if (opt_kind == OPT_SGD) {
            x = __fsub_rn(x, __fmul_rn(o.lr, g));
        } else if (opt_kind == OPT_ADAGRAD) {
            float acc = __fadd_rn(s1[k], __fmul_rn(g, g));
            s1[k] = acc;
            x = __fsub_rn(x, __fdiv_rn(__fmul_rn(o.lr, g), __fsqrt_rn(acc)));
        } else {  // ApplyAdam: m += (g-m)(1-b1); v += (g*g-v)(1-b2); var -= (m*lr_t)/(sqrt(v)+eps)
            float m = s1[k], v = s2[k];
            m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), o.omb1));
            v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), o.omb2));
            s1[k] = m; s2[k] = v;
            x = __fsub_rn(x, __fdiv_rn(__fmul_rn(m, o.lr_t), __fadd_rn(__fsqrt_rn(v), o.eps)));
        }
        w[k] = x;
    }
}

int crb_launch_dense_apply(crb_handle* h, float* w, float* s1, float* s2, const float* parts, int n_parts, int n, int opt_kind, const OptDev& od,
                           cudaStream_t s) {
    dense_apply_kernel<<<(n + 31) / 32, 256, 0, s>>>(w, s1, s2, parts, n_parts, n, opt_kind == OPT_ADAM_TF1 ? OPT_ADAM_LAZY : opt_kind, od);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

int crb_launch_dense_apply_strided(crb_handle* h, float* w, float* s1, float* s2, const float* parts, int n_parts, int n, int part_stride,
                                   int opt_kind, const OptDev& od, cudaStream_t s) {
    dense_apply_kernel<<<(n + 31) / 32, 256, 0, s>>>(w, s1, s2, parts, n_parts, n, opt_kind == OPT_ADAM_TF1 ? OPT_ADAM_LAZY : opt_kind, od, part_stride);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

template <int LANES, int VPL>
static int launch_pw_t(crb_handle* h, const PwArgs& a, int opt_kind, bool gmf, cudaStream_t s) {
    const int gpb = 256 / LANES;
#define CRB_PW_CASE(O)                                                                                       \
    case O:                                                                                                  \
        if (gmf) {                                                                                           \
            const int grid = persistent_grid(h, pointwise_step_kernel<LANES, VPL, O, true>, a.batch, gpb);   \
            pointwise_step_kernel<LANES, VPL, O, true><<<grid, 256, 0, s>>>(a);                              \
        } else {                                                                                             \
            const int grid = persistent_grid(h, pointwise_step_kernel<LANES, VPL, O, false>, a.batch, gpb);  \
            pointwise_step_kernel<LANES, VPL, O, false><<<grid, 256, 0, s>>>(a);                             \
        }                                                                                                    \
        break;
    switch (opt_kind) {
        CRB_PW_CASE(OPT_SGD)
        CRB_PW_CASE(OPT_ADAGRAD)
        CRB_PW_CASE(OPT_ADAM_LAZY)
        CRB_PW_CASE(OPT_ADAM_TF1)
    }
#undef CRB_PW_CASE
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

static int pointwise_checks(crb_handle* h, int32_t kind, const crb_table* P, const crb_table* Q, float* hvec, float* h_s1, float* h_s2,
                            const crb_opt* opt, int32_t loss_kind, int64_t batch, int64_t steps, OptDev* od, int* opt_kind, cudaStream_t s) {
    int rc = bpr_common_checks(h, P, Q, opt, od, opt_kind, batch, steps, s);
    if (rc) return rc;
    CRB_CHECK_ARG(kind == CRB_SCORE_DOT || kind == CRB_SCORE_GMF, "kind must be CRB_SCORE_DOT (MF) or CRB_SCORE_GMF");
    CRB_CHECK_ARG(loss_kind == CRB_LOSS_CROSS_ENTROPY || loss_kind == CRB_LOSS_SQUARE, "pointwise loss must be cross_entropy or square");
    if (kind == CRB_SCORE_GMF) {
        CRB_CHECK_ARG(hvec && crb_is_device_ptr(hvec), "GMF needs the device vector h");
        CRB_CHECK_ARG(*opt_kind == OPT_SGD || h_s1, "h optimizer slot s1 is NULL");
        CRB_CHECK_ARG((*opt_kind != OPT_ADAM_LAZY && *opt_kind != OPT_ADAM_TF1) || h_s2, "h optimizer slot s2 is NULL");
        const int64_t need = (int64_t)h->loss_blocks * P->dim;
        if (need > h->cap_dense) {
            CRB_CUDA(cudaStreamSynchronize(s));
            cudaFree(h->dense_grad);
            h->dense_grad = nullptr;
            CRB_CUDA(cudaMalloc(&h->dense_grad, sizeof(float) * need));
            h->cap_dense = need;
        }
    }
    return CRB_OK;
}

// K2 .. K5 of one pointwise step; `counted`: the sampler already bumped the multiplicities (K1)
static int pointwise_step_device(crb_handle* h, bool gmf, const crb_table* P, const crb_table* Q, float* hvec, float* h_s1, float* h_s2,
                                 const OptDev& od, int opt_kind, int32_t loss_kind, const int32_t* du, const int32_t* di, const float* dy,
                                 int64_t batch, float reg, bool counted, double* loss_dev, cudaStream_t s, bool assigned = false) {
    int rc;
    const int32_t* idx[3] = {du, di, nullptr};
    const int role_table[3] = {0, 1, 0};
    if (!counted && (rc = crb_count_rows(h, batch, 2, idx, role_table, s))) return rc;
    if (!assigned && (rc = crb_launch_assign(h, batch, 2, idx, role_table, s))) return rc;
    PwArgs a;
    a.P = crb_to_dev(P); a.Q = crb_to_dev(Q);
    a.metaU = h->meta[0]; a.metaI = h->meta[1];
    a.u = du; a.i = di; a.y = dy;
    a.rk[0] = h->rank[0]; a.rk[1] = h->rank[1];
    a.hvec = gmf ? hvec : nullptr; a.hpart = h->dense_grad;
    a.batch = batch; a.dim = P->dim; a.loss_kind = loss_kind; a.reg = reg; a.opt = od;
    a.dup_grad = h->dup_grad; a.dup_t = h->dup_t; a.block_loss = h->block_loss;
    if ((rc = crb_prof_begin(h, s))) return rc;
    rc = CRB_DIM_DISPATCH(a.dim, launch_pw_t, h, a, opt_kind, gmf, s);
    if (rc) return rc;
    if ((rc = crb_prof_end(h, s))) return rc;
    DupArgs d;
    d.tab[0] = a.P; d.tab[1] = a.Q;
    d.meta[0] = h->meta[0]; d.meta[1] = h->meta[1];
    d.dim = a.dim; d.opt = od;
    d.dup_rows = h->dup_rows; d.work = h->work; d.multi = h->multi;
    d.dup_grad = h->dup_grad; d.dup_t = h->dup_t; d.partial = h->partial; d.ctr = h->ctr;
    const bool tail = crb_dup_tail_enabled(batch);
    if ((rc = tail ? crb_launch_dup_tail(h, d, opt_kind, loss_dev, s) : crb_launch_dup_pipeline(h, d, opt_kind, s))) return rc;
    if (gmf) {
        dense_apply_kernel<<<(P->dim + 31) / 32, 256, 0, s>>>(hvec, h_s1, h_s2, h->dense_grad, h->step_grid, P->dim, opt_kind == OPT_ADAM_TF1 ? OPT_ADAM_LAZY : opt_kind, od);
        h->launches++;
        CRB_CUDA(cudaGetLastError());
    }
    return tail ? CRB_OK : crb_launch_loss_final(h, loss_dev, s);
}

extern "C" int crb_train_step_pointwise(crb_handle* h, int32_t kind, const crb_table* P, const crb_table* Q, float* hvec, float* h_s1,
                                        float* h_s2, const crb_opt* opt, int32_t loss_kind, const int32_t* u, const int32_t* i,
                                        const float* y, int64_t batch, float reg, double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    OptDev od;
    int opt_kind = 0;
    int rc = pointwise_checks(h, kind, P, Q, hvec, h_s1, h_s2, opt, loss_kind, batch, 1, &od, &opt_kind, s);
    if (rc) return rc;
    CRB_CHECK_ARG(u && i && y, "null feed");
    const int32_t *du, *di;
    if ((rc = stage_i32(h, u, 0, batch, &du, s, P->rows, "u_idx"))) return rc;
    if ((rc = stage_i32(h, i, 1, batch, &di, s, Q->rows, "i_idx"))) return rc;
    const float* dy = y;
    if (!crb_is_device_ptr(y)) {
        CRB_CUDA(cudaMemcpyAsync(h->yv, y, sizeof(float) * batch, cudaMemcpyHostToDevice, s));
        dy = h->yv;
    }
    if ((rc = crb_zero_step_counters(h, s))) return rc;
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    if ((rc = pointwise_step_device(h, kind == CRB_SCORE_GMF, P, Q, hvec, h_s1, h_s2, od, opt_kind, loss_kind, du, di, dy, batch, reg, false, ld, s)))
        return rc;
    return finish_loss(h, loss_out, 1, s);
}

// RankingRecommender.train_model's pointwise loop (:48-60) with the sampler fused in: n_steps iterations in one call, step k trains on
// epoch rows [first + k*batch, min(first + (k+1)*batch, epoch_rows)) of pointwise_ranking_sampler's epoch (utils/sampler.py:10-43).
// One stream, no host round trips: at the batch sizes the reference ships (6144) a step is a handful of microsecond kernels and the
// per-call host cost of stepping from Python was twice the device time.
extern "C" int crb_train_epoch_pointwise(crb_handle* h, int32_t kind, const crb_table* P, const crb_table* Q, float* hvec, float* h_s1,
                                         float* h_s2, const crb_opt* opt, int32_t loss_kind, uint64_t seed, uint32_t epoch, int64_t first,
                                         int64_t batch, int64_t n_steps, int32_t neg_ratio, float reg, double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    OptDev od;
    int opt_kind = 0;
    CRB_CHECK_ARG(n_steps >= 1, "n_steps");
    int rc = pointwise_checks(h, kind, P, Q, hvec, h_s1, h_s2, opt, loss_kind, batch, n_steps, &od, &opt_kind, s);
    if (rc) return rc;
    const int64_t rows = crb_epoch_rows(h, neg_ratio, 1);
    CRB_CHECK_ARG(first >= 0 && first < rows, "first row outside the epoch");
    const bool host_loss = !(loss_out && crb_is_device_ptr(loss_out));
    crb_opt step_opt = *opt;
    // As in crb_train_epoch_bpr: step k's sampler and slot assignment run on the auxiliary stream into copy k & 1 of the step state
    // while step k-1's K3 / tail / dense apply run on the caller's stream from the other copy.  The labels travel in the copy's third
    // index buffer (unused by a two-role step; same size, reinterpreted as float), so they are double-buffered with the indices.
    const bool overlap = n_steps > 1;
    struct Restore { crb_handle* h; ~Restore() { if (h->alt_active) crb_alt_swap(h); } } restore{h};
    cudaStream_t ps = s;
    if (overlap) {
        if ((rc = crb_alt_reserve(h, s))) return rc;
        ps = h->aux_stream;
        CRB_CUDA(cudaEventRecord(h->ev_entry, s));
        CRB_CUDA(cudaStreamWaitEvent(ps, h->ev_entry, 0));
    }
    for (int64_t k = 0; k < n_steps; ++k) {
        const int64_t lo = first + k * batch;
        if (lo >= rows) { crb_set_error("step %lld starts past the end of the epoch", (long long)k); return CRB_ERR_ARG; }
        const int64_t b = (rows - lo) < batch ? (rows - lo) : batch;
        const int set = (int)(k & 1);
        if (overlap && h->alt_active != set) crb_alt_swap(h);
        step_opt.step = opt->step + k;
        if ((rc = crb_opt_to_dev(h, &step_opt, &od, &opt_kind, s))) return rc;
        if (overlap && k >= 2) CRB_CUDA(cudaStreamWaitEvent(ps, h->ev_done[set], 0));
        if ((rc = crb_zero_step_counters(h, ps))) return rc;
        float* labels = reinterpret_cast<float*>(h->idx[2]);
        if ((rc = crb_launch_sample_pointwise(h, seed, epoch, lo, b, neg_ratio, h->idx[0], h->idx[1], labels, true, ps))) return rc;
        {
            const int32_t* idx[3] = {h->idx[0], h->idx[1], nullptr};
            const int role_table[3] = {0, 1, 0};
            if ((rc = crb_launch_assign(h, b, 2, idx, role_table, ps))) return rc;
        }
        if (overlap) {
            CRB_CUDA(cudaEventRecord(h->ev_prep[set], ps));
            CRB_CUDA(cudaStreamWaitEvent(s, h->ev_prep[set], 0));
        }
        double* ld = host_loss ? h->loss_dev + k : loss_out + k;
        if ((rc = pointwise_step_device(h, kind == CRB_SCORE_GMF, P, Q, hvec, h_s1, h_s2, od, opt_kind, loss_kind, h->idx[0], h->idx[1], labels, b, reg,
                                        true, ld, s, true)))
            return rc;
        if (overlap) CRB_CUDA(cudaEventRecord(h->ev_done[set], s));
    }
    if (overlap) {   // the sampler's sticky error word lives in the step counters: fold the alternate copy's into the primary's
        if (h->alt_active) crb_alt_swap(h);
        merge_sampler_err_kernel<<<1, 1, 0, s>>>(h->ctr, h->alt.ctr);
        CRB_CUDA(cudaGetLastError());
    }
    if (loss_out && host_loss) {
        if ((rc = finish_loss(h, loss_out, n_steps, s))) return rc;
        unsigned int err = 0;
        CRB_CUDA(cudaMemcpy(&err, &h->ctr->sampler_err, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) {
            CRB_CUDA(cudaMemset(&h->ctr->sampler_err, 0, sizeof(err)));
            crb_set_error("sampler: %u rows found no admissible negative", err);
            return CRB_ERR_SAMPLER;
        }
    }
    return CRB_OK;
}
