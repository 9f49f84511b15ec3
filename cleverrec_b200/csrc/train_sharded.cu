// Multi-GPU BPR step over NVLink peer memory.  One process per GPU.  Users (rows of P, their histories, their
// sampling) are partitioned across ranks; the item table Q is row-sharded (owner = item % G, local row = item / G) and every
// rank maps every shard through CUDA IPC, so no NCCL all-to-all, no id de-duplication and no host synchronisation is needed:
//
//   phase 1  shard_step_kernel   each rank, for its own B triplets: gathers p_u locally and q_i, q_j STRAIGHT FROM THE OWNER'S HBM
//                                (peer loads over NVLink), forward/backward, applies the user row locally (same multiplicity
//                                machinery as the single-GPU step) and writes the two item gradients into the OWNER's inbox:
//                                one system-scope atomic claims a slot, then 128-bit peer stores.
//   -- cross-rank barrier (a one-element NCCL all-reduce enqueued on the same stream by the caller) --
//   phase 2  inbox_apply_kernel  each owner de-duplicates its inbox (count -> assign), applies rows that arrived once in place and
//                                reduces the others through the duplicate-slot pipeline: one optimizer apply per unique row with
//                                the gradient summed over ALL ranks -- the semantics of one TF step on the union batch.
//   -- barrier --
#include <cstddef>
#include <stdlib.h>

#include "rowopt.cuh"

struct ShardDev {
    int n_ranks, rank;
    int64_t inbox_cap;
    TableDev q[CRB_MAX_RANKS];
    float* inbox_grad[CRB_MAX_RANKS];
    int32_t* inbox_row[CRB_MAX_RANKS];
    uint32_t* inbox_key[CRB_MAX_RANKS];
    unsigned int* inbox_cnt[CRB_MAX_RANKS];   // [0] entries, [1] overflow flag
};

struct ShardStepArgs {
    TableDev P;
    unsigned long long* metaU;
    ShardDev sh;
    const int32_t* u;   // local user rows
    const int32_t* i;   // GLOBAL item ids
    const int32_t* j;
    const uint32_t* rk_u;
    int64_t batch;
    int dim;
    float reg;
    OptDev opt;
    float* dup_grad;
    uint32_t* dup_t;
    double* block_loss;
    int debug;   // experiments only (-DSH_DEBUG_SWITCHES + CRB_SH_DEBUG): 1 = no gradient sends, 2 = item rows read from the local shard, 3 = both
};

#define SH_CHUNK 32u   // inbox slots a warp reserves per system-scope atomic (unused ones stay holes: row = -1)

// Takes one slot of the owner's inbox for every active group of the warp and stores the gradient there.  Slots are handed out
// from per-warp reservations of SH_CHUNK (one system-scope atomic on the owner's counter per 32 gradients instead of one per
// gradient: the single hot counter was the bottleneck of the first version).
template <int LANES, int VPL>
__device__ __forceinline__ void send_item_grad(const ShardDev& sh, unsigned int* s_base, unsigned int* s_left, int owner, int32_t local_row,
                                               uint32_t key, const float4* g, int dim, int gl, int sub, bool active) {
    constexpr int GPW = 32 / LANES;
    unsigned int slot = 0xFFFFFFFFu;
#pragma unroll
    for (int q = 0; q < GPW; ++q) {
        if (sub == q && gl == 0 && active) {
            if (s_left[owner] == 0u) { s_base[owner] = atomicAdd_system(sh.inbox_cnt[owner], SH_CHUNK); s_left[owner] = SH_CHUNK; }
            slot = s_base[owner]++;
            s_left[owner]--;
        }
        __syncwarp();
    }
    slot = __shfl_sync(0xffffffffu, slot, sub * LANES);
    if (!active) return;
    if ((int64_t)slot >= sh.inbox_cap) {
        if (gl == 0) atomicExch_system(sh.inbox_cnt[owner] + 1, 1u);   // overflow: reported by crb_shard_inbox_overflow
        return;
    }
    if (gl == 0) { sh.inbox_key[owner][slot] = key; sh.inbox_row[owner][slot] = local_row; }
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c = (gl + LANES * v) * 4;
        if (c < dim) st4(sh.inbox_grad[owner] + (int64_t)slot * dim + c, g[v]);
    }
}

template <int LANES, int VPL, int OPT>
__global__ void __launch_bounds__(256, 3) shard_step_kernel(ShardStepArgs a) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31, gl = lane % LANES, sub = lane / LANES;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int G = a.sh.n_ranks;
#ifdef SH_DEBUG_SWITCHES
    const int RG = (a.debug & 2) ? 1 : G;              // debug: every read goes to the local shard
    const int RO = (a.debug & 2) ? a.sh.rank : 0;
#define SH_Q(item) a.sh.q[(a.debug & 2) ? RO : (item) % RG]
#define SH_SEND (!(a.debug & 1))
#else
#define SH_Q(item) a.sh.q[(item) % G]
#define SH_SEND true
#endif
    __shared__ unsigned int s_resv[8][2][CRB_MAX_RANKS];   // per warp: next slot / slots left of the current reservation per owner
    if (lane < CRB_MAX_RANKS) { s_resv[threadIdx.x >> 5][0][lane] = 0u; s_resv[threadIdx.x >> 5][1][lane] = 0u; }
    __syncwarp();
    unsigned int* s_base = s_resv[threadIdx.x >> 5][0];
    unsigned int* s_left = s_resv[threadIdx.x >> 5][1];
    double loss_acc = 0.0;
    // Software pipeline.  Item rows mostly live on other GPUs (2-3 us away over NVLink), so they are requested TWO iterations ahead;
    // the local user row one iteration ahead; indices three iterations ahead.  Out-of-range iterations clamp to the last triplet
    // (loads only).
    const int64_t stride = n_warps * GPW;
    const int64_t base0 = warp * GPW;
    auto clampt = [&](int64_t b) { const int64_t t_ = b + sub; return t_ < a.batch ? t_ : a.batch - 1; };
    int32_t u1, i1, j1;          // indices of iteration n+1 (its item rows are in flight, its user row is about to be)
    int32_t u2, i2, j2;          // indices of iteration n+2
    int32_t nu, ni, nj;          // indices of the current iteration
    { const int64_t t0 = clampt(base0); nu = a.u[t0]; ni = a.i[t0]; nj = a.j[t0]; }
    { const int64_t t1 = clampt(base0 + stride); u1 = a.u[t1]; i1 = a.i[t1]; j1 = a.j[t1]; }
    { const int64_t t2 = clampt(base0 + 2 * stride); u2 = a.u[t2]; i2 = a.i[t2]; j2 = a.j[t2]; }
    RowRegs<LANES, VPL> nru, nri, nrj, fri, frj;
    row_load_w<LANES, VPL>(nri, SH_Q(ni), ni / G, a.dim, gl);
    row_load_w<LANES, VPL>(nrj, SH_Q(nj), nj / G, a.dim, gl);
    row_load_w<LANES, VPL>(fri, SH_Q(i1), i1 / G, a.dim, gl);
    row_load_w<LANES, VPL>(frj, SH_Q(j1), j1 / G, a.dim, gl);
    row_load_w<LANES, VPL>(nru, a.P, nu, a.dim, gl);
    for (int64_t base = base0; base < a.batch; base += stride) {
        const int64_t t = base + sub;
        const bool active = t < a.batch;
        const int64_t tt = active ? t : a.batch - 1;
        const int32_t u = nu, i = ni, j = nj;
        RowRegs<LANES, VPL> ru = nru, ri = nri, rj = nrj;
        // rotate the pipeline: iteration n+1's item rows were requested last time; request n+2's (peer loads when the owner is
        // another rank) and n+1's user row
        nu = u1; ni = i1; nj = j1;
        nri = fri; nrj = frj;
        row_load_w<LANES, VPL>(nru, a.P, nu, a.dim, gl);
        row_load_w<LANES, VPL>(fri, SH_Q(i2), i2 / G, a.dim, gl);
        row_load_w<LANES, VPL>(frj, SH_Q(j2), j2 / G, a.dim, gl);
        u1 = u2; i1 = i2; j1 = j2;
        { const int64_t t3 = clampt(base + 3 * stride); u2 = a.u[t3]; i2 = a.i[t3]; j2 = a.j[t3]; }
        const int oi = i % G, oj = j % G;
        const int32_t li = i / G, lj = j / G;
        const unsigned long long mu = a.metaU[u];
        // item rows are always current at a step boundary (the owner decays its whole shard in phase 2), so only the local user
        // row can have missed steps to replay
        ru.last = OptTraits<OPT>::replay ? a.P.last[u] : 0;
        const bool su = (uint32_t)mu == 1u || replay_pending<OPT>(ru.last, a.opt);
        if (OptTraits<OPT>::has_s1 && su) row_load_state<LANES, VPL, OPT>(ru, a.P, u, a.dim, gl);
        if (replay_pending<OPT>(ru.last, a.opt)) row_replay<LANES, VPL, OPT>(ru, a.opt, a.opt.step);
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
        const float g = -sigmoid_f(-x);
        if (active && gl == 0) loss_acc += (double)(softplus_neg(x) + a.reg * 0.5f * sq);
        float4 gu[VPL], gi[VPL], gj[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const float4 p = ru.w[v], qi = ri.w[v], qj = rj.w[v];
            gu[v] = make_float4(fmaf(g, qi.x - qj.x, a.reg * p.x), fmaf(g, qi.y - qj.y, a.reg * p.y), fmaf(g, qi.z - qj.z, a.reg * p.z),
                                fmaf(g, qi.w - qj.w, a.reg * p.w));
            gi[v] = make_float4(fmaf(g, p.x, a.reg * qi.x), fmaf(g, p.y, a.reg * qi.y), fmaf(g, p.z, a.reg * qi.z), fmaf(g, p.w, a.reg * qi.w));
            gj[v] = make_float4(fmaf(-g, p.x, a.reg * qj.x), fmaf(-g, p.y, a.reg * qj.y), fmaf(-g, p.z, a.reg * qj.z), fmaf(-g, p.w, a.reg * qj.w));
        }
        if (active)
            emit_row<LANES, VPL, OPT>(ru, gu, a.P, a.metaU, u, mu, a.rk_u[t], (uint32_t)t, 0u, a.dim, gl, a.opt, a.dup_grad, a.dup_t);
        // ordering key of the occurrence: unique and identical from run to run -> deterministic duplicate sums at the owner
        const uint32_t kbase = ((uint32_t)a.sh.rank * (uint32_t)a.batch + (uint32_t)tt) << 1;
        if (SH_SEND) {
            send_item_grad<LANES, VPL>(a.sh, s_base, s_left, oi, li, kbase, gi, a.dim, gl, sub, active);
            send_item_grad<LANES, VPL>(a.sh, s_base, s_left, oj, lj, kbase | 1u, gj, a.dim, gl, sub, active);
        }
    }
    __threadfence_system();   // peer stores visible before the kernel retires (the barrier that follows orders them across ranks)
    block_loss_store(loss_acc, a.block_loss);
}

struct InboxArgs {
    unsigned long long* meta;
    const int32_t* row;
    const uint32_t* key;
    const unsigned int* cnt;
    const uint32_t* rk;
    int64_t cap;
    uint32_t* dup_src;
    uint32_t* dup_t;
};

// Phase 2 never copies a gradient: every row that received at least one gradient owns a slot range (assign_kernel with
// alloc_rank 0), this kernel -- one thread per inbox entry -- only records WHERE the row's gradients are (slot -> inbox entry, plus
// the ordering key), and dup_reduce_kernel sums them straight out of the inbox (DupArgs::dup_src) in key order and applies the
// optimizer once per row.  A gradient that crossed NVLink is read from HBM exactly once.
__global__ void __launch_bounds__(256) inbox_index_kernel(InboxArgs a) {
    int64_t n = *a.cnt;
    if (n > a.cap) n = a.cap;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const int32_t row = a.row[e];
        if (row < 0) continue;   // hole of a partly used reservation
        const uint32_t slot = (uint32_t)(a.meta[row] >> 32) + a.rk[e];
        a.dup_src[slot] = (uint32_t)e;
        a.dup_t[slot] = a.key[e];
    }
}

static int shard_to_dev(const crb_shard* sh, ShardDev* d) {
    CRB_CHECK_ARG(sh && sh->n_ranks >= 1 && sh->n_ranks <= CRB_MAX_RANKS && sh->rank >= 0 && sh->rank < sh->n_ranks, "shard descriptor");
    CRB_CHECK_ARG(sh->inbox_cap > 0, "inbox capacity");
    d->n_ranks = sh->n_ranks; d->rank = sh->rank; d->inbox_cap = sh->inbox_cap;
    for (int r = 0; r < sh->n_ranks; ++r) {
        CRB_CHECK_ARG(sh->q[r].w && sh->inbox_grad[r] && sh->inbox_row[r] && sh->inbox_key[r] && sh->inbox_cnt[r], "shard descriptor: null peer pointer");
        d->q[r] = crb_to_dev(&sh->q[r]);
        d->inbox_grad[r] = sh->inbox_grad[r]; d->inbox_row[r] = sh->inbox_row[r]; d->inbox_key[r] = sh->inbox_key[r];
        d->inbox_cnt[r] = sh->inbox_cnt[r];
    }
    return CRB_OK;
}

template <typename K>
static int one_wave(crb_handle* h, K kernel, int64_t groups, int gpb) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 2; }
    int64_t grid = (int64_t)h->sm_count * occ;
    const int64_t need = (groups + gpb - 1) / gpb;
    if (grid > need) grid = need;
    if (grid > h->loss_blocks) grid = h->loss_blocks;
    if (grid < 1) grid = 1;
    return (int)grid;
}

template <int LANES, int VPL>
static int launch_shard_t(crb_handle* h, const ShardStepArgs& a, int opt_kind, cudaStream_t s) {
    const int gpb = 256 / LANES;
#define CRB_SH_CASE(O)                                                                          \
    case O: {                                                                                   \
        const int grid = one_wave(h, shard_step_kernel<LANES, VPL, O>, a.batch, gpb);           \
        h->step_grid = grid;                                                                    \
        shard_step_kernel<LANES, VPL, O><<<grid, 256, 0, s>>>(a);                               \
        break;                                                                                  \
    }
    switch (opt_kind) { CRB_SH_CASE(OPT_SGD) CRB_SH_CASE(OPT_ADAGRAD) CRB_SH_CASE(OPT_ADAM_LAZY) CRB_SH_CASE(OPT_ADAM_TF1) }
#undef CRB_SH_CASE
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

#define CRB_DIM_DISPATCH(dim, FN, ...)                                   \
    ((dim) <= 32 ? FN<8, 1>(__VA_ARGS__)                                 \
     : (dim) <= 64 ? FN<16, 1>(__VA_ARGS__)                              \
     : (dim) <= 128 ? FN<32, 1>(__VA_ARGS__)                             \
     : (dim) <= 256 ? FN<32, 2>(__VA_ARGS__)                             \
                    : FN<32, 4>(__VA_ARGS__))

static void fill_dup(crb_handle* h, DupArgs* d, const TableDev& t0, const TableDev& t1, int dim, const OptDev& od) {
    d->tab[0] = t0; d->tab[1] = t1;
    d->meta[0] = h->meta[0]; d->meta[1] = h->meta[1];
    d->dim = dim; d->opt = od;
    d->dup_rows = h->dup_rows; d->work = h->work; d->multi = h->multi;
    d->dup_grad = h->dup_grad; d->dup_t = h->dup_t; d->partial = h->partial; d->ctr = h->ctr;
}

// Optional phase 0: sample rows [first, first+batch) of this rank's epoch and count / assign the user rows NOW, on the handle's
// auxiliary stream, into the alternate copy of the step state -- typically called right after crb_shard_step_compute of the
// previous step, so that it overlaps that step's barriers and inbox phase (none of it touches the tables).  The next
// crb_shard_step_compute with u == NULL and the same (seed, epoch, first, neg_ratio, batch) consumes it.  reserve_rows >= batch
// sizes the workspace once for both phases (pass the inbox capacity) so that no later call reallocates it under the prepared step.
// With feed_u / feed_i / feed_j (HOST or DEVICE int32 [batch]: the caller's own triplets, local user rows and global item ids) the
// step is staged from them instead of being sampled -- the host -> device copies then run on the copy stream beside the previous
// step's kernels; (seed, epoch, first, neg_ratio, batch) only serve as the ticket the consuming crb_shard_step_compute presents.
extern "C" int crb_shard_step_prepare(crb_handle* h, const crb_table* P, uint64_t seed, uint32_t epoch, int64_t first, int32_t neg_ratio,
                                      int64_t batch, int64_t reserve_rows, const int32_t* feed_u, const int32_t* feed_i,
                                      const int32_t* feed_j, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && P, "null argument");
    CRB_CHECK_ARG(batch > 0, "batch");
    CRB_CUDA(cudaSetDevice(h->device));
    int rc;
    if ((rc = crb_ws_reserve(h, reserve_rows > batch ? reserve_rows : batch, P->dim, 4, s))) return rc;
    if ((rc = crb_meta_reserve(h, 0, P->rows, s))) return rc;
    if ((rc = crb_alt_reserve(h, s))) return rc;
    cudaStream_t ps = h->aux_stream;
    // the alternate copy is free once the last compute that used it has finished; a freshly zeroed meta needs the caller's stream
    CRB_CUDA(cudaEventRecord(h->ev_entry, s));
    CRB_CUDA(cudaStreamWaitEvent(ps, h->ev_entry, 0));
    crb_alt_swap(h);
    struct Restore { crb_handle* h; ~Restore() { if (h->alt_active) crb_alt_swap(h); } } restore{h};
    if ((rc = crb_zero_step_counters(h, ps))) return rc;
    if (feed_u) {
        CRB_CHECK_ARG(feed_i && feed_j, "null index feed");
        const int32_t* src[3] = {feed_u, feed_i, feed_j};
        for (int q = 0; q < 3; ++q)
            CRB_CUDA(cudaMemcpyAsync(h->idx[q], src[q], sizeof(int32_t) * batch, crb_is_device_ptr(src[q]) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ps));
    } else if ((rc = crb_launch_sample_pairwise(h, seed, epoch, first, batch, neg_ratio, h->idx[0], h->idx[1], h->idx[2], nullptr, false, ps))) return rc;
    const int32_t* idx[3] = {h->idx[0], nullptr, nullptr};
    const int role_table[3] = {0, 0, 0};
    if ((rc = crb_count_rows(h, batch, 1, idx, role_table, ps))) return rc;
    if ((rc = crb_launch_assign(h, batch, 1, idx, role_table, ps))) return rc;
    CRB_CUDA(cudaEventRecord(h->ev_prep[1], ps));
    h->prep_valid = 1; h->prep_seed = seed; h->prep_epoch = epoch; h->prep_first = first; h->prep_batch = batch; h->prep_neg_ratio = neg_ratio;
    return CRB_OK;
}

// phase 1.  u: local user rows, i/j: global item ids (DEVICE or HOST); or u == NULL: sample rows [first, first+batch) of this rank's
// epoch (crb_set_history holds the rank's own users with GLOBAL item ids).
extern "C" int crb_shard_step_compute(crb_handle* h, const crb_table* P, const crb_shard* shard, const crb_opt* opt, const int32_t* u,
                                      const int32_t* i, const int32_t* j, uint64_t seed, uint32_t epoch, int64_t first, int32_t neg_ratio,
                                      int64_t batch, float reg, double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && P && shard, "null argument");
    CRB_CHECK_ARG(batch > 0 && (int64_t)shard->n_ranks * batch < 0x20000000LL, "batch");
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    if ((rc = crb_table_check(P, opt_kind, "P"))) return rc;
    ShardStepArgs a;
    if ((rc = shard_to_dev(shard, &a.sh))) return rc;
    CRB_CHECK_ARG(shard->q[shard->rank].dim == P->dim, "P.dim != Q.dim");
    CRB_CUDA(cudaSetDevice(h->device));
    if ((rc = crb_ws_reserve(h, batch, P->dim, 4, s))) return rc;
    if ((rc = crb_meta_reserve(h, 0, P->rows, s))) return rc;
    const bool prepared = !u && h->prep_valid && h->prep_seed == seed && h->prep_epoch == epoch && h->prep_first == first &&
                          h->prep_batch == batch && h->prep_neg_ratio == neg_ratio;
    struct Restore { crb_handle* h; ~Restore() { if (h->alt_active) crb_alt_swap(h); } } restore{h};
    if (prepared) {
        h->prep_valid = 0;
        CRB_CUDA(cudaStreamWaitEvent(s, h->ev_prep[1], 0));
        crb_alt_swap(h);   // K3 / K4 / K5 below read the prepared copy
    } else if ((rc = crb_zero_step_counters(h, s))) return rc;
    const int32_t *du = u, *di = i, *dj = j;
    if (prepared) {
        du = h->idx[0]; di = h->idx[1]; dj = h->idx[2];
    } else if (!u) {
        if ((rc = crb_launch_sample_pairwise(h, seed, epoch, first, batch, neg_ratio, h->idx[0], h->idx[1], h->idx[2], nullptr, false, s))) return rc;
        du = h->idx[0]; di = h->idx[1]; dj = h->idx[2];
    } else {
        CRB_CHECK_ARG(i && j, "null index feed");
        if (!crb_is_device_ptr(u)) { CRB_CUDA(cudaMemcpyAsync(h->idx[0], u, 4 * batch, cudaMemcpyHostToDevice, s)); du = h->idx[0]; }
        if (!crb_is_device_ptr(i)) { CRB_CUDA(cudaMemcpyAsync(h->idx[1], i, 4 * batch, cudaMemcpyHostToDevice, s)); di = h->idx[1]; }
        if (!crb_is_device_ptr(j)) { CRB_CUDA(cudaMemcpyAsync(h->idx[2], j, 4 * batch, cudaMemcpyHostToDevice, s)); dj = h->idx[2]; }
    }
    const int32_t* idx[3] = {du, nullptr, nullptr};
    const int role_table[3] = {0, 0, 0};
    if (!prepared) {
        if ((rc = crb_count_rows(h, batch, 1, idx, role_table, s))) return rc;
        if ((rc = crb_launch_assign(h, batch, 1, idx, role_table, s))) return rc;
    }
    a.P = crb_to_dev(P); a.metaU = h->meta[0]; a.u = du; a.i = di; a.j = dj; a.rk_u = h->rank[0];
    a.batch = batch; a.dim = P->dim; a.reg = reg; a.opt = od; a.dup_grad = h->dup_grad; a.dup_t = h->dup_t; a.block_loss = h->block_loss;
#ifdef SH_DEBUG_SWITCHES
    a.debug = getenv("CRB_SH_DEBUG") ? atoi(getenv("CRB_SH_DEBUG")) : 0;
#else
    a.debug = 0;
#endif
    if ((rc = crb_prof_begin(h, s))) return rc;
    if ((rc = CRB_DIM_DISPATCH(a.dim, launch_shard_t, h, a, opt_kind, s))) return rc;
    if ((rc = crb_prof_end(h, s))) return rc;
    DupArgs d;
    fill_dup(h, &d, a.P, a.P, a.dim, od);
    if ((rc = crb_launch_dup_pipeline(h, d, opt_kind, s))) return rc;
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    if ((rc = crb_launch_loss_final(h, ld, s))) return rc;
    if (loss_out && !crb_is_device_ptr(loss_out)) {
        CRB_CUDA(cudaMemcpyAsync(loss_out, h->loss_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
        CRB_CUDA(cudaStreamSynchronize(s));
    }
    return CRB_OK;
}

// phase 2 (after the cross-rank barrier): reduce + apply this rank's inbox, then empty it.
extern "C" int crb_shard_apply_inbox(crb_handle* h, const crb_shard* shard, const crb_opt* opt, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && shard, "null argument");
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    ShardDev sd;
    if ((rc = shard_to_dev(shard, &sd))) return rc;
    const crb_table* Q = &shard->q[shard->rank];
    if ((rc = crb_table_check(Q, opt_kind, "Q shard"))) return rc;
    CRB_CUDA(cudaSetDevice(h->device));
    const int64_t cap = shard->inbox_cap;
    if ((rc = crb_ws_reserve(h, cap, Q->dim, 4, s))) return rc;
    if ((rc = crb_meta_reserve(h, 1, Q->rows, s))) return rc;
    if ((rc = crb_zero_step_counters(h, s))) return rc;
    const int r = shard->rank;
    const int32_t* idx[3] = {sd.inbox_row[r], nullptr, nullptr};
    const int role_table[3] = {1, 0, 0};
    if ((rc = crb_count_rows(h, cap, 1, idx, role_table, s, sd.inbox_cnt[r]))) return rc;
    if ((rc = crb_launch_assign(h, cap, 1, idx, role_table, s, sd.inbox_cnt[r], /*every_row=*/true))) return rc;
    InboxArgs a;
    a.meta = h->meta[1]; a.row = sd.inbox_row[r]; a.key = sd.inbox_key[r]; a.cnt = sd.inbox_cnt[r];
    a.rk = h->rank[0]; a.cap = cap; a.dup_src = h->dup_src; a.dup_t = h->dup_t;
    {
        int64_t blocks = (cap + 255) / 256, capb = (int64_t)h->sm_count * 16;
        inbox_index_kernel<<<(int)(blocks < capb ? blocks : capb), 256, 0, s>>>(a);
        h->launches++;
        CRB_CUDA(cudaGetLastError());
    }
    DupArgs d;
    fill_dup(h, &d, sd.q[r], sd.q[r], Q->dim, od);
    d.dup_src = h->dup_src;
    d.src_grad = sd.inbox_grad[r];
    if ((rc = crb_launch_dup_pipeline(h, d, opt_kind, s))) return rc;
    // CRB_ADAM_TF1: bring every row of the shard to this step (rows not touched now take their decay-only step), so that the next
    // step's readers -- local or over NVLink -- never need a row's `last` (a dependent 4-byte peer load cost 2.3 ms per step)
    if ((rc = crb_adam_flush(h, Q, opt, stream))) return rc;
    // overflow flag is sticky until read by crb_shard_inbox_overflow; the entry counter is reset for the next step
    CRB_CUDA(cudaMemsetAsync(sd.inbox_cnt[r], 0, sizeof(unsigned int), s));
    CRB_CUDA(cudaMemsetAsync(sd.inbox_row[r], 0xFF, sizeof(int32_t) * (size_t)cap, s));   // every slot is a hole until written
    return CRB_OK;
}

extern "C" int crb_shard_inbox_overflow(crb_handle* h, const crb_shard* shard, int32_t* overflowed, void* stream) {
    CRB_CHECK_ARG(h && shard && overflowed, "null argument");
    unsigned int v = 0;
    CRB_CUDA(cudaMemcpyAsync(&v, shard->inbox_cnt[shard->rank] + 1, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CRB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *overflowed = (int32_t)v;
    return CRB_OK;
}

// ------------------------------------------------------------------------------------------------ peer memory plumbing
extern "C" int crb_malloc(crb_handle* h, int64_t bytes, void** out) {
    CRB_CHECK_ARG(h && out && bytes > 0, "bad argument");
    CRB_CUDA(cudaSetDevice(h->device));
    CRB_CUDA(cudaMalloc(out, (size_t)bytes));
    CRB_CUDA(cudaMemset(*out, 0, (size_t)bytes));
    return CRB_OK;
}
extern "C" int crb_free(crb_handle* h, void* p) {
    CRB_CHECK_ARG(h, "null handle");
    CRB_CUDA(cudaSetDevice(h->device));
    CRB_CUDA(cudaFree(p));
    return CRB_OK;
}
extern "C" int crb_ipc_export(crb_handle* h, void* dev_ptr, unsigned char handle64[64]) {
    CRB_CHECK_ARG(h && dev_ptr && handle64, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    cudaIpcMemHandle_t hd;
    CRB_CUDA(cudaSetDevice(h->device));
    CRB_CUDA(cudaIpcGetMemHandle(&hd, dev_ptr));
    memcpy(handle64, &hd, 64);
    return CRB_OK;
}
extern "C" int crb_ipc_open(crb_handle* h, const unsigned char handle64[64], void** out) {
    CRB_CHECK_ARG(h && handle64 && out, "null argument");
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    CRB_CUDA(cudaSetDevice(h->device));
    CRB_CUDA(cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess));
    return CRB_OK;
}
extern "C" int crb_ipc_close(crb_handle* h, void* p) {
    CRB_CHECK_ARG(h, "null handle");
    CRB_CUDA(cudaSetDevice(h->device));
    CRB_CUDA(cudaIpcCloseMemHandle(p));
    return CRB_OK;
}
