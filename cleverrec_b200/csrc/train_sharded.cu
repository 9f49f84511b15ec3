// Multi-GPU BPR step over NVLink peer memory.  One process per GPU.  Users (rows of P, their histories, their
// sampling) are partitioned across ranks; the item table Q is row-sharded (owner = item % G, local row = item / G) and every
// rank maps every shard and every inbox through CUDA IPC, so there is no NCCL all-to-all and no host synchronisation.
//
// What crosses the wire is de-duplicated PER RANK first: an item row that a rank's batch draws more than once is fetched from its
// owner once and its gradients are summed locally before one peer store (2^21 item draws over 2M items: 62 % distinct rows).
//
//   phase 0  (auxiliary stream, overlaps the previous step's phase 2)  sample, count every user / item row's occurrences,
//            give repeated rows gradient slots (K1 / K2 of the single-GPU step) and resolve each triplet's item sources
//   phase 1a item_fetch_kernel   every REPEATED item row: one peer load from the owner into a local staging row
//   phase 1b shard_step_kernel   per triplet: p_u locally, q_i / q_j from the staging buffer (repeated rows) or STRAIGHT FROM THE
//                                OWNER'S HBM (rows that occur once: peer loads requested two iterations ahead), forward/backward,
//                                the user row applied locally (multiplicity machinery of the single-GPU step), the gradient of a
//                                once-occurring item row stored straight into the owner's DIRECT-MAPPED inbox
//                                (slot = (source rank, local row): no reservation, no atomics), a repeated row's into a local slot
//   phase 1c dup_reduce_kernel<SHARD>  user rows: summed and applied as on one GPU; item rows: summed in triplet order and SENT once
//   -- crb_shard_barrier (flag barrier in peer memory, on the stream) --
//   phase 2  inbox_apply_kernel  one pass over the owner's rows: the <= G gradients that arrived for a row are summed in source-rank
//                                order and applied once -- the semantics of one TF step on the union batch; CRB_ADAM_TF1 rows that
//                                received nothing take their decay-only step (remote readers never need a row's `last`)
//   -- crb_shard_barrier --
#include <cstddef>
#include <stdlib.h>

#include "dup.cuh"

struct ShardDev {
    int n_ranks, rank;
    TableDev q[CRB_MAX_RANKS];
    ShardSend send;
};

// ------------------------------------------------------------------------------------------------ pointwise family (MF / GMF)
struct ShardPwArgs {
    TableDev P;
    unsigned long long* metaU;
    unsigned long long* metaI;
    ShardDev sh;
    const int32_t* u;   // local user rows
    const int32_t* i;   // GLOBAL item ids
    const float* y;
    const int32_t* sbi;
    const uint32_t* rk[2];
    const float* stage;
    const float* hvec;  // GMF's h (replicated), NULL for MF
    float* hpart;       // [gridDim.x, dim] per-block partial gradient of h
    int64_t batch;
    int dim;
    int loss_kind;
    float reg;
    OptDev opt;
    float* dup_grad;
    uint32_t* dup_t;
    double* block_loss;
};

struct ShardStepArgs {
    TableDev P;
    unsigned long long* metaU;
    unsigned long long* metaI;   // indexed by GLOBAL item id: this rank's multiplicities
    ShardDev sh;
    const int32_t* u;   // local user rows
    const int32_t* i;   // GLOBAL item ids
    const int32_t* j;
    const int32_t* sbi; // per triplet: first gradient slot (= staging row) of item i / j when the row repeats in this rank's batch, else -1
    const int32_t* sbj;
    const uint32_t* rk[3];
    const float* stage;
    int64_t batch;
    int dim;
    float reg;
    OptDev opt;
    float* dup_grad;
    uint32_t* dup_t;
    double* block_loss;
    int debug;   // experiments only (-DSH_DEBUG_SWITCHES + CRB_SH_DEBUG): 1 = no gradient sends, 2 = item rows read from the local shard, 3 = both
};

// ------------------------------------------------------------------------------------------------ phase 0: item sources
__global__ void __launch_bounds__(256) shard_resolve_kernel(const unsigned long long* __restrict__ metaI, const int32_t* __restrict__ i,
                                                            const int32_t* __restrict__ j, int64_t batch, int32_t* __restrict__ sbi,
                                                            int32_t* __restrict__ sbj) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < batch; t += stride) {
        const unsigned long long mi = metaI[i[t]], mj = metaI[j[t]];
        sbi[t] = (uint32_t)mi > 1u ? (int32_t)(mi >> 32) : -1;
        sbj[t] = (uint32_t)mj > 1u ? (int32_t)(mj >> 32) : -1;
    }
}

// ------------------------------------------------------------------------------------------------ phase 1a: fetch repeated rows
// One lane group per duplicate-row descriptor; item rows (table 1) are copied from the owner's shard to stage[base].  The loop is
// two deep: the next descriptor's row is requested before the current one is stored.
template <int LANES, int VPL>
__global__ void __launch_bounds__(256) item_fetch_kernel(ShardDev sh, const crb_dup_row* __restrict__ dup_rows, const crb_step_ctr* ctr,
                                                         float* __restrict__ stage, int dim) {
    const int gl = threadIdx.x % LANES;
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / LANES;
    const int64_t n = ctr->dup_rows;
    const int G = sh.n_ranks;
    for (int64_t k = group; k < n; k += n_groups) {
        const crb_dup_row d = dup_rows[k];
        if (d.table != 1) continue;
        const float* src = sh.q[d.row % G].w + (int64_t)(d.row / G) * dim;
        float* dst = stage + (int64_t)d.base * dim;
        float4 r[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            if (c < dim) r[v] = ld4(src + c);
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            if (c < dim) st4(dst + c, r[v]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ phase 1b: the fused step
template <int LANES, int VPL>
__device__ __forceinline__ void item_row_load(RowRegs<LANES, VPL>& r, const ShardStepArgs& a, int32_t item, int32_t sb, int dim, int gl) {
    const int G = a.sh.n_ranks;
    const float* src = sb >= 0 ? a.stage + (int64_t)sb * dim : a.sh.q[item % G].w + (int64_t)(item / G) * dim;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c = (gl + LANES * v) * 4;
        r.w[v] = c < dim ? ld4(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int LANES, int VPL, int OPT>
__global__ void __launch_bounds__(256, 3) shard_step_kernel(ShardStepArgs a) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31, gl = lane % LANES, sub = lane / LANES;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double loss_acc = 0.0;
    // Software pipeline.  Once-occurring item rows live on other GPUs (2-3 us away over NVLink), so item rows are requested TWO
    // iterations ahead; the local user row one iteration ahead; indices and item sources three iterations ahead.  Out-of-range
    // iterations clamp to the last triplet (loads only).
    const int64_t stride = n_warps * GPW;
    const int64_t base0 = warp * GPW;
    auto clampt = [&](int64_t b) { const int64_t t_ = b + sub; return t_ < a.batch ? t_ : a.batch - 1; };
    int32_t u1, i1, j1;                  // indices of iteration n+1 (its item rows are in flight, its user row is about to be)
    int32_t u2, i2, j2, si2, sj2;        // indices and item sources of iteration n+2
    int32_t nu, ni, nj;                  // indices of the current iteration
    RowRegs<LANES, VPL> nru, nri, nrj, fri, frj;
    {
        const int64_t t0 = clampt(base0), t1 = clampt(base0 + stride), t2 = clampt(base0 + 2 * stride);
        nu = a.u[t0]; ni = a.i[t0]; nj = a.j[t0];
        u1 = a.u[t1]; i1 = a.i[t1]; j1 = a.j[t1];
        u2 = a.u[t2]; i2 = a.i[t2]; j2 = a.j[t2]; si2 = a.sbi[t2]; sj2 = a.sbj[t2];
        item_row_load<LANES, VPL>(nri, a, ni, a.sbi[t0], a.dim, gl);
        item_row_load<LANES, VPL>(nrj, a, nj, a.sbj[t0], a.dim, gl);
        item_row_load<LANES, VPL>(fri, a, i1, a.sbi[t1], a.dim, gl);
        item_row_load<LANES, VPL>(frj, a, j1, a.sbj[t1], a.dim, gl);
        row_load_w<LANES, VPL>(nru, a.P, nu, a.dim, gl);
    }

    for (int64_t base = base0; base < a.batch; base += stride) {
        const int64_t t = base + sub;
        const bool active = t < a.batch;
        const int64_t tt = active ? t : a.batch - 1;
        const int32_t u = nu, i = ni, j = nj;
        RowRegs<LANES, VPL> ru = nru, ri = nri, rj = nrj;
        // this triplet's own slots (re-read: cheaper than carrying them through the pipeline registers)
        const int32_t sbi = a.sbi[tt], sbj = a.sbj[tt];
        const uint32_t rki = a.rk[1][tt], rkj = a.rk[2][tt];
        // rotate the pipeline: iteration n+1's item rows were requested last time; request n+2's and n+1's user row
        nu = u1; ni = i1; nj = j1;
        nri = fri; nrj = frj;
        row_load_w<LANES, VPL>(nru, a.P, nu, a.dim, gl);
        item_row_load<LANES, VPL>(fri, a, i2, si2, a.dim, gl);
        item_row_load<LANES, VPL>(frj, a, j2, sj2, a.dim, gl);
        u1 = u2; i1 = i2; j1 = j2;
        { const int64_t t3 = clampt(base + 3 * stride); u2 = a.u[t3]; i2 = a.i[t3]; j2 = a.j[t3]; si2 = a.sbi[t3]; sj2 = a.sbj[t3]; }
        // (Fetching this word and `last` one iteration ahead was measured: the three extra live registers spill at the 80-register
        // budget of three CTAs per SM and the kernel went from 1.35 to 1.59 ms at N=2.)
        const unsigned long long mu = a.metaU[u];
        // item rows are always current at a step boundary (the owner brings its whole shard to the step in phase 2), so only the
        // local user row can have missed steps to replay
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
        if (!active) continue;
        emit_row<LANES, VPL, OPT>(ru, gu, a.P, a.metaU, u, mu, a.rk[0][t], (uint32_t)t, 0u, a.dim, gl, a.opt, a.dup_grad, a.dup_t);
        // item gradients: a once-occurring row goes straight into its owner's inbox; a repeated row's into its local slot (the slot
        // key orders the local sum by triplet: identical from run to run)
#pragma unroll
        for (int role = 1; role <= 2; ++role) {
            const int32_t item = role == 1 ? i : j, sb = role == 1 ? sbi : sbj;
            const float4* gv = role == 1 ? gi : gj;
            if (sb < 0) {
#ifdef SH_DEBUG_SWITCHES
                if (a.debug & 1) continue;
#endif
                shard_send<LANES, VPL>(a.sh.send, item, gv, a.dim, gl);
                if (gl == 0) a.metaI[item] = 0ULL;
            } else {
                const uint32_t slot = (uint32_t)sb + (role == 1 ? rki : rkj);
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    const int c = (gl + LANES * v) * 4;
                    if (c < a.dim) st4(a.dup_grad + (int64_t)slot * a.dim + c, gv[v]);
                }
                if (gl == 0) a.dup_t[slot] = ((uint32_t)t << 2) | (uint32_t)role;
            }
        }
    }
    block_loss_store(loss_acc, a.block_loss);
}

// ------------------------------------------------------------------------------------------------ phase 1b, asynchronous staging
// The same step with the gathers taken off the registers (dim <= 128).  shard_step_kernel holds every prefetched row in registers
// (five rows ahead of their use at 80 registers per thread) and still stalls twice per triplet on the user row's multiplicity word
// and optimizer slots: 3.6 TB/s.  Here every warp owns a private slice of shared memory:
//   * HEADERS of the next 16 iterations (ids, item sources, the user row's multiplicity / slot base / last step, occurrence ranks),
//     loaded lane-per-iteration in three spaced phases (ids -> multiplicity words -> shared memory) half a block ahead, so nothing on
//     the index side is ever waited for;
//   * three ROW STAGES: the rows of iterations n+1 and n+2 are in flight (16-byte cp.async, global / peer memory -> shared, no
//     registers) while iteration n computes.  A lane copies exactly the 16-byte chunks it later reads, so the only synchronisation is
//     its own cp.async.wait_group -- no barrier, no mbarrier, no cross-warp hand-off.
// Arithmetic, order and device functions are those of shard_step_kernel: bit-identical tables.
#define SA_NBUF 3
#define SA_HB 16
#define SA_NROW 5   // p_u, q_i, q_j, s1_u, s2_u
struct SaHdr {
    int32_t u, i, j, sbi, sbj;   // u < 0: no triplet (tail)
    uint32_t cnt, base;          // the user row's multiplicity word
    int32_t last;
    uint32_t rku, rki, rkj, pad;
};
static_assert(sizeof(SaHdr) == 48, "header size");

__device__ __forceinline__ void sa_copy16(void* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int LANES, int OPT>
__global__ void __launch_bounds__(256, 3) shard_step_async_kernel(ShardStepArgs a) {
    constexpr int GPW = 32 / LANES, VPL = 1;
    extern __shared__ __align__(16) unsigned char sa_sm[];
    const int lane = threadIdx.x & 31, gl = lane % LANES, sub = lane / LANES;
    const uint32_t rb = (uint32_t)a.dim * 4u;
    const size_t hdr_bytes = 2 * SA_HB * GPW * sizeof(SaHdr);
    const size_t warp_bytes = hdr_bytes + (size_t)SA_NBUF * GPW * SA_NROW * rb;
    unsigned char* wbase = sa_sm + (threadIdx.x >> 5) * warp_bytes;
    SaHdr* hdr = reinterpret_cast<SaHdr*>(wbase);
    unsigned char* rows = wbase + hdr_bytes;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_groups = (a.batch + GPW - 1) / GPW;
    const int64_t n_it = n_groups > gwarp ? (n_groups - gwarp + n_warps - 1) / n_warps : 0;
    const int G = a.sh.n_ranks;
    const bool my_chunk = gl * 4 < a.dim;
    double loss_acc = 0.0;

    // ---- header pipeline: lane l < SA_HB owns iteration (block * SA_HB + l) of the block being loaded
    int32_t hu[GPW], hi[GPW], hj[GPW], hsi[GPW], hsj[GPW], hl[GPW];
    uint32_t hru[GPW], hri[GPW], hrj[GPW];
    unsigned long long hm[GPW];
    auto phase_ids = [&](int64_t blk) {
        const int64_t m = blk * SA_HB + lane;
#pragma unroll
        for (int k = 0; k < GPW; ++k) {
            const int64_t t = (gwarp + m * n_warps) * GPW + k;
            hu[k] = -1;
            if (lane < SA_HB && m < n_it && t < a.batch) {
                hu[k] = a.u[t]; hi[k] = a.i[t]; hj[k] = a.j[t]; hsi[k] = a.sbi[t]; hsj[k] = a.sbj[t];
                hru[k] = a.rk[0][t]; hri[k] = a.rk[1][t]; hrj[k] = a.rk[2][t];
            }
        }
    };
    auto phase_meta = [&]() {
#pragma unroll
        for (int k = 0; k < GPW; ++k) {
            hm[k] = 0ULL; hl[k] = 0;
            if (hu[k] >= 0) {
                hm[k] = a.metaU[hu[k]];
                hl[k] = OptTraits<OPT>::replay ? a.P.last[hu[k]] : 0;
            }
        }
    };
    auto phase_store = [&](int64_t blk) {
        if (lane < SA_HB) {
#pragma unroll
            for (int k = 0; k < GPW; ++k) {
                SaHdr H;
                H.u = hu[k]; H.i = hi[k]; H.j = hj[k]; H.sbi = hsi[k]; H.sbj = hsj[k];
                H.cnt = (uint32_t)hm[k]; H.base = (uint32_t)(hm[k] >> 32); H.last = hl[k];
                H.rku = hru[k]; H.rki = hri[k]; H.rkj = hrj[k]; H.pad = 0u;
                hdr[((blk & 1) * SA_HB + lane) * GPW + k] = H;
            }
        }
        __syncwarp();
    };
    auto header = [&](int64_t m) -> const SaHdr& { return hdr[(((m / SA_HB) & 1) * SA_HB + (m % SA_HB)) * GPW + sub]; };
    // ---- row stage m: this lane's 16-byte chunk of every row iteration m needs
    auto issue = [&](int64_t m) {
        if (m < n_it) {
            const SaHdr& H = header(m);
            if (H.u >= 0 && my_chunk) {
                unsigned char* st = rows + ((size_t)(m % SA_NBUF) * GPW + sub) * SA_NROW * rb + (size_t)gl * 16;
                const bool su = H.cnt == 1u || replay_pending<OPT>(H.last, a.opt);
                sa_copy16(st, a.P.w + (int64_t)H.u * a.dim + gl * 4);
                const float* qi = H.sbi >= 0 ? a.stage + (int64_t)H.sbi * a.dim : a.sh.q[H.i % G].w + (int64_t)(H.i / G) * a.dim;
                const float* qj = H.sbj >= 0 ? a.stage + (int64_t)H.sbj * a.dim : a.sh.q[H.j % G].w + (int64_t)(H.j / G) * a.dim;
                sa_copy16(st + rb, qi + gl * 4);
                sa_copy16(st + 2 * rb, qj + gl * 4);
                if (OptTraits<OPT>::has_s1 && su) sa_copy16(st + 3 * rb, a.P.s1 + (int64_t)H.u * a.dim + gl * 4);
                if (OptTraits<OPT>::has_s2 && su) sa_copy16(st + 4 * rb, a.P.s2 + (int64_t)H.u * a.dim + gl * 4);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");   // one group per iteration, empty or not: the wait below counts groups
    };

    phase_ids(0); phase_meta(); phase_store(0);
    issue(0);
    issue(1);
    for (int64_t it = 0; it < n_it; ++it) {
        const int ph = (int)(it % SA_HB);
        const int64_t blk = it / SA_HB;
        if (ph == 0) phase_ids(blk + 1);
        else if (ph == 5) phase_meta();
        else if (ph == 10) phase_store(blk + 1);
        issue(it + 2);
        asm volatile("cp.async.wait_group 2;" ::: "memory");   // the copies of iteration `it` (issued two iterations ago) have landed
        const SaHdr H = header(it);
        const bool active = H.u >= 0;
        const int32_t u = active ? H.u : 0, i = H.i, j = H.j;
        const unsigned long long mu = ((unsigned long long)H.base << 32) | H.cnt;
        RowRegs<LANES, VPL> ru, ri, rj;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        ru.w[0] = ri.w[0] = rj.w[0] = z4;
        ru.last = H.last;
        const bool su = active && (H.cnt == 1u || replay_pending<OPT>(H.last, a.opt));
        if (active && my_chunk) {
            const unsigned char* st = rows + ((size_t)(it % SA_NBUF) * GPW + sub) * SA_NROW * rb + (size_t)gl * 16;
            ru.w[0] = *reinterpret_cast<const float4*>(st);
            ri.w[0] = *reinterpret_cast<const float4*>(st + rb);
            rj.w[0] = *reinterpret_cast<const float4*>(st + 2 * rb);
            if (OptTraits<OPT>::has_s1 && su) ru.s1[0] = *reinterpret_cast<const float4*>(st + 3 * rb);
            if (OptTraits<OPT>::has_s2 && su) ru.s2[0] = *reinterpret_cast<const float4*>(st + 4 * rb);
        } else if (su) {   // padding lanes (dim not a multiple of 4 * LANES): see row_load_state
            const float pad1 = OptTraits<OPT>::has_s2 ? 0.f : 1.f;
            if (OptTraits<OPT>::has_s1) ru.s1[0] = make_float4(pad1, pad1, pad1, pad1);
            if (OptTraits<OPT>::has_s2) ru.s2[0] = make_float4(1.f, 1.f, 1.f, 1.f);
        }
        if (active && replay_pending<OPT>(ru.last, a.opt)) row_replay<LANES, VPL, OPT>(ru, a.opt, a.opt.step);
        float x = 0.f, sq = 0.f;
        {
            const float4 p = ru.w[0], qi = ri.w[0], qj = rj.w[0];
            const float4 dq = make_float4(qi.x - qj.x, qi.y - qj.y, qi.z - qj.z, qi.w - qj.w);
            x += dot4(p, dq);
            sq += dot4(p, p) + dot4(qi, qi) + dot4(qj, qj);
        }
        x = group_sum<LANES>(x);
        sq = group_sum<LANES>(sq);
        const float g = -sigmoid_f(-x);
        if (active && gl == 0) loss_acc += (double)(softplus_neg(x) + a.reg * 0.5f * sq);
        float4 gu[VPL], gi[VPL], gj[VPL];
        {
            const float4 p = ru.w[0], qi = ri.w[0], qj = rj.w[0];
            gu[0] = make_float4(fmaf(g, qi.x - qj.x, a.reg * p.x), fmaf(g, qi.y - qj.y, a.reg * p.y), fmaf(g, qi.z - qj.z, a.reg * p.z),
                                fmaf(g, qi.w - qj.w, a.reg * p.w));
            gi[0] = make_float4(fmaf(g, p.x, a.reg * qi.x), fmaf(g, p.y, a.reg * qi.y), fmaf(g, p.z, a.reg * qi.z), fmaf(g, p.w, a.reg * qi.w));
            gj[0] = make_float4(fmaf(-g, p.x, a.reg * qj.x), fmaf(-g, p.y, a.reg * qj.y), fmaf(-g, p.z, a.reg * qj.z), fmaf(-g, p.w, a.reg * qj.w));
        }
        if (!active) continue;
        const uint32_t t = (uint32_t)((gwarp + it * n_warps) * GPW + sub);
        emit_row<LANES, VPL, OPT>(ru, gu, a.P, a.metaU, u, mu, H.rku, t, 0u, a.dim, gl, a.opt, a.dup_grad, a.dup_t);
#pragma unroll
        for (int role = 1; role <= 2; ++role) {
            const int32_t item = role == 1 ? i : j, sb = role == 1 ? H.sbi : H.sbj;
            const float4* gv = role == 1 ? gi : gj;
            if (sb < 0) {
                shard_send<LANES, VPL>(a.sh.send, item, gv, a.dim, gl);
                if (gl == 0) a.metaI[item] = 0ULL;
            } else {
                const uint32_t slot = (uint32_t)sb + (role == 1 ? H.rki : H.rkj);
                if (my_chunk) st4(a.dup_grad + (int64_t)slot * a.dim + gl * 4, gv[0]);
                if (gl == 0) a.dup_t[slot] = (t << 2) | (uint32_t)role;
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    block_loss_store(loss_acc, a.block_loss);
}

// MEASURED (S-large, B = 2^20 per GPU, Adam tf1; whole step, ms): N=2 2.71 async vs 2.79 register-staged; N=4 2.61 vs 2.71; N=8 2.70 vs
// 2.65 (step kernel 1.31 vs 1.22 ms: with 7/8 of the direct reads remote, two iterations of look-ahead no longer cover the NVLink
// round trip, and a deeper ring would cost the third CTA per SM).  Hence: asynchronous staging up to four ranks.
// CRB_SHARD_ASYNC=0 / 1 forces one or the other (A/B).
static bool shard_async_enabled(int dim, int n_ranks) {
    if (dim > 128) return false;
    const char* e = getenv("CRB_SHARD_ASYNC");
    if (e) return atoi(e) != 0;
    return n_ranks <= 4;
}

// ------------------------------------------------------------------------------------------------ phase 2: the owner's pass
struct InboxApplyArgs {
    TableDev Q;           // this rank's shard
    int64_t rows;         // its rows
    int64_t rows_cap;
    int n_ranks;
    int dim;
    uint32_t stamp;
    const float* grad;    // [n_ranks][rows_cap][dim]
    const uint32_t* stamps;
    OptDev opt;
};

// GMAX = 2 / 4 / 8 >= n_ranks bounds the gradient registers (GMAX float4 per VPL): with two ranks the kernel keeps 8 of them instead of
// 32 and a fourth CTA fits on the SM.  The row's weights and optimizer slots are requested TOGETHER with its stamps (every touched row
// needs them, and so does every CRB_ADAM_TF1 row that has ever been updated), so a row costs two dependent round trips -- stamps +
// row, then the arrived gradients -- instead of three.
template <int LANES, int VPL, int OPT, int GMAX>
__global__ void __launch_bounds__(256) inbox_apply_kernel(InboxApplyArgs a) {
    const int lane = threadIdx.x & 31, gl = lane % LANES;
    const uint32_t gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << (lane - gl));
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / LANES;
    // whole lane groups leave together (a row index is per lane group; the votes below are group-wide)
    for (int64_t row = group; row < a.rows; row += n_groups) {
        // lane s of the group looks at source rank s (LANES >= 8 >= n_ranks)
        const uint32_t st = gl < a.n_ranks ? a.stamps[(int64_t)gl * a.rows_cap + row] : 0u;
        RowRegs<LANES, VPL> r;
        r.last = OptTraits<OPT>::replay ? a.Q.last[row] : 0;
        constexpr bool EARLY = OptTraits<OPT>::replay;   // the other optimizers leave untouched rows alone: no speculative reads for them
        if (EARLY) {
            row_load_w<LANES, VPL>(r, a.Q, row, a.dim, gl);
            row_load_state<LANES, VPL, OPT>(r, a.Q, row, a.dim, gl);
        }
        uint32_t present = __ballot_sync(gmask, gl < a.n_ranks && st == a.stamp);
        present = LANES == 32 ? present : (present >> (lane - gl)) & ((1u << LANES) - 1u);
        if (!EARLY && present != 0u) {
            row_load_w<LANES, VPL>(r, a.Q, row, a.dim, gl);
            row_load_state<LANES, VPL, OPT>(r, a.Q, row, a.dim, gl);
        }
        if (present == 0u) {
            // nothing arrived: CRB_ADAM_TF1 rows decay (tf.train.AdamOptimizer moves every row every step); others are untouched
            if (OptTraits<OPT>::replay && r.last != 0 && r.last < a.opt.step) {
                row_replay<LANES, VPL, OPT>(r, a.opt, a.opt.step + 1);
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    const int c = (gl + LANES * v) * 4;
                    if (c >= a.dim) continue;
                    st4(a.Q.w + row * a.dim + c, r.w[v]);
                    st4(a.Q.s1 + row * a.dim + c, r.s1[v]);
                    st4(a.Q.s2 + row * a.dim + c, r.s2[v]);
                }
                if (gl == 0) a.Q.last[row] = a.opt.step;
            }
            continue;
        }
        // all arrived gradients are requested together (one round trip), then added in source-rank order
        float4 acc[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 g[GMAX][VPL];
#pragma unroll
        for (int q = 0; q < GMAX; ++q) {
            const bool on = (present >> q) & 1u;
            const float* src = a.grad + ((int64_t)q * a.rows_cap + row) * a.dim;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c = (gl + LANES * v) * 4;
                g[q][v] = (on && c < a.dim) ? ld4(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int q = 0; q < GMAX; ++q) {
            if (!((present >> q) & 1u)) continue;
#pragma unroll
            for (int v = 0; v < VPL; ++v) { acc[v].x += g[q][v].x; acc[v].y += g[q][v].y; acc[v].z += g[q][v].z; acc[v].w += g[q][v].w; }
        }
        row_replay<LANES, VPL, OPT>(r, a.opt, a.opt.step);
        row_apply_store<LANES, VPL, OPT>(r, acc, a.Q, row, a.dim, gl, a.opt);
    }
}

// ------------------------------------------------------------------------------------------------ flag barrier in peer memory
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct BarrierArgs {
    int n_ranks, rank;
    uint32_t* flags[CRB_MAX_RANKS];
    uint32_t ticket;
    unsigned long long timeout_ns;
};

__global__ void shard_barrier_kernel(BarrierArgs a) {
    const int r = threadIdx.x;
    if (r >= a.n_ranks) return;
    // everything this stream did before the barrier (peer stores included) is complete: the kernel boundary orders it; the fence
    // makes it visible system-wide before the flag that announces it
    __threadfence_system();
    volatile uint32_t* theirs = a.flags[r] + a.rank;
    *theirs = a.ticket;
    __threadfence_system();
    volatile uint32_t* mine = a.flags[a.rank] + r;
    const unsigned long long t0 = global_ns();
    while ((int32_t)(*mine - a.ticket) < 0) {
        if (global_ns() - t0 > a.timeout_ns) { a.flags[a.rank][CRB_SHARD_ERR] = 1u; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

// ------------------------------------------------------------------------------------------------ host side
static int shard_to_dev(const crb_shard* sh, ShardDev* d, uint32_t stamp) {
    CRB_CHECK_ARG(sh && sh->n_ranks >= 1 && sh->n_ranks <= CRB_MAX_RANKS && sh->rank >= 0 && sh->rank < sh->n_ranks, "shard descriptor");
    CRB_CHECK_ARG(sh->rows_cap > 0, "shard descriptor: rows_cap");
    d->n_ranks = sh->n_ranks; d->rank = sh->rank;
    d->send.n_ranks = sh->n_ranks; d->send.rank = sh->rank; d->send.rows_cap = sh->rows_cap; d->send.stamp = stamp;
    for (int r = 0; r < CRB_MAX_RANKS; ++r) { d->send.grad[r] = nullptr; d->send.stamps[r] = nullptr; d->q[r] = TableDev{nullptr, nullptr, nullptr, nullptr}; }
    for (int r = 0; r < sh->n_ranks; ++r) {
        CRB_CHECK_ARG(sh->q[r].w && sh->inbox_grad[r] && sh->inbox_stamp[r] && sh->flags[r], "shard descriptor: null peer pointer");
        CRB_CHECK_ARG(sh->q[r].rows <= sh->rows_cap, "shard descriptor: rows_cap smaller than a shard");
        d->q[r] = crb_to_dev(&sh->q[r]);
        d->send.grad[r] = sh->inbox_grad[r]; d->send.stamps[r] = sh->inbox_stamp[r];
    }
    return CRB_OK;
}

static int64_t shard_items(const crb_shard* sh) {
    int64_t n = 0;
    for (int r = 0; r < sh->n_ranks; ++r) n += sh->q[r].rows;
    return n;
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
    int rc;
    // phase 1a: repeated item rows -> staging buffer (grid: every SM full, the loop covers the descriptor list)
    if ((rc = crb_prof_begin(h, s, 1))) return rc;
    item_fetch_kernel<LANES, VPL><<<h->sm_count * 8, 256, 0, s>>>(a.sh, h->dup_rows, h->ctr, h->stage, a.dim);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    if ((rc = crb_prof_end(h, s, 1))) return rc;
    if ((rc = crb_prof_begin(h, s))) return rc;
    const bool use_async = VPL == 1 && shard_async_enabled(a.dim, a.sh.n_ranks);
    const size_t sa_smem = 8 * (2 * SA_HB * (32 / LANES) * sizeof(SaHdr) + (size_t)SA_NBUF * (32 / LANES) * SA_NROW * a.dim * 4);
#define CRB_SH_CASE(O)                                                                                                          \
    case O: {                                                                                                                   \
        if (use_async) {                                                                                                        \
            CRB_CUDA(cudaFuncSetAttribute(shard_step_async_kernel<LANES, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sa_smem)); \
            int occ = 0;                                                                                                        \
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, shard_step_async_kernel<LANES, O>, 256, sa_smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; } \
            int64_t grid = (int64_t)h->sm_count * occ;                                                                          \
            const int64_t need = (a.batch + gpb - 1) / gpb;                                                                     \
            if (grid > need) grid = need;                                                                                       \
            if (grid > h->loss_blocks) grid = h->loss_blocks;                                                                   \
            h->step_grid = (int)grid;                                                                                           \
            shard_step_async_kernel<LANES, O><<<(int)grid, 256, sa_smem, s>>>(a);                                               \
        } else {                                                                                                                \
            const int grid = one_wave(h, shard_step_kernel<LANES, VPL, O>, a.batch, gpb);                                       \
            h->step_grid = grid;                                                                                                \
            shard_step_kernel<LANES, VPL, O><<<grid, 256, 0, s>>>(a);                                                           \
        }                                                                                                                       \
        break;                                                                                                                  \
    }
    switch (opt_kind) { CRB_SH_CASE(OPT_SGD) CRB_SH_CASE(OPT_ADAGRAD) CRB_SH_CASE(OPT_ADAM_LAZY) CRB_SH_CASE(OPT_ADAM_TF1) }
#undef CRB_SH_CASE
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    if ((rc = crb_prof_end(h, s))) return rc;
    // phase 1c: duplicate rows -- users applied locally, items summed and sent
    DupArgs d;
    d.tab[0] = a.P; d.tab[1] = a.P;   // table 1 rows never reach an apply in SHARD kernels
    d.meta[0] = a.metaU; d.meta[1] = a.metaI;
    d.dim = a.dim; d.opt = a.opt;
    d.dup_rows = h->dup_rows; d.work = h->work; d.multi = h->multi;
    d.dup_grad = h->dup_grad; d.dup_t = h->dup_t; d.partial = h->partial; d.ctr = h->ctr;
    d.send = a.sh.send;
    const int grid = h->sm_count * 4;
    if ((rc = crb_prof_begin(h, s, 2))) return rc;
#define CRB_SHDUP_CASE(O)                                                        \
    case O:                                                                      \
        dup_reduce_kernel<LANES, VPL, O, true><<<grid, 256, 0, s>>>(d);          \
        dup_final_kernel<LANES, VPL, O, true><<<h->sm_count, 256, 0, s>>>(d);    \
        break;
    switch (opt_kind) { CRB_SHDUP_CASE(OPT_SGD) CRB_SHDUP_CASE(OPT_ADAGRAD) CRB_SHDUP_CASE(OPT_ADAM_LAZY) CRB_SHDUP_CASE(OPT_ADAM_TF1) }
#undef CRB_SHDUP_CASE
    h->launches += 2;
    CRB_CUDA(cudaGetLastError());
    return crb_prof_end(h, s, 2);
}

template <int LANES, int VPL>
static int launch_inbox_t(crb_handle* h, const InboxApplyArgs& a, int opt_kind, cudaStream_t s) {
    const int gpb = 256 / LANES;
    int64_t grid = (a.rows + gpb - 1) / gpb;
    const int64_t cap = (int64_t)h->sm_count * 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    int rc;
    if ((rc = crb_prof_begin(h, s, 3))) return rc;
#define CRB_IN_CASE(O)                                                                               \
    case O:                                                                                          \
        if (a.n_ranks <= 2) inbox_apply_kernel<LANES, VPL, O, 2><<<(int)grid, 256, 0, s>>>(a);       \
        else if (a.n_ranks <= 4) inbox_apply_kernel<LANES, VPL, O, 4><<<(int)grid, 256, 0, s>>>(a);  \
        else inbox_apply_kernel<LANES, VPL, O, 8><<<(int)grid, 256, 0, s>>>(a);                      \
        break;
    switch (opt_kind) { CRB_IN_CASE(OPT_SGD) CRB_IN_CASE(OPT_ADAGRAD) CRB_IN_CASE(OPT_ADAM_LAZY) CRB_IN_CASE(OPT_ADAM_TF1) }
#undef CRB_IN_CASE
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return crb_prof_end(h, s, 3);
}

#define CRB_DIM_DISPATCH(dim, FN, ...)                                   \
    ((dim) <= 32 ? FN<8, 1>(__VA_ARGS__)                                 \
     : (dim) <= 64 ? FN<16, 1>(__VA_ARGS__)                              \
     : (dim) <= 128 ? FN<32, 1>(__VA_ARGS__)                             \
     : (dim) <= 256 ? FN<32, 2>(__VA_ARGS__)                             \
                    : FN<32, 4>(__VA_ARGS__))

static int stage_reserve(crb_handle* h, int32_t dim, cudaStream_t s) {
    const int64_t rows = 3 * h->cap_batch;
    if (h->stage && h->stage_rows >= rows && h->stage_dim >= dim) return CRB_OK;
    CRB_CUDA(cudaStreamSynchronize(s));
    cudaFree(h->stage);
    h->stage = nullptr; h->stage_rows = 0; h->stage_dim = 0;
    CRB_CUDA(cudaMalloc(&h->stage, sizeof(float) * (size_t)rows * dim));
    h->stage_rows = rows; h->stage_dim = dim;
    return CRB_OK;
}

// everything of a step that depends only on its indices: occurrence counts (unless the sampler counted), slot assignment, item sources
static int shard_index_work(crb_handle* h, const int32_t* u, const int32_t* i, const int32_t* j, int64_t batch, bool counted, cudaStream_t s) {
    const int32_t* idx[3] = {u, i, j};
    const int role_table[3] = {0, 1, 1};
    int rc;
    if (!counted && (rc = crb_count_rows(h, batch, 3, idx, role_table, s))) return rc;
    if ((rc = crb_launch_assign(h, batch, 3, idx, role_table, s))) return rc;
    int64_t blocks = (batch + 255) / 256;
    const int64_t capb = (int64_t)h->sm_count * 16;
    shard_resolve_kernel<<<(int)(blocks < capb ? blocks : capb), 256, 0, s>>>(h->meta[1], i, j, batch, h->sb[0], h->sb[1]);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

extern "C" int crb_shard_step_prepare(crb_handle* h, const crb_table* P, uint64_t seed, uint32_t epoch, int64_t first, int32_t neg_ratio,
                                      int64_t batch, int64_t n_items, const int32_t* feed_u, const int32_t* feed_i,
                                      const int32_t* feed_j, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && P, "null argument");
    CRB_CHECK_ARG(batch > 0 && batch < 0x20000000LL && n_items > 0, "batch / n_items");
    CRB_CUDA(cudaSetDevice(h->device));
    int rc;
    if ((rc = crb_ws_reserve(h, batch, P->dim, 4, s))) return rc;
    if ((rc = crb_meta_reserve(h, 0, P->rows, s))) return rc;
    if ((rc = crb_meta_reserve(h, 1, n_items, s))) return rc;
    if ((rc = crb_alt_reserve(h, s))) return rc;
    if ((rc = stage_reserve(h, P->dim, s))) return rc;
    cudaStream_t ps = h->aux_stream;
    // the alternate copy is free once the last compute that used it has finished; a freshly zeroed meta needs the caller's stream
    CRB_CUDA(cudaEventRecord(h->ev_entry, s));
    CRB_CUDA(cudaStreamWaitEvent(ps, h->ev_entry, 0));
    crb_alt_swap(h);
    struct Restore { crb_handle* h; ~Restore() { if (h->alt_active) crb_alt_swap(h); } } restore{h};
    if ((rc = crb_zero_step_counters(h, ps))) return rc;
    bool counted = false;
    if (feed_u) {
        CRB_CHECK_ARG(feed_i && feed_j, "null index feed");
        const int32_t* src[3] = {feed_u, feed_i, feed_j};
        for (int q = 0; q < 3; ++q)
            CRB_CUDA(cudaMemcpyAsync(h->idx[q], src[q], sizeof(int32_t) * batch, crb_is_device_ptr(src[q]) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ps));
    } else {
        if ((rc = crb_launch_sample_pairwise(h, seed, epoch, first, batch, neg_ratio, h->idx[0], h->idx[1], h->idx[2], nullptr, true, ps))) return rc;
        counted = true;
    }
    if ((rc = shard_index_work(h, h->idx[0], h->idx[1], h->idx[2], batch, counted, ps))) return rc;
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
    CRB_CHECK_ARG(batch > 0 && batch < 0x20000000LL, "batch");
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    if ((rc = crb_table_check(P, opt_kind, "P"))) return rc;
    ShardStepArgs a;
    if ((rc = shard_to_dev(shard, &a.sh, (uint32_t)od.step))) return rc;
    CRB_CHECK_ARG(shard->q[shard->rank].dim == P->dim, "P.dim != Q.dim");
    const int64_t n_items = shard_items(shard);
    CRB_CUDA(cudaSetDevice(h->device));
    if ((rc = crb_ws_reserve(h, batch, P->dim, 4, s))) return rc;
    if ((rc = crb_meta_reserve(h, 0, P->rows, s))) return rc;
    if ((rc = crb_meta_reserve(h, 1, n_items, s))) return rc;
    if ((rc = stage_reserve(h, P->dim, s))) return rc;
    const bool prepared = !u && h->prep_valid && h->prep_seed == seed && h->prep_epoch == epoch && h->prep_first == first &&
                          h->prep_batch == batch && h->prep_neg_ratio == neg_ratio;
    struct Restore { crb_handle* h; ~Restore() { if (h->alt_active) crb_alt_swap(h); } } restore{h};
    if (prepared) {
        h->prep_valid = 0;
        CRB_CUDA(cudaStreamWaitEvent(s, h->ev_prep[1], 0));
        crb_alt_swap(h);   // the kernels below read the prepared copy
    } else if ((rc = crb_zero_step_counters(h, s))) return rc;
    const int32_t *du = u, *di = i, *dj = j;
    if (prepared) {
        du = h->idx[0]; di = h->idx[1]; dj = h->idx[2];
    } else {
        bool counted = false;
        if (!u) {
            if ((rc = crb_launch_sample_pairwise(h, seed, epoch, first, batch, neg_ratio, h->idx[0], h->idx[1], h->idx[2], nullptr, true, s))) return rc;
            du = h->idx[0]; di = h->idx[1]; dj = h->idx[2];
            counted = true;
        } else {
            CRB_CHECK_ARG(i && j, "null index feed");
            if (!crb_is_device_ptr(u)) { CRB_CUDA(cudaMemcpyAsync(h->idx[0], u, 4 * batch, cudaMemcpyHostToDevice, s)); du = h->idx[0]; }
            if (!crb_is_device_ptr(i)) { CRB_CUDA(cudaMemcpyAsync(h->idx[1], i, 4 * batch, cudaMemcpyHostToDevice, s)); di = h->idx[1]; }
            if (!crb_is_device_ptr(j)) { CRB_CUDA(cudaMemcpyAsync(h->idx[2], j, 4 * batch, cudaMemcpyHostToDevice, s)); dj = h->idx[2]; }
        }
        if ((rc = shard_index_work(h, du, di, dj, batch, counted, s))) return rc;
    }
    a.P = crb_to_dev(P); a.metaU = h->meta[0]; a.metaI = h->meta[1]; a.u = du; a.i = di; a.j = dj; a.sbi = h->sb[0]; a.sbj = h->sb[1];
    a.rk[0] = h->rank[0]; a.rk[1] = h->rank[1]; a.rk[2] = h->rank[2]; a.stage = h->stage;
    a.batch = batch; a.dim = P->dim; a.reg = reg; a.opt = od; a.dup_grad = h->dup_grad; a.dup_t = h->dup_t; a.block_loss = h->block_loss;
#ifdef SH_DEBUG_SWITCHES
    a.debug = getenv("CRB_SH_DEBUG") ? atoi(getenv("CRB_SH_DEBUG")) : 0;
#else
    a.debug = 0;
#endif
    if ((rc = CRB_DIM_DISPATCH(a.dim, launch_shard_t, h, a, opt_kind, s))) return rc;
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    if ((rc = crb_launch_loss_final(h, ld, s))) return rc;
    if (loss_out && !crb_is_device_ptr(loss_out)) {
        CRB_CUDA(cudaMemcpyAsync(loss_out, h->loss_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
        CRB_CUDA(cudaStreamSynchronize(s));
    }
    return CRB_OK;
}

// phase 2 (after the cross-rank barrier): sum + apply what arrived for each of this rank's rows.
extern "C" int crb_shard_apply_inbox(crb_handle* h, const crb_shard* shard, const crb_opt* opt, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && shard, "null argument");
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    ShardDev sd;
    if ((rc = shard_to_dev(shard, &sd, (uint32_t)od.step))) return rc;
    const crb_table* Q = &shard->q[shard->rank];
    if ((rc = crb_table_check(Q, opt_kind, "Q shard"))) return rc;
    CRB_CUDA(cudaSetDevice(h->device));
    InboxApplyArgs a;
    a.Q = sd.q[shard->rank]; a.rows = Q->rows; a.rows_cap = shard->rows_cap; a.n_ranks = shard->n_ranks; a.dim = Q->dim;
    a.stamp = (uint32_t)od.step; a.grad = shard->inbox_grad[shard->rank]; a.stamps = shard->inbox_stamp[shard->rank]; a.opt = od;
    return CRB_DIM_DISPATCH(a.dim, launch_inbox_t, h, a, opt_kind, s);
}

extern "C" int crb_shard_barrier(crb_handle* h, const crb_shard* shard, uint32_t ticket, int32_t timeout_ms, void* stream) {
    CRB_CHECK_ARG(h && shard && ticket >= 1, "null argument / ticket");
    CRB_CHECK_ARG(shard->n_ranks >= 1 && shard->n_ranks <= CRB_MAX_RANKS, "shard descriptor");
    BarrierArgs a;
    a.n_ranks = shard->n_ranks; a.rank = shard->rank; a.ticket = ticket;
    a.timeout_ns = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 10000) * 1000000ULL;
    for (int r = 0; r < CRB_MAX_RANKS; ++r) a.flags[r] = r < shard->n_ranks ? shard->flags[r] : nullptr;
    for (int r = 0; r < shard->n_ranks; ++r) CRB_CHECK_ARG(a.flags[r], "shard descriptor: null flag pointer");
    CRB_CUDA(cudaSetDevice(h->device));
    shard_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

extern "C" int crb_sampler_errors(crb_handle* h, uint32_t* n_rows, void* stream) {
    CRB_CHECK_ARG(h && n_rows, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CUDA(cudaSetDevice(h->device));
    if (h->aux_stream) CRB_CUDA(cudaStreamSynchronize(h->aux_stream));
    CRB_CUDA(cudaStreamSynchronize(s));
    unsigned int err[2] = {0, 0};
    const bool alt_active = h->alt_active;
    crb_step_ctr* c0 = alt_active ? h->alt.ctr : h->ctr;
    crb_step_ctr* c1 = alt_active ? h->ctr : h->alt.ctr;
    CRB_CUDA(cudaMemcpy(&err[0], &c0->sampler_err, sizeof(unsigned int), cudaMemcpyDeviceToHost));
    if (c1) CRB_CUDA(cudaMemcpy(&err[1], &c1->sampler_err, sizeof(unsigned int), cudaMemcpyDeviceToHost));
    if (err[0]) CRB_CUDA(cudaMemset(&c0->sampler_err, 0, sizeof(unsigned int)));
    if (err[1]) CRB_CUDA(cudaMemset(&c1->sampler_err, 0, sizeof(unsigned int)));
    *n_rows = err[0] + err[1];
    return CRB_OK;
}

extern "C" int crb_shard_check(crb_handle* h, const crb_shard* shard, void* stream) {
    CRB_CHECK_ARG(h && shard, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t serr = 0;
    int rc = crb_sampler_errors(h, &serr, stream);
    if (rc) return rc;
    unsigned int v = 0;
    uint32_t* word = shard->flags[shard->rank] + CRB_SHARD_ERR;
    CRB_CUDA(cudaMemcpyAsync(&v, word, sizeof(v), cudaMemcpyDeviceToHost, s));
    CRB_CUDA(cudaStreamSynchronize(s));
    if (v) {
        CRB_CUDA(cudaMemsetAsync(word, 0, sizeof(v), s));
        crb_set_error("multi-GPU step: a cross-rank barrier timed out (a peer rank failed or fell behind)");
        return CRB_ERR_STATE;
    }
    if (serr) {
        crb_set_error("sampler: %u rows found no admissible negative", serr);
        return CRB_ERR_SAMPLER;
    }
    return CRB_OK;
}

// ------------------------------------------------------------------------------------------------ pointwise step (MF / GMF)
// sess.run([train, loss], {u_idx, i_idx, y}) on this rank's slice of the union batch (GMF.py:37-49): logit = sum_k p_uk q_ik [h_k];
// loss = get_loss(loss_func, y, logit) + reg * (l2(p_u) + l2(q_i)).  The item row comes from the staging buffer (repeated in this
// rank's batch) or straight from its owner; its gradient goes to a local slot or straight into the owner's inbox -- as in
// shard_step_kernel.  The next iteration's rows are requested before the current one is computed.
template <int LANES, int VPL, int OPT, bool GMF>
__global__ void __launch_bounds__(256) shard_pw_step_kernel(ShardPwArgs a) {
    constexpr int GPW = 32 / LANES;
    __shared__ float4 s_h[GMF ? 256 * VPL : 1];
    const int lane = threadIdx.x & 31, gl = lane % LANES, sub = lane / LANES;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int G = a.sh.n_ranks;
    float4 hreg[VPL], gh[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c = (gl + LANES * v) * 4;
        hreg[v] = (GMF && c < a.dim) ? ld4(a.hvec + c) : make_float4(1.f, 1.f, 1.f, 1.f);
        gh[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    double loss_acc = 0.0;
    const int64_t stride = n_warps * GPW;
    auto clampt = [&](int64_t b) { const int64_t t_ = b + sub; return t_ < a.batch ? t_ : a.batch - 1; };
    auto load_item = [&](RowRegs<LANES, VPL>& r, int32_t item, int32_t sb) {
        const float* src = sb >= 0 ? a.stage + (int64_t)sb * a.dim : a.sh.q[item % G].w + (int64_t)(item / G) * a.dim;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (gl + LANES * v) * 4;
            r.w[v] = c < a.dim ? ld4(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    int32_t nu, ni, nsb;
    RowRegs<LANES, VPL> nru, nri;
    { const int64_t t0 = clampt(warp * GPW); nu = a.u[t0]; ni = a.i[t0]; nsb = a.sbi[t0]; }
    row_load_w<LANES, VPL>(nru, a.P, nu, a.dim, gl);
    load_item(nri, ni, nsb);
    for (int64_t base = warp * GPW; base < a.batch; base += stride) {
        const int64_t t = base + sub;
        const bool active = t < a.batch;
        const int64_t tt = active ? t : a.batch - 1;
        const int32_t u = nu, i = ni, sbi = nsb;
        RowRegs<LANES, VPL> ru = nru, ri = nri;
        { const int64_t t1 = clampt(base + stride); nu = a.u[t1]; ni = a.i[t1]; nsb = a.sbi[t1]; }
        row_load_w<LANES, VPL>(nru, a.P, nu, a.dim, gl);
        load_item(nri, ni, nsb);
        const float y = a.y[tt];
        const unsigned long long mu = a.metaU[u];
        ru.last = OptTraits<OPT>::replay ? a.P.last[u] : 0;
        const bool su = (uint32_t)mu == 1u || replay_pending<OPT>(ru.last, a.opt);
        if (OptTraits<OPT>::has_s1 && su) row_load_state<LANES, VPL, OPT>(ru, a.P, u, a.dim, gl);
        if (replay_pending<OPT>(ru.last, a.opt)) row_replay<LANES, VPL, OPT>(ru, a.opt, a.opt.step);
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
            if (sbi < 0) {
                shard_send<LANES, VPL>(a.sh.send, i, gi, a.dim, gl);
                if (gl == 0) a.metaI[i] = 0ULL;
            } else {
                const uint32_t slot = (uint32_t)sbi + a.rk[1][t];
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    const int c = (gl + LANES * v) * 4;
                    if (c < a.dim) st4(a.dup_grad + (int64_t)slot * a.dim + c, gi[v]);
                }
                if (gl == 0) a.dup_t[slot] = ((uint32_t)t << 2) | 1u;
            }
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

__global__ void __launch_bounds__(256) shard_resolve1_kernel(const unsigned long long* __restrict__ metaI, const int32_t* __restrict__ i, int64_t batch,
                                                             int32_t* __restrict__ sbi) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < batch; t += stride) {
        const unsigned long long mi = metaI[i[t]];
        sbi[t] = (uint32_t)mi > 1u ? (int32_t)(mi >> 32) : -1;
    }
}

// this rank's total gradient of the replicated dense variable (block partials summed in block order) -> slot `rank` of EVERY rank's
// dense inbox
struct DenseSendArgs {
    const float* parts;
    int n_parts, n, n_ranks, rank;
    float* inbox[CRB_MAX_RANKS];
};
__global__ void __launch_bounds__(256) shard_dense_send_kernel(DenseSendArgs a) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < a.n; k += gridDim.x * blockDim.x) {
        float g = 0.f;
        for (int p = 0; p < a.n_parts; ++p) g += a.parts[(int64_t)p * a.n + k];
        for (int r = 0; r < a.n_ranks; ++r) a.inbox[r][a.rank * CRB_SHARD_DENSE + k] = g;
    }
}

template <int LANES, int VPL>
static int launch_shard_pw_t(crb_handle* h, const ShardPwArgs& a, int opt_kind, bool gmf, cudaStream_t s) {
    const int gpb = 256 / LANES;
    int rc;
    if ((rc = crb_prof_begin(h, s, 1))) return rc;
    item_fetch_kernel<LANES, VPL><<<h->sm_count * 8, 256, 0, s>>>(a.sh, h->dup_rows, h->ctr, h->stage, a.dim);
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    if ((rc = crb_prof_end(h, s, 1))) return rc;
    if ((rc = crb_prof_begin(h, s))) return rc;
#define CRB_SHPW_CASE(O)                                                                                  \
    case O: {                                                                                             \
        if (gmf) {                                                                                        \
            const int grid = one_wave(h, shard_pw_step_kernel<LANES, VPL, O, true>, a.batch, gpb);        \
            h->step_grid = grid;                                                                          \
            shard_pw_step_kernel<LANES, VPL, O, true><<<grid, 256, 0, s>>>(a);                            \
        } else {                                                                                          \
            const int grid = one_wave(h, shard_pw_step_kernel<LANES, VPL, O, false>, a.batch, gpb);       \
            h->step_grid = grid;                                                                          \
            shard_pw_step_kernel<LANES, VPL, O, false><<<grid, 256, 0, s>>>(a);                           \
        }                                                                                                 \
        break;                                                                                            \
    }
    switch (opt_kind) { CRB_SHPW_CASE(OPT_SGD) CRB_SHPW_CASE(OPT_ADAGRAD) CRB_SHPW_CASE(OPT_ADAM_LAZY) CRB_SHPW_CASE(OPT_ADAM_TF1) }
#undef CRB_SHPW_CASE
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    if ((rc = crb_prof_end(h, s))) return rc;
    DupArgs d;
    d.tab[0] = a.P; d.tab[1] = a.P;
    d.meta[0] = a.metaU; d.meta[1] = a.metaI;
    d.dim = a.dim; d.opt = a.opt;
    d.dup_rows = h->dup_rows; d.work = h->work; d.multi = h->multi;
    d.dup_grad = h->dup_grad; d.dup_t = h->dup_t; d.partial = h->partial; d.ctr = h->ctr;
    d.send = a.sh.send;
    const int grid = h->sm_count * 4;
    if ((rc = crb_prof_begin(h, s, 2))) return rc;
#define CRB_SHDUP_CASE(O)                                                        \
    case O:                                                                      \
        dup_reduce_kernel<LANES, VPL, O, true><<<grid, 256, 0, s>>>(d);          \
        dup_final_kernel<LANES, VPL, O, true><<<h->sm_count, 256, 0, s>>>(d);    \
        break;
    switch (opt_kind) { CRB_SHDUP_CASE(OPT_SGD) CRB_SHDUP_CASE(OPT_ADAGRAD) CRB_SHDUP_CASE(OPT_ADAM_LAZY) CRB_SHDUP_CASE(OPT_ADAM_TF1) }
#undef CRB_SHDUP_CASE
    h->launches += 2;
    CRB_CUDA(cudaGetLastError());
    return crb_prof_end(h, s, 2);
}

extern "C" int crb_shard_step_compute_pointwise(crb_handle* h, int32_t kind, const crb_table* P, const crb_shard* shard, const float* hvec,
                                                const crb_opt* opt, int32_t loss_kind, const int32_t* u, const int32_t* i, const float* y,
                                                uint64_t seed, uint32_t epoch, int64_t first, int32_t neg_ratio, int64_t batch, float reg,
                                                double* loss_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && P && shard, "null argument");
    CRB_CHECK_ARG(batch > 0 && batch < 0x20000000LL, "batch");
    CRB_CHECK_ARG(kind == CRB_SCORE_DOT || kind == CRB_SCORE_GMF, "kind must be CRB_SCORE_DOT (MF) or CRB_SCORE_GMF");
    CRB_CHECK_ARG(loss_kind == CRB_LOSS_CROSS_ENTROPY || loss_kind == CRB_LOSS_SQUARE, "pointwise loss must be cross_entropy or square");
    const bool gmf = kind == CRB_SCORE_GMF;
    CRB_CHECK_ARG(!gmf || (hvec && crb_is_device_ptr(hvec)), "GMF needs the device vector h");
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    if ((rc = crb_table_check(P, opt_kind, "P"))) return rc;
    ShardPwArgs a;
    if ((rc = shard_to_dev(shard, &a.sh, (uint32_t)od.step))) return rc;
    CRB_CHECK_ARG(shard->q[shard->rank].dim == P->dim, "P.dim != Q.dim");
    CRB_CHECK_ARG(!gmf || P->dim <= CRB_SHARD_DENSE, "GMF: dim exceeds the dense inbox");
    for (int r = 0; gmf && r < shard->n_ranks; ++r) CRB_CHECK_ARG(shard->dense_inbox[r], "shard descriptor: null dense inbox");
    const int64_t n_items = shard_items(shard);
    CRB_CUDA(cudaSetDevice(h->device));
    if (h->alt_active) crb_alt_swap(h);
    if ((rc = crb_ws_reserve(h, batch, P->dim, 4, s))) return rc;
    if ((rc = crb_meta_reserve(h, 0, P->rows, s))) return rc;
    if ((rc = crb_meta_reserve(h, 1, n_items, s))) return rc;
    if ((rc = stage_reserve(h, P->dim, s))) return rc;
    if (gmf) {
        const int64_t need = (int64_t)h->loss_blocks * P->dim;
        if (need > h->cap_dense) {
            CRB_CUDA(cudaStreamSynchronize(s));
            cudaFree(h->dense_grad);
            h->dense_grad = nullptr;
            CRB_CUDA(cudaMalloc(&h->dense_grad, sizeof(float) * need));
            h->cap_dense = need;
        }
    }
    if ((rc = crb_zero_step_counters(h, s))) return rc;
    const int32_t *du = u, *di = i;
    const float* dy = y;
    bool counted = false;
    if (!u) {
        if ((rc = crb_launch_sample_pointwise(h, seed, epoch, first, batch, neg_ratio, h->idx[0], h->idx[1], h->yv, true, s))) return rc;
        du = h->idx[0]; di = h->idx[1]; dy = h->yv;
        counted = true;
    } else {
        CRB_CHECK_ARG(i && y, "null feed");
        if (!crb_is_device_ptr(u)) { CRB_CUDA(cudaMemcpyAsync(h->idx[0], u, 4 * batch, cudaMemcpyHostToDevice, s)); du = h->idx[0]; }
        if (!crb_is_device_ptr(i)) { CRB_CUDA(cudaMemcpyAsync(h->idx[1], i, 4 * batch, cudaMemcpyHostToDevice, s)); di = h->idx[1]; }
        if (!crb_is_device_ptr(y)) { CRB_CUDA(cudaMemcpyAsync(h->yv, y, 4 * batch, cudaMemcpyHostToDevice, s)); dy = h->yv; }
    }
    const int32_t* idx[3] = {du, di, nullptr};
    const int role_table[3] = {0, 1, 0};
    if (!counted && (rc = crb_count_rows(h, batch, 2, idx, role_table, s))) return rc;
    if ((rc = crb_launch_assign(h, batch, 2, idx, role_table, s))) return rc;
    {
        int64_t blocks = (batch + 255) / 256;
        const int64_t capb = (int64_t)h->sm_count * 16;
        shard_resolve1_kernel<<<(int)(blocks < capb ? blocks : capb), 256, 0, s>>>(h->meta[1], di, batch, h->sb[0]);
        h->launches++;
        CRB_CUDA(cudaGetLastError());
    }
    a.P = crb_to_dev(P); a.metaU = h->meta[0]; a.metaI = h->meta[1]; a.u = du; a.i = di; a.y = dy; a.sbi = h->sb[0];
    a.rk[0] = h->rank[0]; a.rk[1] = h->rank[1]; a.stage = h->stage; a.hvec = gmf ? hvec : nullptr; a.hpart = h->dense_grad;
    a.batch = batch; a.dim = P->dim; a.loss_kind = loss_kind; a.reg = reg; a.opt = od;
    a.dup_grad = h->dup_grad; a.dup_t = h->dup_t; a.block_loss = h->block_loss;
    if ((rc = CRB_DIM_DISPATCH(a.dim, launch_shard_pw_t, h, a, opt_kind, gmf, s))) return rc;
    if (gmf) {
        DenseSendArgs ds;
        ds.parts = h->dense_grad; ds.n_parts = h->step_grid; ds.n = P->dim; ds.n_ranks = shard->n_ranks; ds.rank = shard->rank;
        for (int r = 0; r < CRB_MAX_RANKS; ++r) ds.inbox[r] = r < shard->n_ranks ? shard->dense_inbox[r] : nullptr;
        shard_dense_send_kernel<<<1, 256, 0, s>>>(ds);
        h->launches++;
        CRB_CUDA(cudaGetLastError());
    }
    double* ld = (loss_out && crb_is_device_ptr(loss_out)) ? loss_out : h->loss_dev;
    if ((rc = crb_launch_loss_final(h, ld, s))) return rc;
    if (loss_out && !crb_is_device_ptr(loss_out)) {
        CRB_CUDA(cudaMemcpyAsync(loss_out, h->loss_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
        CRB_CUDA(cudaStreamSynchronize(s));
    }
    return CRB_OK;
}

// after the barrier: h -= dense optimizer step on the gradient summed over the ranks in rank order (identical on every rank)
extern "C" int crb_shard_apply_dense(crb_handle* h, const crb_shard* shard, const crb_opt* opt, float* hvec, float* h_s1, float* h_s2, int32_t dim,
                                     void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CHECK_ARG(h && shard && hvec && dim > 0 && dim <= CRB_SHARD_DENSE, "bad argument");
    CRB_CHECK_ARG(shard->dense_inbox[shard->rank], "shard descriptor: null dense inbox");
    OptDev od;
    int opt_kind = 0;
    int rc = crb_opt_to_dev(h, opt, &od, &opt_kind, s);
    if (rc) return rc;
    CRB_CHECK_ARG(opt_kind == OPT_SGD || h_s1, "h optimizer slot s1 is NULL");
    CRB_CHECK_ARG((opt_kind != OPT_ADAM_LAZY && opt_kind != OPT_ADAM_TF1) || h_s2, "h optimizer slot s2 is NULL");
    CRB_CUDA(cudaSetDevice(h->device));
    // the n_ranks partial vectors sit CRB_SHARD_DENSE floats apart in this rank's dense inbox
    return crb_launch_dense_apply_strided(h, hvec, h_s1, h_s2, shard->dense_inbox[shard->rank], shard->n_ranks, dim, CRB_SHARD_DENSE, opt_kind, od, s);
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
