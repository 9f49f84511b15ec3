// K3 of the BPR step (model/ranking/BPR.py:31-44) with the row gathers moved off the registers: a warp-specialised persistent
// kernel, one CTA per SM.
//
//   producer warps  a STAGE is the GPW triplets one consumer warp-iteration works on.  A producer reads the stage's ids, multiplicity
//                   words, `last` steps and occurrence ranks (lane per stage, 32 stages at a time), decides which rows the step
//                   needs (the weights always; the optimizer slots only for a row this triplet will update in place or must
//                   replay), writes a 64-byte header and copies every row into the stage's slot of a shared-memory ring with
//                   16-byte cp.async (LDGSTS: global -> shared, no registers, 512 coalesced bytes per warp instruction); the
//                   slot's `full` mbarrier completes when every lane's copies have landed.
//   consumer warps  wait on `full`, read header + rows with LDS.128, release the slot (a per-slot lap counter) and run exactly the
//                   arithmetic of bpr_step_kernel (same device functions: forward, backward, in-place apply or gradient slot), so the
//                   tables are bit-identical to the register-staged kernel.
//
// What it buys: the bytes in flight per SM are bounded by shared memory (~200 KB of ring) instead of by the registers of the
// resident warps (24 warps x ~3.5 KB), and nothing on the index side (ids -> multiplicities -> rows) is on a consumer's critical path.
#include <stdlib.h>

#include "bpr_args.cuh"

#define RG_WARPS 16           // producer + consumer warps of the CTA (the split is a launch parameter)
#define RG_THREADS (RG_WARPS * 32)
#define RG_SMEM_BUDGET (200 * 1024)

struct RingHdr {              // 64 bytes per triplet
    int32_t u, i, j;
    uint32_t flags;           // bit 0: active; bits 1..3: optimizer slots of the u / i / j row were fetched
    unsigned long long mu, mi, mj;
    int32_t lu, li, lj;
    uint32_t rku, rki, rkj;
};
static_assert(sizeof(RingHdr) == 64, "header size");

__device__ __forceinline__ uint32_t rg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rg_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rg_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void rg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rg_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rg_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(rg_smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
    } while (!done);
}
// 16 bytes global -> shared without passing through registers (LDGSTS, L2 only)
__device__ __forceinline__ void rg_copy16(void* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rg_smem_u32(dst)), "l"(src) : "memory");
}

template <int OPT> struct RingRows { static constexpr int per_triplet = 3 * (1 + (OptTraits<OPT>::has_s1 ? 1 : 0) + (OptTraits<OPT>::has_s2 ? 1 : 0)); };

template <int LANES, int VPL>
__device__ __forceinline__ void ring_read_row(float4* dst, const float* src, int dim, int gl, float4 pad) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c = (gl + LANES * v) * 4;
        dst[v] = c < dim ? *reinterpret_cast<const float4*>(src + c) : pad;
    }
}

template <int LANES, int VPL, int OPT>
__global__ void __launch_bounds__(RG_THREADS, 1) bpr_ring_kernel(BprArgs a, int n_stages_ring, int n_producers) {
    const int RG_PRODUCERS = n_producers, RG_CONSUMERS = RG_WARPS - n_producers;
    constexpr int GPW = 32 / LANES;
    constexpr int ROWS = RingRows<OPT>::per_triplet;
    extern __shared__ __align__(128) unsigned char rg_sm[];
    const int NS = n_stages_ring;
    const uint32_t rb = (uint32_t)a.dim * 4u;                          // bytes per row
    const uint32_t stage_bytes = GPW * (64u + ROWS * rb);
    // Hand-off per slot.  `full[s]` is an mbarrier (it has to count the bulk copies' bytes); its waiters test a phase PARITY, which
    // is only meaningful for a waiter at most one phase away.  Successive occupants of a slot are produced by different lanes and
    // consumed by different warps, so the order between laps is carried by a plain counter instead: done[s] = number of occupants
    // of slot s that have been consumed.  The producer of occupant k waits for done[s] == k before it touches the slot; the consumer
    // of occupant k waits for done[s] == k (occupant k-1 consumed, hence full[s] is in phase k) before it tests full[s]'s parity.
    uint64_t* full = reinterpret_cast<uint64_t*>(rg_sm);
    volatile uint32_t* done = reinterpret_cast<volatile uint32_t*>(full + NS);
    unsigned char* ring = rg_sm + ((2 * NS * 8 + 127) & ~127);
    __shared__ double s_loss[RG_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { rg_mbar_init(full + s, 33); done[s] = 0u; }   // 32 asynchronous copy arrivals + the header's
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // stage q of this CTA = warp-iteration group (blockIdx.x + q * gridDim.x): triplets [g * GPW, g * GPW + GPW)
    const int64_t n_groups = (a.batch + GPW - 1) / GPW;
    const int64_t n_q = n_groups > blockIdx.x ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (warp < RG_PRODUCERS) {
        // ------------------------------------------------------------------------------------------------ producers
        // A producer warp takes 32 consecutive stages at a time.  Index side, lane per stage: ids and occurrence ranks, then the
        // multiplicity words and `last` steps (two dependent round trips for 32 stages at once) -> which rows the stage needs.
        // Copy side, the whole warp per stage: a row is 16-byte cp.async (LDGSTS) chunks, one per lane -- 512 coalesced bytes per
        // warp instruction, no registers; the lanes' asynchronous arrivals on the slot's mbarrier fire when their copies have landed.
        // (1-D cp.async.bulk copies of one 512-byte row each were measured first: ~19 copies per microsecond per SM whatever the
        // optimizer, i.e. 1.4 TB/s -- the per-copy cost of the TMA path is too high for rows this small.)
        const int chunks = (int)(rb / 16u);
        for (int64_t q0 = (int64_t)warp * 32; q0 < n_q; q0 += 32 * RG_PRODUCERS) {
            const int64_t q = q0 + lane;
            const int64_t g = blockIdx.x + q * (int64_t)gridDim.x;
            RingHdr hd[GPW];
#pragma unroll
            for (int k = 0; k < GPW; ++k) {
                const int64_t t = g * GPW + k;
                RingHdr& H = hd[k];
                H.flags = 0u; H.u = 0; H.i = 0; H.j = 0;
                if (q >= n_q || t >= a.batch) continue;
                H.u = a.u[t]; H.i = a.i[t]; H.j = a.j[t];
                H.rku = a.rk[0][t]; H.rki = a.rk[1][t]; H.rkj = a.rk[2][t];
            }
#pragma unroll
            for (int k = 0; k < GPW; ++k) {
                const int64_t t = g * GPW + k;
                RingHdr& H = hd[k];
                if (q >= n_q || t >= a.batch) continue;
                H.mu = a.metaU[H.u]; H.mi = a.metaI[H.i]; H.mj = a.metaI[H.j];
                H.lu = OptTraits<OPT>::replay ? a.P.last[H.u] : 0;
                H.li = OptTraits<OPT>::replay ? a.Q.last[H.i] : 0;
                H.lj = OptTraits<OPT>::replay ? a.Q.last[H.j] : 0;
            }
#pragma unroll
            for (int k = 0; k < GPW; ++k) {
                const int64_t t = g * GPW + k;
                RingHdr& H = hd[k];
                if (q >= n_q || t >= a.batch) continue;
                const bool su = OptTraits<OPT>::has_s1 && ((uint32_t)H.mu == 1u || replay_pending<OPT>(H.lu, a.opt));
                const bool si = OptTraits<OPT>::has_s1 && ((uint32_t)H.mi == 1u || replay_pending<OPT>(H.li, a.opt));
                const bool sj = OptTraits<OPT>::has_s1 && ((uint32_t)H.mj == 1u || replay_pending<OPT>(H.lj, a.opt));
                H.flags = 1u | (su ? 2u : 0u) | (si ? 4u : 0u) | (sj ? 8u : 0u);
            }
            const int n_here = (int)(n_q - q0 < 32 ? n_q - q0 : 32);
            for (int s = 0; s < n_here; ++s) {
                const int64_t qs = q0 + s;
                const int slot = (int)(qs % NS);
                const uint32_t occ = (uint32_t)(qs / NS);
                unsigned char* st = ring + (size_t)slot * stage_bytes;
                if (lane == s) {
                    if (occ) while (done[slot] < occ) __nanosleep(32);   // every earlier occupant of the slot has been consumed
                    __threadfence_block();
#pragma unroll
                    for (int k = 0; k < GPW; ++k) *reinterpret_cast<RingHdr*>(st + k * 64) = hd[k];
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < GPW; ++k) {
                    const uint32_t fl = __shfl_sync(0xffffffffu, hd[k].flags, s);
                    const int32_t ru_ = __shfl_sync(0xffffffffu, hd[k].u, s), ri_ = __shfl_sync(0xffffffffu, hd[k].i, s),
                                  rj_ = __shfl_sync(0xffffffffu, hd[k].j, s);
                    if (!(fl & 1u)) continue;   // warp-uniform
                    unsigned char* rows = st + GPW * 64 + (size_t)k * ROWS * rb;
                    const int64_t ou = (int64_t)ru_ * a.dim, oi = (int64_t)ri_ * a.dim, oj = (int64_t)rj_ * a.dim;
                    for (int c = lane; c < chunks; c += 32) {
                        rg_copy16(rows + 0 * rb + 16 * c, a.P.w + ou + 4 * c);
                        rg_copy16(rows + 1 * rb + 16 * c, a.Q.w + oi + 4 * c);
                        rg_copy16(rows + 2 * rb + 16 * c, a.Q.w + oj + 4 * c);
                        if (OptTraits<OPT>::has_s1) {
                            if (fl & 2u) rg_copy16(rows + 3 * rb + 16 * c, a.P.s1 + ou + 4 * c);
                            if (fl & 4u) rg_copy16(rows + 4 * rb + 16 * c, a.Q.s1 + oi + 4 * c);
                            if (fl & 8u) rg_copy16(rows + 5 * rb + 16 * c, a.Q.s1 + oj + 4 * c);
                        }
                        if (OptTraits<OPT>::has_s2) {
                            if (fl & 2u) rg_copy16(rows + 6 * rb + 16 * c, a.P.s2 + ou + 4 * c);
                            if (fl & 4u) rg_copy16(rows + 7 * rb + 16 * c, a.Q.s2 + oi + 4 * c);
                            if (fl & 8u) rg_copy16(rows + 8 * rb + 16 * c, a.Q.s2 + oj + 4 * c);
                        }
                    }
                }
                // every lane: "arrive on full[slot] when all my copies so far have landed"; lane 0 adds the arrival that publishes
                // the header (written before the __syncwarp above)
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(rg_smem_u32(full + slot)) : "memory");
                if (lane == 0) rg_mbar_arrive(full + slot);
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------------ consumers
        const int cw = warp - RG_PRODUCERS;
        const int gl = lane % LANES, sub = lane / LANES;
        double loss_acc = 0.0;
        const float pad1 = OptTraits<OPT>::has_s2 ? 0.f : 1.f;   // padding lanes: see row_load_state
        for (int64_t q = cw; q < n_q; q += RG_CONSUMERS) {
            const int slot = (int)(q % NS);
            const uint32_t occ = (uint32_t)(q / NS);
            if (occ) while (done[slot] < occ) __nanosleep(32);
            rg_mbar_wait(full + slot, occ & 1u);
            const unsigned char* st = ring + (size_t)slot * stage_bytes;
            const RingHdr H = *reinterpret_cast<const RingHdr*>(st + sub * 64);
            const bool active = (H.flags & 1u) != 0u;
            const float* rows = reinterpret_cast<const float*>(st + GPW * 64 + (size_t)sub * ROWS * rb);
            const int rf = a.dim;   // floats per row
            RowRegs<LANES, VPL> ru, ri, rj;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) {
                ring_read_row<LANES, VPL>(ru.w, rows + 0 * rf, a.dim, gl, z4);
                ring_read_row<LANES, VPL>(ri.w, rows + 1 * rf, a.dim, gl, z4);
                ring_read_row<LANES, VPL>(rj.w, rows + 2 * rf, a.dim, gl, z4);
                if (OptTraits<OPT>::has_s1) {
                    const float4 p1 = make_float4(pad1, pad1, pad1, pad1);
                    if (H.flags & 2u) ring_read_row<LANES, VPL>(ru.s1, rows + 3 * rf, a.dim, gl, p1);
                    if (H.flags & 4u) ring_read_row<LANES, VPL>(ri.s1, rows + 4 * rf, a.dim, gl, p1);
                    if (H.flags & 8u) ring_read_row<LANES, VPL>(rj.s1, rows + 5 * rf, a.dim, gl, p1);
                }
                if (OptTraits<OPT>::has_s2) {
                    const float4 p2 = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (H.flags & 2u) ring_read_row<LANES, VPL>(ru.s2, rows + 6 * rf, a.dim, gl, p2);
                    if (H.flags & 4u) ring_read_row<LANES, VPL>(ri.s2, rows + 7 * rf, a.dim, gl, p2);
                    if (H.flags & 8u) ring_read_row<LANES, VPL>(rj.s2, rows + 8 * rf, a.dim, gl, p2);
                }
            } else {
#pragma unroll
                for (int v = 0; v < VPL; ++v) { ru.w[v] = z4; ri.w[v] = z4; rj.w[v] = z4; }
            }
            __syncwarp();
            if (lane == 0) { __threadfence_block(); done[slot] = occ + 1u; }   // everything of the stage is in registers: the slot may be refilled
            ru.last = H.lu; ri.last = H.li; rj.last = H.lj;
            if (active) {
                if (replay_pending<OPT>(ru.last, a.opt)) row_replay<LANES, VPL, OPT>(ru, a.opt, a.opt.step);
                if (replay_pending<OPT>(ri.last, a.opt)) row_replay<LANES, VPL, OPT>(ri, a.opt, a.opt.step);
                if (replay_pending<OPT>(rj.last, a.opt)) row_replay<LANES, VPL, OPT>(rj, a.opt, a.opt.step);
            }
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
                const uint32_t t = (uint32_t)((blockIdx.x + q * (int64_t)gridDim.x) * GPW + sub);
                emit_row<LANES, VPL, OPT>(ru, gu, a.P, a.metaU, H.u, H.mu, H.rku, t, 0u, a.dim, gl, a.opt, a.dup_grad, a.dup_t);
                emit_row<LANES, VPL, OPT>(ri, gi, a.Q, a.metaI, H.i, H.mi, H.rki, t, 1u, a.dim, gl, a.opt, a.dup_grad, a.dup_t);
                emit_row<LANES, VPL, OPT>(rj, gj, a.Q, a.metaI, H.j, H.mj, H.rkj, t, 2u, a.dim, gl, a.opt, a.dup_grad, a.dup_t);
            }
        }
        for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
        if (lane == 0) s_loss[cw] = loss_acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < RG_CONSUMERS; ++k) t += s_loss[k];
        a.block_loss[blockIdx.x] = t;
    }
}

bool crb_bpr_ring_enabled(int dim) {
    const char* e = getenv("CRB_BPR_RING");   // 0 = register-staged kernel, 1 = ring kernel (read per call: A/B inside one process)
    const int mode = e ? atoi(e) : 0;
    return mode == 1 && dim >= 32;
}

template <int LANES, int VPL>
static int launch_ring_t(crb_handle* h, const BprArgs& a, int opt_kind, cudaStream_t s) {
    constexpr int GPW = 32 / LANES;
    const int64_t n_groups = (a.batch + GPW - 1) / GPW;
    int grid = (int)(n_groups < h->sm_count ? n_groups : h->sm_count);
    if (grid > h->loss_blocks) grid = h->loss_blocks;
    h->step_grid = grid;
    int n_prod = getenv("CRB_RING_PRODUCERS") ? atoi(getenv("CRB_RING_PRODUCERS")) : 2;
    if (n_prod < 1) n_prod = 1;
    if (n_prod > RG_WARPS - 1) n_prod = RG_WARPS - 1;
#define CRB_RING_CASE(O)                                                                                                     \
    case O: {                                                                                                                \
        const size_t stage_bytes = (size_t)GPW * (64 + RingRows<O>::per_triplet * (size_t)a.dim * 4);                        \
        int ns = (int)((RG_SMEM_BUDGET - 1024) / (stage_bytes + 16));                                                        \
        if (ns > 256) ns = 256;                                                                                              \
        if (ns < 2) { crb_set_error("ring kernel: row too large for the shared-memory ring"); return CRB_ERR_UNSUPPORTED; }  \
        const size_t smem = ((2 * (size_t)ns * 8 + 127) & ~(size_t)127) + ns * stage_bytes;                                  \
        CRB_CUDA(cudaFuncSetAttribute(bpr_ring_kernel<LANES, VPL, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        bpr_ring_kernel<LANES, VPL, O><<<grid, RG_THREADS, smem, s>>>(a, ns, n_prod);                                                \
        break;                                                                                                               \
    }
    switch (opt_kind) {
        CRB_RING_CASE(OPT_SGD)
        CRB_RING_CASE(OPT_ADAGRAD)
        CRB_RING_CASE(OPT_ADAM_LAZY)
        CRB_RING_CASE(OPT_ADAM_TF1)
    }
#undef CRB_RING_CASE
    h->launches++;
    CRB_CUDA(cudaGetLastError());
    return CRB_OK;
}

int crb_launch_bpr_ring(crb_handle* h, const BprArgs& a, int opt_kind, cudaStream_t s) {
    return a.dim <= 32    ? launch_ring_t<8, 1>(h, a, opt_kind, s)
           : a.dim <= 64  ? launch_ring_t<16, 1>(h, a, opt_kind, s)
           : a.dim <= 128 ? launch_ring_t<32, 1>(h, a, opt_kind, s)
           : a.dim <= 256 ? launch_ring_t<32, 2>(h, a, opt_kind, s)
                          : launch_ring_t<32, 4>(h, a, opt_kind, s);
}
