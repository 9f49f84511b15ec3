// K4: the duplicate-row reduction shared by every row-sparse training step (see rowopt.cuh for the multiplicity split).
// Templates only: train.cu instantiates the single-GPU kernels (SHARD = false), train_sharded.cu the multi-GPU ones (SHARD = true:
// the summed gradient of an ITEM row is not applied here but written into its owner's inbox over NVLink peer memory).
#pragma once
#include "rowopt.cuh"
// ------------------------------------------------------------------------------------------------ K4
// Sum of gradient slots [lo, hi) of one duplicate row into acc (lane-group view).  Rows with <= 32 occurrences
// are summed in ascending triplet order (deterministic); longer ones in slot order.
template <int LANES, int VPL>
__device__ __forceinline__ void sum_slots(float4* acc, const float* __restrict__ dup_grad, const uint32_t* __restrict__ dup_t,
                                          const uint32_t* __restrict__ dup_src, uint32_t lo, uint32_t hi, bool ordered, int dim, int gl) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t n = hi - lo;
    // Two to four occurrences (all but ~1e-3 of the duplicated rows of a uniform batch): every slot is requested at once, so the
    // row costs one memory round trip instead of one per occurrence; the adds then run in ascending key order from registers
    // (for two occurrences the order is immaterial: (0 + a) + b == (0 + b) + a bit for bit).
    if (ordered && n >= 2u && n <= 4u && VPL <= 2) {
        float4 G[4][VPL];
        uint32_t key[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool on = (uint32_t)q < n;
            const uint32_t sq = lo + (on ? q : 0);
            key[q] = on ? (n > 2u ? dup_t[sq] : (uint32_t)q) : 0xFFFFFFFFu;
            const int64_t src = dup_src ? dup_src[sq] : sq;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c = (gl + LANES * v) * 4;
                G[q][v] = (on && c < dim) ? ld4(dup_grad + src * dim + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        // rank of each slot among the keys (keys are unique; absent slots carry the maximum and sort last)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if ((uint32_t)k >= n) break;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int rk = 0;
#pragma unroll
                for (int o = 0; o < 4; ++o) rk += (key[o] < key[q]) ? 1 : 0;
                if (rk == k && (uint32_t)q < n) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        acc[v].x += G[q][v].x; acc[v].y += G[q][v].y; acc[v].z += G[q][v].z; acc[v].w += G[q][v].w;
                    }
                }
            }
        }
        return;
    }
    // Up to LANES occurrences (a shard owner's inbox at 8 GPUs averages ~8 per row): lane q of the group holds key q, the group ranks the
    // keys with shuffles (keys are unique), and the gradient rows are then requested in ascending key order with addresses that do not
    // depend on earlier loads -- two memory round trips for the whole row instead of one per occurrence, same summation order.
    if (ordered && n > 4u && n <= (uint32_t)LANES) {
        const int lane = threadIdx.x & 31;
        const uint32_t gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << (lane - gl));
        const bool on = (uint32_t)gl < n;
        const uint32_t key = on ? dup_t[lo + gl] : 0xFFFFFFFFu;
        const uint32_t myslot = on ? (dup_src ? dup_src[lo + gl] : lo + gl) : 0u;
        uint32_t rk = 0;
        for (uint32_t o = 0; o < n; ++o) rk += (__shfl_sync(gmask, key, (int)o, LANES) < key) ? 1u : 0u;
        int src = 0;   // lane (inside the group) whose key has rank gl
        for (uint32_t o = 0; o < n; ++o) src = (__shfl_sync(gmask, rk, (int)o, LANES) == (uint32_t)gl) ? (int)o : src;
        const uint32_t ordslot = __shfl_sync(gmask, myslot, src, LANES);
#pragma unroll 4
        for (uint32_t k = 0; k < n; ++k) {
            const int64_t slot = __shfl_sync(gmask, ordslot, (int)k, LANES);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c = (gl + LANES * v) * 4;
                if (c < dim) {
                    const float4 g = ld4(dup_grad + slot * dim + c);
                    acc[v].x += g.x; acc[v].y += g.y; acc[v].z += g.z; acc[v].w += g.w;
                }
            }
        }
        return;
    }
    if (ordered) {
        int64_t prev = -1;
        for (uint32_t k = lo; k < hi; ++k) {
            int64_t best = 0x7fffffffffffLL;
            uint32_t bs = lo;
            for (uint32_t q = lo; q < hi; ++q) {
                int64_t tq = (int64_t)dup_t[q];
                if (tq > prev && tq < best) { best = tq; bs = q; }
            }
            prev = best;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                int c = (gl + LANES * v) * 4;
                if (c < dim) {
                    float4 g = ld4(dup_grad + (int64_t)(dup_src ? dup_src[bs] : bs) * dim + c);
                    acc[v].x += g.x; acc[v].y += g.y; acc[v].z += g.z; acc[v].w += g.w;
                }
            }
        }
    } else {
        for (uint32_t q = lo; q < hi; ++q) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                int c = (gl + LANES * v) * 4;
                if (c < dim) {
                    float4 g = ld4(dup_grad + (int64_t)(dup_src ? dup_src[q] : q) * dim + c);
                    acc[v].x += g.x; acc[v].y += g.y; acc[v].z += g.z; acc[v].w += g.w;
                }
            }
        }
    }
}

// The optimizer apply of one duplicate row is split in two so that the row's own loads (w, slots, last) are in flight while
// its gradient slots are being summed: dup_load issues them, dup_finish replays / applies / stores.
template <int LANES, int VPL, int OPT>
__device__ __forceinline__ void dup_load(RowRegs<LANES, VPL>& r, const DupArgs& a, const crb_dup_row& d, int gl) {
    const TableDev& T = a.tab[d.table];
    row_load_w<LANES, VPL>(r, T, d.row, a.dim, gl);
    r.last = OptTraits<OPT>::replay ? T.last[d.row] : 0;
    row_load_state<LANES, VPL, OPT>(r, T, d.row, a.dim, gl);
}

template <int LANES, int VPL, int OPT>
__device__ __forceinline__ void dup_finish(RowRegs<LANES, VPL>& r, const DupArgs& a, const OptDev& o, const crb_dup_row& d, const float4* acc,
                                           int gl) {
    const TableDev& T = a.tab[d.table];
    row_replay<LANES, VPL, OPT>(r, o, o.step);
    row_apply_store<LANES, VPL, OPT>(r, acc, T, d.row, a.dim, gl, o);
    if (gl == 0) a.meta[d.table][d.row] = 0ULL;
}

template <int LANES, int VPL, int OPT>
__device__ __forceinline__ void dup_apply(const DupArgs& a, const OptDev& o, const crb_dup_row& d, const float4* acc, int gl) {
    RowRegs<LANES, VPL> r;
    dup_load<LANES, VPL, OPT>(r, a, d, gl);
    dup_finish<LANES, VPL, OPT>(r, a, o, d, acc, gl);
}

// One lane group per work item (= one duplicate row, or one 256-slot chunk of a very frequent one).  The loop is software
// pipelined: the next item's descriptors (work -> dup_rows, two dependent loads) are fetched while the current item's slots
// and row are in flight, so a group's critical path per item is one memory round trip instead of four.
// PDL: the kernel was launched with programmatic stream serialisation behind the step kernel (dup_tail_kernel): what K2 left (counters,
// work list, row descriptors) is read at once, `griddepcontrol.wait` comes before the first access to anything the step kernel writes.
template <int LANES, int VPL, int OPT, bool SHARD, bool PDL = false>
__device__ __forceinline__ void dup_reduce_body(const DupArgs& a, const OptDev& o) {
    const int gl = threadIdx.x % LANES;
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / LANES;
    const int64_t n_work = a.ctr->work_items;
    int64_t k = group;
    if (k >= n_work) return;
    crb_work w = a.work[k];
    crb_dup_row d = a.dup_rows[w.dup];
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
    while (true) {
        const int64_t kn = k + n_groups;
        const bool more = kn < n_work;
        crb_work wn = w;
        if (more) wn = a.work[kn];
        const uint32_t lo = d.base + w.chunk * CRB_DUP_CHUNK;
        const uint32_t hi = min(d.base + d.cnt, lo + CRB_DUP_CHUNK);
        RowRegs<LANES, VPL> r;
        const bool send = SHARD && d.table == 1;   // multi-GPU: an item row's sum goes to its owner instead of being applied here
        if (d.nchunk == 1 && !send) dup_load<LANES, VPL, OPT>(r, a, d, gl);
        float4 acc[VPL];
        sum_slots<LANES, VPL>(acc, a.dup_src ? a.src_grad : a.dup_grad, a.dup_t, a.dup_src, lo, hi, d.cnt <= 32u, a.dim, gl);
        crb_dup_row dn = d;
        if (more) dn = a.dup_rows[wn.dup];
        if (d.nchunk == 1 && send) {
            shard_send<LANES, VPL>(a.send, d.row, acc, a.dim, gl);
            if (gl == 0) a.meta[1][d.row] = 0ULL;
        } else if (d.nchunk == 1) {
            dup_finish<LANES, VPL, OPT>(r, a, o, d, acc, gl);
        } else {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                int c = (gl + LANES * v) * 4;
                if (c < a.dim) st4(a.partial + (int64_t)(d.pbase + w.chunk) * a.dim + c, acc[v]);
            }
        }
        if (!more) break;
        k = kn; w = wn; d = dn;
    }
}

template <int LANES, int VPL, int OPT, bool SHARD = false>
__global__ void __launch_bounds__(256) dup_reduce_kernel(const DupArgs a) {
    OptDev o = a.opt;   // a resolved COPY: writing into the parameter struct would move all of it to local memory
    opt_resolve(o);
    dup_reduce_body<LANES, VPL, OPT, SHARD>(a, o);
}

// multi-chunk rows: lane groups [group, group + n_groups, ...) of the launch (or of ONE block, dup_tail_kernel)
template <int LANES, int VPL, int OPT, bool SHARD>
__device__ __forceinline__ void dup_final_body(const DupArgs& a, const OptDev& o, int64_t group, int64_t n_groups) {
    const int gl = threadIdx.x % LANES;
    const uint32_t n_multi = a.ctr->multi_rows;
    for (int64_t k = group; k < n_multi; k += n_groups) {
        const crb_dup_row d = a.dup_rows[a.multi[k]];
        float4 acc[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t q = 0; q < d.nchunk; ++q) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                int c = (gl + LANES * v) * 4;
                if (c < a.dim) {
                    float4 g = ld4(a.partial + (int64_t)(d.pbase + q) * a.dim + c);
                    acc[v].x += g.x; acc[v].y += g.y; acc[v].z += g.z; acc[v].w += g.w;
                }
            }
        }
        if (SHARD && d.table == 1) {
            shard_send<LANES, VPL>(a.send, d.row, acc, a.dim, gl);
            if (gl == 0) a.meta[1][d.row] = 0ULL;
        } else {
            dup_apply<LANES, VPL, OPT>(a, o, d, acc, gl);
        }
    }
}

template <int LANES, int VPL, int OPT, bool SHARD = false>
__global__ void __launch_bounds__(256) dup_final_kernel(const DupArgs a) {
    OptDev o = a.opt;
    opt_resolve(o);
    dup_final_body<LANES, VPL, OPT, SHARD>(a, o, ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES, (int64_t)gridDim.x * blockDim.x / LANES);
}

// Small batches (the shapes the reference ships: a step is a handful of microsecond kernels and costs launch count x launch latency):
// K4 and K5 and the reset of the step counters in ONE launch.  Every block does its share of dup_reduce; the block that finishes last
// (ticket in crb_step_ctr::tail_done, the threadFenceReduction pattern) then finishes the multi-chunk rows -- there are none unless a
// row occurs more than CRB_DUP_CHUNK times in the batch --, sums the per-block loss partials in loss_final_kernel's fixed order and
// zeroes the counters for the step that uses this counter set next.  Same arithmetic per row and the same loss bits as the three
// separate launches.
struct DupTail {
    crb_step_ctr* ctr;
    const double* block_loss;
    int n_block_loss;
    double* loss_out;
};

template <int LANES, int VPL, int OPT>
__global__ void __launch_bounds__(256) dup_tail_kernel(const DupArgs a, const DupTail t) {
    OptDev o = a.opt;
    opt_resolve(o);
    const uint32_t n_multi = a.ctr->multi_rows;   // final since K2; same sector as work_items, which dup_reduce_body reads next
    // K5: the per-block loss partials are K3's, so any block may sum them at any time: the last block of the grid does it first (its
    // lane groups are the ones without a work item when there are fewer duplicate rows than groups), in loss_final_kernel's order
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x < 32) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        double v = 0.0;
        for (int k = threadIdx.x; k < t.n_block_loss; k += 32) v += t.block_loss[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (threadIdx.x == 0) *t.loss_out = v;
    }
    dup_reduce_body<LANES, VPL, OPT, false, true>(a, o);
    __shared__ bool s_last;
    if (n_multi) __threadfence();   // partial sums of multi-chunk rows must be visible to the block that finishes them
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&t.ctr->tail_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if (n_multi) {
        asm volatile("griddepcontrol.wait;" ::: "memory");   // threads without a work item have not waited yet
        __threadfence();
        dup_final_body<LANES, VPL, OPT, false>(a, o, threadIdx.x / LANES, blockDim.x / LANES);
        __syncthreads();   // every thread of the block has read ctr->multi_rows
    }
    // every block read the counters before it took its ticket: they can be zeroed for the step that uses this copy next
    if (threadIdx.x == 0) {
        t.ctr->dup_slots = 0; t.ctr->dup_rows = 0; t.ctr->work_items = 0; t.ctr->multi_rows = 0; t.ctr->partial_slots = 0;
        t.ctr->tail_done = 0;
    }
}
