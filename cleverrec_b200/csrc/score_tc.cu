// Full-rank scoring on the 5th-generation tensor cores (reference: tf.matmul(u_embed, Q, transpose_b=True) + np.argsort +
// seen filter, model/ranking/BPR.py:51 and model/RankingRecommender.py:221-240).
//
//   prep_*_kernel     fp32 tables -> bf16 operand copies (K padded to 64; bias / -|q|^2 folded in as two extra K columns)
//   score_tc_kernel   persistent, warp-specialised: warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle), warp 1 = one
//                     elected thread issuing tcgen05.mma (M=128 x N=128 x K=16, two M halves per CTA = 256 users share every
//                     item tile), accumulators double-buffered in TMEM (2 x 256 columns), warps 4-11 = epilogue: tcgen05.ld,
//                     seen items masked with a per-row cursor into the user's sorted history, running per-row threshold,
//                     candidates appended to a per-row list; the score matrix never reaches HBM.
//   rescore_kernel    canonical fp32 re-scoring of each user's candidates + certificate: every unlisted item has
//                     approx <= theta, |approx - canonical| <= eps, so if the K-th canonical score beats theta + eps the
//                     returned ids are exactly those of the fp32 path.  Uncertified users go to fullrank_exact_kernel.
#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_pipeline.h>

#include "score_common.cuh"

#define TC_C 256          // candidate slots per (user, split, column half)
#define TC_EPL (TC_C / 32) // list entries per lane in a compaction
#define TC_BM 256         // users per CTA (2 x UMMA_M=128)
#define TC_BN 128         // items per tile
#ifndef TC_CH
#define TC_CH 2            // column halves of a tile, each scanned by its own set of 8 warps with its own candidate lists
#endif
#define TC_THREADS (128 + 256 * TC_CH)   // 4 control warps + 8 epilogue warps per column half
#define TC_MAX_KB 3       // K blocks of 64 -> d_pad <= 192

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"  // sleeps in hardware up to the hint (ns)
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same MMA with the two shared-memory descriptors given as (low word, common high word): the high word (SBO, version, swizzle
// mode) never changes and the low word (start address >> 4, LBO) moves by a compile-time constant from one MMA to the next, so
// the issuing loop is one add per operand instead of rebuilding two 64-bit descriptors (21 SASS instructions per MMA before).
__device__ __forceinline__ void umma_bf16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMEM load of 32 consecutive fp32 columns of this thread's lane, and the wait that makes the registers valid (one asm block:
// the compiler must not schedule a use between the two).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]),
          "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
          "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr)
        : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4, LBO (unused
// for swizzled K-major) = 1, SBO = 1024 B (8 rows x 128 B) >> 4, version = 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major, N = 128, M = 128
#define TC_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))

// ------------------------------------------------------------------------------------------------ operand preparation
struct PrepArgs {
    const float* src;       // fp32 table
    const float* hvec;      // GMF h / item bias
    const int32_t* rows;    // row ids to take (users) or NULL for 0..n-1
    int64_t n, n_pad;
    int dim, d_pad, kind;
    __nv_bfloat16* dst;     // [n_pad, d_pad]
    float* norm;            // users: |p| (or |p.h|);   items: unused
    float* aux;             // users (SQDIST): sum p^2 in fp32
    unsigned int* maxbits;  // items: max |q| as float bits; [1]: max |bias|
};

// one warp per row.  is_user selects the A-operand rules.  Rows are read 16 bytes and written 8 bytes per lane when the table allows it
// (dim % 4 == 0, 16-byte aligned base: every table the library trains); the item-side maxima are kept per warp and merged with ONE atomic
// per warp at the end -- one atomicMax per row on a single address serialised in L2 and held the 2M-row conversion at 0.8 TB/s.
template <bool IS_USER>
__global__ void __launch_bounds__(256) prep_kernel(PrepArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool vec = (a.dim & 3) == 0 && ((uintptr_t)a.src & 15) == 0 && (!(IS_USER && a.kind == CRB_SCORE_GMF) || ((uintptr_t)a.hvec & 15) == 0);
    float max_norm = 0.f, max_bias = 0.f;
    for (int64_t r = warp; r < a.n_pad; r += n_warps) {
        __nv_bfloat16* out = a.dst + r * a.d_pad;
        if (r >= a.n) {
            for (int k = lane; k < a.d_pad; k += 32) out[k] = __float2bfloat16(0.f);
            continue;
        }
        const float* src = a.src + (int64_t)(a.rows ? a.rows[r] : r) * a.dim;
        float sq = 0.f;
        if (vec) {
            for (int k = lane * 4; k < a.dim; k += 128) {
                float4 x = *reinterpret_cast<const float4*>(src + k);
                if (IS_USER && a.kind == CRB_SCORE_GMF) {
                    const float4 hh = *reinterpret_cast<const float4*>(a.hvec + k);
                    x.x = __fmul_rn(x.x, hh.x); x.y = __fmul_rn(x.y, hh.y); x.z = __fmul_rn(x.z, hh.z); x.w = __fmul_rn(x.w, hh.w);
                }
                sq = fmaf(x.x, x.x, sq); sq = fmaf(x.y, x.y, sq); sq = fmaf(x.z, x.z, sq); sq = fmaf(x.w, x.w, sq);
                if (!IS_USER && a.kind == CRB_SCORE_SQDIST) { x.x *= 2.f; x.y *= 2.f; x.z *= 2.f; x.w *= 2.f; }
                const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(out + k) = pk;
            }
        } else {
            for (int k = lane; k < a.dim; k += 32) {
                float x = src[k];
                if (IS_USER && a.kind == CRB_SCORE_GMF) x = __fmul_rn(x, a.hvec[k]);
                sq = fmaf(x, x, sq);
                if (!IS_USER && a.kind == CRB_SCORE_SQDIST) x = 2.f * x;
                out[k] = __float2bfloat16(x);
            }
        }
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        // two augmentation columns carry a per-item fp32 constant as hi + lo bf16 parts (A side holds 1, 1)
        const bool aug = a.kind == CRB_SCORE_SQDIST || a.kind == CRB_SCORE_DOT_BIAS;
        for (int k = a.dim + lane; k < a.d_pad; k += 32) {
            float x = 0.f;
            if (aug && k < a.dim + 2) {
                if (IS_USER) {
                    x = 1.f;
                } else {
                    const float c = a.kind == CRB_SCORE_SQDIST ? -sq : a.hvec[r];
                    const float hi = __bfloat162float(__float2bfloat16(c));
                    x = (k == a.dim) ? hi : (c - hi);
                }
            }
            out[k] = __float2bfloat16(x);
        }
        if (IS_USER) {
            if (lane == 0) {
                a.norm[r] = sqrtf(sq);
                if (a.aux) a.aux[r] = sq;
            }
        } else {
            max_norm = fmaxf(max_norm, sqrtf(sq));
            if (a.kind == CRB_SCORE_DOT_BIAS) max_bias = fmaxf(max_bias, fabsf(a.hvec[r]));
        }
    }
    if (!IS_USER && lane == 0) {   // non-negative floats order like their bit patterns
        atomicMax(a.maxbits, __float_as_uint(max_norm));
        if (a.kind == CRB_SCORE_DOT_BIAS) atomicMax(a.maxbits + 1, __float_as_uint(max_bias));
    }
}

// ------------------------------------------------------------------------------------------------ main kernel
struct TcArgs {
    int64_t n_users;        // real users in this pass
    int64_t n_users_pad;    // multiple of TC_BM
    int64_t n_items;        // real items
    int n_tiles;            // item tiles (n_items_pad / TC_BN)
    int n_splits, tiles_per_split;
    int kb;                 // K blocks of 64
    int stages;             // smem stages for B
    const int32_t* users;       // [n_users] row ids (history lookup when hist_users == NULL)
    const int32_t* hist_users;  // or NULL
    const int64_t* seen_rowptr;
    const int32_t* seen_cols;
    unsigned long long* cand;   // [n_splits, n_users_pad, TC_C]  (approx score bits << 32 | item)
    int32_t* cand_cnt;          // [n_splits, n_users_pad]
    float* cand_thr;            // [n_splits, n_users_pad]
    int keep_lo, keep_hi;       // a compaction keeps between keep_lo and keep_hi entries (keep_lo >= K)
    int trig;                   // a list longer than this is compacted after the tile (keep_hi < trig <= TC_C - 64)
    unsigned long long* dbg;    // experiments only (-DTC_DEBUG_SWITCHES, CRB_TC_DEBUG_CYCLES=1): cycle counters, see score_topk_tc
    int debug;                  // experiments only (-DTC_DEBUG_SWITCHES, CRB_TC_DEBUG): 1 = read TMEM but skip the scan, 2 = skip the TMEM read too, 3 = scan only
};

__device__ __forceinline__ uint32_t ord_bits(uint32_t f) { return (f & 0x80000000u) ? ~f : (f | 0x80000000u); }

// Warp-cooperative compaction of one row's candidate list (n <= TC_C entries, TC_EPL per lane).  ANY threshold T is valid for the
// certificate as long as every dropped entry has score <= T; the list only needs to shrink enough to make room.  So instead of
// an exact selection (a 32-step radix descent was the straggler that stalled the 2-deep accumulator pipeline) T is found by
// bisection on the order-preserving score bits between the list's min and max, stopping as soon as the number of kept entries
// (score > T) lies in [keep_lo, keep_hi]: typically 4-6 warp reductions.  Kept entries are packed to the front.
// Returns T as a float; *kept receives the new count.
// The threshold only moves at a compaction, and until the next one the list takes every score above it: the tighter the kept range
// and the earlier the next compaction, the fewer scores pass the threshold test at all -- and a chunk in which ANY of a warp's 32
// users has a passing score leaves the branch-free path for the whole warp.  Measured (4096 users x 2M items, one item split): with
// keep in [32, 64] and a compaction at 192 entries the candidate path cost 3.8 of 16.2 ms; see score_topk_tc for the choice now.
__device__ __forceinline__ float compact_list(unsigned long long* list, int n, int lane, int* kept, int keep_lo, int keep_hi) {
    unsigned long long e[TC_EPL];
    uint32_t o[TC_EPL];
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
#pragma unroll
    for (int t = 0; t < TC_EPL; ++t) {
        const int idx = lane + 32 * t;
        e[t] = idx < n ? __ldcg(list + idx) : 0ULL;
        o[t] = idx < n ? ord_bits((uint32_t)(e[t] >> 32)) : 0u;
        if (idx < n) { lo = min(lo, o[t]); hi = max(hi, o[t]); }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    // invariants: count(o > hi) <= keep_hi (0 at the start);  count(o > lo - 1) would be n (too many)
    uint32_t T = hi;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        int c = 0;
#pragma unroll
        for (int t = 0; t < TC_EPL; ++t) c += (o[t] > mid);
        const int tot = __reduce_add_sync(0xffffffffu, c);
        if (tot > keep_hi) { lo = mid + 1; T = hi; }
        else { hi = mid; T = mid; if (tot >= keep_lo) break; }
    }
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int t = 0; t < TC_EPL; ++t) {
        const bool keep = o[t] > T;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) __stcg(list + base + __popc(m & ((1u << lane) - 1u)), e[t]);
        base += __popc(m);
    }
    __syncwarp();
    *kept = base;
    const uint32_t f = (T & 0x80000000u) ? (T & 0x7fffffffu) : ~T;  // inverse of ord_bits
    return __uint_as_float(f);
}

__global__ void __launch_bounds__(TC_THREADS, 1) score_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                  const __grid_constant__ CUtensorMap map_b, TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    // 128-byte swizzle atoms need 1024-byte aligned operand tiles: align by hand (the launch reserves the slack)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_bytes = 2u * a.kb * 16384u;  // 2 halves x kb boxes of [128 rows x 128 B]
    const uint32_t b_bytes = (uint32_t)a.kb * 16384u;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)a.stages * b_bytes);
    uint64_t* b_full = bars;                 // [stages]
    uint64_t* b_empty = bars + 4;            // [stages]
    uint64_t* a_full = bars + 8;
    uint64_t* a_empty = bars + 9;
    // accumulator hand-off barriers, one pair per (stage, M half): a half's 8 epilogue warps start as soon as ITS 8 MMAs have
    // retired and release it independently, so the MMA pipe stays busy as long as epilogue + hand-off latency <= 1.5 tile times
    // (with whole-tile barriers the bound was 1.0 and the two sides ping-ponged: each waited on the other ~30% of the time)
    uint64_t* t_full = bars + 10;            // [2 stages][2 halves]
    uint64_t* t_empty = bars + 14;           // [2 stages][2 halves]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < 4; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 4 * TC_CH); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int64_t m_tiles = a.n_users_pad / TC_BM;
    const int64_t n_work = m_tiles * a.n_splits;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, a_phase = 0;
            for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int64_t mt = w / a.n_splits;
                const int sp = (int)(w % a.n_splits);
                const int t0 = sp * a.tiles_per_split, t1 = min(a.n_tiles, t0 + a.tiles_per_split);
                mbar_wait(a_empty, a_phase ^ 1);
                mbar_expect_tx(a_full, a_bytes);
                for (int h = 0; h < 2; ++h)
                    for (int kb = 0; kb < a.kb; ++kb)
                        tma_load_2d(smem_a + (size_t)(h * a.kb + kb) * 16384, &map_a, kb * 64, (int)(mt * TC_BM + h * 128), a_full);
                a_phase ^= 1;
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(b_empty + stage, phase ^ 1);
                    mbar_expect_tx(b_full + stage, b_bytes);
                    for (int kb = 0; kb < a.kb; ++kb)
                        tma_load_2d(smem_b + (size_t)stage * b_bytes + (size_t)kb * 16384, &map_b, kb * 64, t * TC_BN, b_full + stage);
                    if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp runs the loop (so that everything below is warp-uniform and lives in uniform registers); only the
        // tcgen05 instructions themselves are issued by one elected lane.
        const bool leader = elect_one();
        uint32_t stage = 0, phase = 0, a_phase = 0, acc = 0, acc_phase = 0;
        const uint32_t desc_hi = (uint32_t)(umma_desc(0) >> 32);
        const uint32_t a_lo0 = (uint32_t)umma_desc(smem_u32(smem_a)), b_lo0 = (uint32_t)umma_desc(smem_u32(smem_b));
        const uint32_t b_stage_step = b_bytes >> 4;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int sp = (int)(w % a.n_splits);
            const int t0 = sp * a.tiles_per_split, t1 = min(a.n_tiles, t0 + a.tiles_per_split);
            mbar_wait(a_full, a_phase);
            a_phase ^= 1;
#ifdef TC_DEBUG_SWITCHES
            long long dbg_b = 0, dbg_t = 0, dbg_c = 0; const long long dbg_start = clock64();
#define DBG_T0 dbg_c = clock64();
#define DBG_ADD(x) x += clock64() - dbg_c;
#else
#define DBG_T0
#define DBG_ADD(x)
#endif
            for (int t = t0; t < t1; ++t) {
                DBG_T0
                mbar_wait(b_full + stage, phase);          // TMA bytes have landed
                DBG_ADD(dbg_b)
                const uint32_t b_lo = b_lo0 + stage * b_stage_step;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    DBG_T0
                    mbar_wait(t_empty + acc * 2 + h, acc_phase ^ 1);   // this half's epilogue warps have drained the stage
                    DBG_ADD(dbg_t)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (leader) {
                        const uint32_t d_tmem = tmem_base + acc * 256u + (uint32_t)h * 128u;
                        const uint32_t a_lo = a_lo0 + (uint32_t)h * (uint32_t)a.kb * 1024u;   // 16384 B per K block, >> 4
                        for (int kb = 0; kb < a.kb; ++kb) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)  // UMMA_K = 16 bf16 = 32 bytes (>> 4 = 2) inside the 128-byte swizzle atom
                                umma_bf16_lh(d_tmem, a_lo + kb * 1024 + k * 2, b_lo + kb * 1024 + k * 2, desc_hi, TC_IDESC, (kb | k) ? 1u : 0u);
                        }
                        if (h == 1) umma_commit(b_empty + stage);  // smem stage reusable once these MMAs retire
                        umma_commit(t_full + acc * 2 + h);         // this half's accumulator is ready for its epilogue warps
                    }
                    __syncwarp();
                }
                if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
                if (++acc == 2u) { acc = 0; acc_phase ^= 1; }
            }
            if (leader) umma_commit(a_empty);  // A tile reusable after the last item tile's MMAs
            __syncwarp();
#ifdef TC_DEBUG_SWITCHES
            if (a.dbg && lane == 0) { atomicAdd(a.dbg + 0, (unsigned long long)dbg_b); atomicAdd(a.dbg + 1, (unsigned long long)dbg_t); atomicAdd(a.dbg + 5, (unsigned long long)(clock64() - dbg_start)); }
#endif
        }
    } else if (warp >= 4) {
        // ================= epilogue: 16 warps; a thread owns one TMEM lane (= one user) and one column half of every tile ======
        // This loop is instruction-issue bound (profiles/r01_tc_epilogue_decomposition.md): every instruction below is paid
        // 32768 times per 128-item tile and SM.  The common path per 32 scores is: TMEM load, one compare for "anything special
        // in this chunk" (seen item or catalogue end), a 3-input max tree, one compare against the row's threshold.
        const int ew = warp - 4, half = (ew >> 2) & 1, quad = warp & 3, ch = ew >> 3;  // a warp may only touch TMEM lanes 32*(warp%4)..+31
        const int row = half * 128 + quad * 32 + lane;
        const int32_t n_items32 = (int32_t)a.n_items;
        uint32_t acc = 0, acc_phase = 0;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int64_t mt = w / a.n_splits;
            const int sp = (int)(w % a.n_splits);
            const int t0 = sp * a.tiles_per_split, t1 = min(a.n_tiles, t0 + a.tiles_per_split);
            const int64_t g = mt * TC_BM + row;  // user slot of this thread
            const bool live = g < a.n_users;
            const int64_t lslot = (int64_t)(sp * TC_CH + ch) * a.n_users_pad;   // this (split, column half)'s lists
            unsigned long long* list = a.cand + (lslot + g) * TC_C;
#ifdef TC_DEBUG_SWITCHES
            float theta = (!live || a.debug == 3) ? INFINITY : -INFINITY;   // debug 3: fast path only
#else
            float theta = live ? -INFINITY : INFINITY;   // a padding row never lists anything
#endif
            int cnt = 0;
            // cursor into the user's sorted history: first seen item >= first item of this split
            int64_t hp = 0, hend = 0;
            int32_t seen_cur = 0x7fffffff;
            if (live) {
                const int32_t hu = a.hist_users ? a.hist_users[g] : a.users[g];
                if (hu >= 0) {
                    int64_t lo = a.seen_rowptr[hu], hi = a.seen_rowptr[hu + 1];
                    hend = hi;
                    const int32_t first = t0 * TC_BN;
                    while (lo < hi) {
                        const int64_t mid = (lo + hi) >> 1;
                        if (__ldg(a.seen_cols + mid) < first) lo = mid + 1; else hi = mid;
                    }
                    hp = lo;
                    if (hp < hend) seen_cur = __ldg(a.seen_cols + hp);
                }
            }
            // first column at which a chunk needs the special path: the next seen item or the end of the catalogue
            int32_t special_at = live ? min(seen_cur, n_items32) : 0x7fffffff;
#ifdef TC_DEBUG_SWITCHES
            long long dbg_w = 0, dbg_s = 0, dbg_k = 0, dbg_c = 0, dbg_n = 0;
#endif
            for (int t = t0; t < t1; ++t) {
                DBG_T0
                mbar_wait(t_full + acc * 2 + half, acc_phase);
                DBG_ADD(dbg_w)
                DBG_T0
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 256u + (uint32_t)half * 128u;
#pragma unroll 1
                for (int c = ch * (TC_BN / 32 / TC_CH); c < (ch + 1) * (TC_BN / 32 / TC_CH); ++c) {
                    float v[32];
#ifdef TC_DEBUG_SWITCHES
                    if (a.debug == 2) continue;
#endif
                    tmem_ld32(t_lane + (uint32_t)c * 32u, v);
#ifdef TC_DEBUG_SWITCHES
                    if (a.debug == 1) { if (v[0] == 1.2345e30f && v[31] == 5.4321e30f) cnt = 0; continue; }
#endif
                    const int32_t c0 = t * TC_BN + c * 32;
                    uint32_t skip = 0;   // bit k: column c0 + k must not be listed (seen by the user, or past the catalogue)
                    if (special_at < c0 + 32) {
                        // rare: seen items (amortised O(|history|) per user) and the padding past the catalogue.  The scores stay
                        // where they are (masking one of 32 registers by a run-time index costs 64 instructions); the bit mask
                        // is applied where candidates are appended.  A seen item can only make the chunk take the slow path.
                        while (seen_cur < c0 + 32) {
                            const int idx = seen_cur - c0;
                            if (idx >= 0) skip |= 1u << idx;
                            ++hp;
                            seen_cur = hp < hend ? __ldg(a.seen_cols + hp) : 0x7fffffff;
                        }
                        if (c0 + 32 > n_items32) skip |= c0 >= n_items32 ? 0xffffffffu : (0xffffffffu << (n_items32 - c0));
                        special_at = min(seen_cur, n_items32);   // past the catalogue every chunk stays special
                    }
                    // 8 group maxima (3-input max) -> row maximum; only groups that beat the threshold are scanned
                    float gm[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) gm[q] = fmaxf(fmaxf(v[4 * q], v[4 * q + 1]), fmaxf(v[4 * q + 2], v[4 * q + 3]));
                    const float mx = fmaxf(fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3])), fmaxf(fmaxf(gm[4], gm[5]), fmaxf(gm[6], gm[7])));
                    if (mx > theta) {
                        {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                if (gm[q] > theta) {
                                    // bit k: element k of the group passes.  The loop's trip count is data dependent, which keeps
                                    // the compiler from predicating four store sequences per group (110 instructions per hit chunk).
                                    unsigned m = ((v[4 * q] > theta ? 1u : 0u) | (v[4 * q + 1] > theta ? 2u : 0u) |
                                                  (v[4 * q + 2] > theta ? 4u : 0u) | (v[4 * q + 3] > theta ? 8u : 0u)) & ~(skip >> (4 * q));
                                    while (m) {
                                        const int k = __ffs(m) - 1;
                                        m &= m - 1;
                                        const float sc = k == 0 ? v[4 * q] : k == 1 ? v[4 * q + 1] : k == 2 ? v[4 * q + 2] : v[4 * q + 3];
                                        __stcg(list + cnt, ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned long long)(uint32_t)(c0 + 4 * q + k));
                                        ++cnt;
                                    }
                                }
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                DBG_ADD(dbg_s)
                if (lane == 0) mbar_arrive(t_empty + acc * 2 + half);
                if (++acc == 2u) { acc = 0; acc_phase ^= 1; }
                // Compaction is OFF the accumulator hand-off: the TMEM stage has just been released, so the ~1500 cycles a
                // compaction takes (list round trip through L2 + bisection) overlap the next tiles' MMAs instead of stalling them.
                // A list is compacted as soon as the next tile could overflow it (worst case 32 appends per chunk).
                unsigned need = __ballot_sync(0xffffffffu, cnt > a.trig);
                DBG_T0
#ifdef TC_DEBUG_SWITCHES
                dbg_n += __popc(need);
#endif
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const int n = __shfl_sync(0xffffffffu, cnt, src);
                    unsigned long long* lst = a.cand + (lslot + (mt * TC_BM + half * 128 + quad * 32 + src)) * TC_C;
                    int kept;
                    const float thr = compact_list(lst, n, lane, &kept, a.keep_lo, a.keep_hi);
                    if (lane == src) { theta = fmaxf(theta, thr); cnt = kept; }
                }
                DBG_ADD(dbg_k)
            }
#ifdef TC_DEBUG_SWITCHES
            if (a.dbg && lane == 0) { atomicAdd(a.dbg + 2, (unsigned long long)dbg_w); atomicAdd(a.dbg + 3, (unsigned long long)dbg_s); atomicAdd(a.dbg + 4, (unsigned long long)dbg_k); atomicAdd(a.dbg + 6, (unsigned long long)dbg_n); }
#endif
            a.cand_cnt[lslot + g] = live ? cnt : 0;
            a.cand_thr[lslot + g] = theta;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ certified re-scoring
struct RescoreArgs {
    int kind, dim, K, n_splits;
    int64_t n_users, n_users_pad;
    const float* P;
    const float* Q;
    const float* hvec;
    const int32_t* users;
    unsigned long long* cand;
    const int32_t* cand_cnt;
    const float* cand_thr;
    const float* pnorm;
    const float* psq;
    const unsigned int* maxbits;
    float cbound;
    int32_t* out_items;
    float* out_scores;
    int32_t* todo;
    unsigned int* counters;  // [0] uncertified users, [1] max candidates
    int lists_only;          // 1 = every user on rescore_user_lists (CRB_RESCORE_LISTS=1: A/B and the equality test)
};

#define RS_CH 64   // columns of the item rows staged per round (shared memory per warp: 32 rows x 65 + 64 floats + the list counts)
#define RS_ORD (32 * (RS_CH + 1))   // words of a warp's row staging area; before any row is staged it holds the bisection array
#define RS_MAXL (64 * TC_CH)        // candidate lists per user (item splits x column halves)
#define RS_LIVE 8                   // surviving candidates per lane kept in registers (256 per user); a user with more takes the list path
#define RS_WARP_WORDS (RS_ORD + RS_CH + RS_MAXL)

// Canonical scores of up to 32 candidates of one user (lane = candidate): a lane runs the canonical sequential chain (the scores are those
// of canonical_score, bit for bit), but the item rows are staged in shared memory by cp.async, 64 columns of up to 32 rows at a time,
// instead of being read 16 bytes at a time by the lane that consumes them (same scheme as score_pairs_tiled_kernel).
template <int KIND>
__device__ __forceinline__ float rescore_group(const RescoreArgs& a, const float* __restrict__ p, uint32_t item, bool live, unsigned live_mask,
                                               int lane, float* sQ, float* sP) {
    float acc = 0.f;
    for (int kc = 0; kc < a.dim; kc += RS_CH) {
        const int len = a.dim - kc < RS_CH ? a.dim - kc : RS_CH;
        const bool in0 = lane < len, in1 = lane + 32 < len;
        for (unsigned m = live_mask; m; m &= m - 1) {
            const int r = __ffs(m) - 1;
            const float* src = a.Q + (int64_t)__shfl_sync(0xffffffffu, item, r) * a.dim + kc;
            if (in0) __pipeline_memcpy_async(sQ + r * (RS_CH + 1) + lane, src + lane, 4);
            if (in1) __pipeline_memcpy_async(sQ + r * (RS_CH + 1) + lane + 32, src + lane + 32, 4);
        }
        __pipeline_commit();
        if (in0) sP[lane] = p[kc + lane];
        if (in1) sP[lane + 32] = p[kc + lane + 32];
        __pipeline_wait_prior(0);
        __syncwarp();
        if (live) {
            const float* q = sQ + lane * (RS_CH + 1);
            for (int c = 0; c < len; ++c) {
                const float pa = sP[c], qb = q[c];
                if (KIND == CRB_SCORE_DOT || KIND == CRB_SCORE_DOT_BIAS) acc = fmaf(pa, qb, acc);
                else if (KIND == CRB_SCORE_GMF) acc = fmaf(__fmul_rn(pa, qb), __ldg(a.hvec + kc + c), acc);
                else { const float dd = __fsub_rn(pa, qb); acc = fmaf(dd, dd, acc); }
            }
        }
        __syncwarp();
    }
    if (KIND == CRB_SCORE_DOT_BIAS) acc = __fadd_rn(acc, live ? a.hvec[item] : 0.f);
    return acc;
}

// One user, working on the candidate lists where they lie in global memory: any number of candidates.  The pre-filter bisection, the
// re-scoring and the K selection rounds each walk all of the user's lists, so the cost grows with (lists x rounds) dependent L2 round
// trips: this is the fall-back of rescore_user_fast, which handles every user with <= 256 survivors.
template <int KIND>
__device__ __noinline__ void rescore_user_lists(const RescoreArgs& a, int64_t g, int lane, float* sQ, float* sP, float eps, float* theta_out,
                                                unsigned long long* kth_out, int* total_out) {
    constexpr int ASC = KIND == CRB_SCORE_SQDIST ? 1 : 0;
    const float* p = a.P + (int64_t)a.users[g] * a.dim;
    float theta = -INFINITY;
    int total = 0;
    // Pre-filter on the APPROXIMATE scores: with T <= the K-th best approximate score of the user's candidates, a candidate
    // whose approximate score is below T - 2 eps has at least K candidates canonically above it (|approx - canonical| <= eps
    // for both), so it cannot be in the top K and needs no fp32 dot product.  T is found by bisection on the order-preserving
    // score bits (count >= K keeps T a lower bound; stop once the count is within [K, 2K]).  This does not touch theta.
    uint32_t lo_b = 0xFFFFFFFFu, hi_b = 0u;
    for (int sp = 0; sp < a.n_splits; ++sp) {
        const int64_t slot = (int64_t)sp * a.n_users_pad + g;
        const int n = a.cand_cnt[slot];
        const unsigned long long* list = a.cand + slot * TC_C;
        for (int k = lane; k < n; k += 32) {
            const uint32_t o = ord_bits((uint32_t)(list[k] >> 32));
            lo_b = min(lo_b, o); hi_b = max(hi_b, o);
        }
        total += n;
    }
    lo_b = __reduce_min_sync(0xffffffffu, lo_b);
    hi_b = __reduce_max_sync(0xffffffffu, hi_b);
    uint32_t T_b = lo_b;   // count(o >= lo_b) = total: a valid (if useless) lower bound when total >= K
    if (total > 2 * a.K) {
        uint32_t lo = lo_b, hi = hi_b;   // invariant: count(o >= lo) >= K
        for (int it = 0; it < 32 && lo < hi; ++it) {
            const uint32_t mid = lo + ((hi - lo + 1) >> 1);
            int c = 0;
            for (int sp = 0; sp < a.n_splits; ++sp) {
                const int64_t slot = (int64_t)sp * a.n_users_pad + g;
                const int n = a.cand_cnt[slot];
                const unsigned long long* list = a.cand + slot * TC_C;
                for (int k = lane; k < n; k += 32) c += ord_bits((uint32_t)(list[k] >> 32)) >= mid;
            }
            c = __reduce_add_sync(0xffffffffu, c);
            if (c >= a.K) { lo = mid; if (c <= 2 * a.K) break; } else hi = mid - 1;
        }
        T_b = lo;
    }
    const float T_f = __uint_as_float((T_b & 0x80000000u) ? (T_b & 0x7fffffffu) : ~T_b);
    const float cutoff = (total > 2 * a.K) ? T_f - 2.f * eps : -INFINITY;
    for (int sp = 0; sp < a.n_splits; ++sp) {
        const int64_t slot = (int64_t)sp * a.n_users_pad + g;
        const int n = a.cand_cnt[slot];
        theta = fmaxf(theta, a.cand_thr[slot]);
        unsigned long long* list = a.cand + slot * TC_C;
        // canonical score of every surviving candidate, entry rewritten as a ranking key (0 = filtered out)
        for (int k0 = 0; k0 < n; k0 += 32) {
            const int k = k0 + lane;
            const unsigned long long e = k < n ? list[k] : 0ULL;
            const uint32_t item = (uint32_t)e;
            const bool live = k < n && !(__uint_as_float((uint32_t)(e >> 32)) < cutoff);
            if (k < n && !live) list[k] = 0ULL;
            const unsigned live_mask = __ballot_sync(0xffffffffu, live);
            if (!live_mask) continue;
            const float acc = rescore_group<KIND>(a, p, item, live, live_mask, lane, sQ, sP);
            if (live) list[k] = rank_key(acc, item, ASC);
        }
    }
    __syncwarp();
    // K rounds of arg-best over all splits' lists
    unsigned long long prev = ~0ULL, kth = 0ULL;
    for (int r = 0; r < a.K; ++r) {
        unsigned long long best = 0ULL;
        for (int sp = 0; sp < a.n_splits; ++sp) {
            const int64_t slot = (int64_t)sp * a.n_users_pad + g;
            const int n = a.cand_cnt[slot];
            const unsigned long long* list = a.cand + slot * TC_C;
            for (int k = lane; k < n; k += 32) {
                const unsigned long long key = list[k];
                if (key < prev && key > best) best = key;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) {
            a.out_items[g * a.K + r] = best ? (int32_t)key_index(best) : -1;
            if (a.out_scores) a.out_scores[g * a.K + r] = best ? key_score(best, ASC) : 0.f;
        }
        if (best) { prev = best; kth = best; } else { kth = 0ULL; prev = 0ULL; }
    }
    *theta_out = theta; *kth_out = kth; *total_out = total;
}

// One user, everything after one pass over the lists in shared memory / registers.  The user's lists are read in batches of 8 loads per
// lane that do not depend on each other (the flat index runs over (list, position) pairs, whatever the lists' lengths); the bisection of
// the pre-filter runs on a shared-memory copy of the score bits -- up to RS_ORD / lists entries of every list: a threshold with >= K
// entries of a SUBSET at or above it has >= K of all entries at or above it, so the filter stays valid when a list is longer --; the
// survivors (<= 2K + the candidates within 2 eps of the threshold: a few dozen) are compacted, re-scored 32 at a time whatever list they
// came from, and ranked from registers.  Same survivors' canonical scores and same keys as rescore_user_lists, hence the same top K.
// Returns false (nothing written) when the user has more than 32 * RS_LIVE survivors.
template <int KIND>
__device__ __forceinline__ bool rescore_user_fast(const RescoreArgs& a, int64_t g, int lane, float* sQ, float* sP, int* sCnt, float eps,
                                                  float* theta_out, unsigned long long* kth_out, int* total_out) {
    constexpr int ASC = KIND == CRB_SCORE_SQDIST ? 1 : 0;
    const int L = a.n_splits;
    uint32_t* sO = reinterpret_cast<uint32_t*>(sQ);
    int total = 0, maxn = 0;
    float theta = -INFINITY;
    for (int l = lane; l < L; l += 32) {
        const int64_t slot = (int64_t)l * a.n_users_pad + g;
        const int n = a.cand_cnt[slot];
        sCnt[l] = n;
        total += n;
        maxn = max(maxn, n);
        theta = fmaxf(theta, a.cand_thr[slot]);
    }
    total = __reduce_add_sync(0xffffffffu, total);
    maxn = __reduce_max_sync(0xffffffffu, maxn);
    for (int o = 16; o > 0; o >>= 1) theta = fmaxf(theta, __shfl_xor_sync(0xffffffffu, theta, o));
    __syncwarp();
    float cutoff = -INFINITY;
    if (total > 2 * a.K) {
        const int cpl = min(maxn, RS_ORD / L);   // entries of every list that take part in the bisection
        const int space = L * cpl;
        uint32_t lo_b = 0xFFFFFFFFu, hi_b = 0u;
        int sub = 0;
        for (int f0 = 0; f0 < space; f0 += 256) {
            unsigned long long e[8];
            unsigned okm = 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int f = f0 + q * 32 + lane;
                const int l = f / cpl, k = f - l * cpl;
                const bool ok = f < space && k < sCnt[l];
                e[q] = ok ? a.cand[((int64_t)l * a.n_users_pad + g) * TC_C + k] : 0ULL;
                okm |= (ok ? 1u : 0u) << q;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int f = f0 + q * 32 + lane;
                const bool ok = (okm >> q) & 1u;
                const uint32_t o = ok ? ord_bits((uint32_t)(e[q] >> 32)) : 0u;   // 0 sorts below every score: holes never count
                if (f < space) sO[f] = o;
                if (ok) { lo_b = min(lo_b, o); hi_b = max(hi_b, o); ++sub; }
            }
        }
        lo_b = __reduce_min_sync(0xffffffffu, lo_b);
        hi_b = __reduce_max_sync(0xffffffffu, hi_b);
        sub = __reduce_add_sync(0xffffffffu, sub);
        __syncwarp();
        if (sub >= a.K) {
            uint32_t T_b = lo_b;   // count(o >= lo_b) = sub >= K
            if (sub > 2 * a.K) {
                uint32_t lo = lo_b, hi = hi_b;   // invariant: count(o >= lo) >= K
                for (int it = 0; it < 32 && lo < hi; ++it) {
                    const uint32_t mid = lo + ((hi - lo + 1) >> 1);
                    int c = 0;
                    for (int f = lane; f < space; f += 32) c += sO[f] >= mid;
                    c = __reduce_add_sync(0xffffffffu, c);
                    if (c >= a.K) { lo = mid; if (c <= 2 * a.K) break; } else hi = mid - 1;
                }
                T_b = lo;
            }
            cutoff = __uint_as_float((T_b & 0x80000000u) ? (T_b & 0x7fffffffu) : ~T_b) - 2.f * eps;
        }
        __syncwarp();
    }
    // survivors of ALL entries, compacted (the bisection array is dead: its words now hold the survivors' item ids)
    uint32_t* sC = sO;
    int live_n = 0;
    const int space2 = L * maxn;
    for (int f0 = 0; f0 < space2; f0 += 256) {
        unsigned long long e[8];
        unsigned okm = 0u;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int f = f0 + q * 32 + lane;
            const int l = f / maxn, k = f - l * maxn;
            const bool ok = f < space2 && k < sCnt[l];
            e[q] = ok ? a.cand[((int64_t)l * a.n_users_pad + g) * TC_C + k] : 0ULL;
            okm |= (ok ? 1u : 0u) << q;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const bool lv = ((okm >> q) & 1u) && !(__uint_as_float((uint32_t)(e[q] >> 32)) < cutoff);
            const unsigned m = __ballot_sync(0xffffffffu, lv);
            const int pos = live_n + __popc(m & ((1u << lane) - 1u));
            if (lv && pos < 32 * RS_LIVE) sC[pos] = (uint32_t)e[q];
            live_n += __popc(m);
        }
    }
    __syncwarp();
    if (live_n > 32 * RS_LIVE) return false;
    uint32_t item[RS_LIVE];
    unsigned long long key[RS_LIVE];
#pragma unroll
    for (int r = 0; r < RS_LIVE; ++r) {
        item[r] = r * 32 + lane < live_n ? sC[r * 32 + lane] : 0u;
        key[r] = 0ULL;
    }
    __syncwarp();   // the ids are in registers: the staging area is free for item rows
    const float* p = a.P + (int64_t)a.users[g] * a.dim;
#pragma unroll
    for (int r = 0; r < RS_LIVE; ++r) {
        if (r * 32 < live_n) {
            const bool live = r * 32 + lane < live_n;
            const unsigned live_mask = __ballot_sync(0xffffffffu, live);
            const float acc = rescore_group<KIND>(a, p, item[r], live, live_mask, lane, sQ, sP);
            if (live) key[r] = rank_key(acc, item[r], ASC);
        }
    }
    unsigned long long prev = ~0ULL, kth = 0ULL;
    for (int r = 0; r < a.K; ++r) {
        unsigned long long best = 0ULL;
#pragma unroll
        for (int q = 0; q < RS_LIVE; ++q)
            if (key[q] < prev && key[q] > best) best = key[q];
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) {
            a.out_items[g * a.K + r] = best ? (int32_t)key_index(best) : -1;
            if (a.out_scores) a.out_scores[g * a.K + r] = best ? key_score(best, ASC) : 0.f;
        }
        if (best) { prev = best; kth = best; } else { kth = 0ULL; prev = 0ULL; }
    }
    *theta_out = theta; *kth_out = kth; *total_out = total;
    return true;
}

template <int KIND>
__global__ void __launch_bounds__(256) rescore_kernel(RescoreArgs a) {
    constexpr int ASC = KIND == CRB_SCORE_SQDIST ? 1 : 0;
    extern __shared__ float rs_sm[];
    const int lane = threadIdx.x & 31;
    float* sQ = rs_sm + (threadIdx.x >> 5) * RS_WARP_WORDS;
    float* sP = sQ + RS_ORD;
    int* sCnt = reinterpret_cast<int*>(sP + RS_CH);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp; g < a.n_users; g += n_warps) {
        // error bound of one approximate score for this user (same expression as the certificate below)
        float eps;
        {
            const float qmax = __uint_as_float(a.maxbits[0]);
            if (KIND == CRB_SCORE_SQDIST) { const float s_ = a.pnorm[g] + qmax; eps = a.cbound * s_ * s_; }
            else { eps = a.cbound * a.pnorm[g] * qmax; if (KIND == CRB_SCORE_DOT_BIAS) eps += a.cbound * __uint_as_float(a.maxbits[1]); }
        }
        float theta;
        unsigned long long kth;
        int total;
        __syncwarp();   // the previous user's shared-memory words are dead
        if (a.lists_only || !rescore_user_fast<KIND>(a, g, lane, sQ, sP, sCnt, eps, &theta, &kth, &total))
            rescore_user_lists<KIND>(a, g, lane, sQ, sP, eps, &theta, &kth, &total);
        if (lane == 0) atomicMax(a.counters + 1, (unsigned int)total);
        // certificate
        bool ok;
        if (theta == -INFINITY) {
            ok = true;  // nothing was ever dropped: every unseen item is in the lists
        } else if (kth == 0ULL) {
            ok = false; // fewer than K listed although items were dropped
        } else {
            const float qmax = __uint_as_float(a.maxbits[0]);
            const float ek = key_score(kth, ASC);
            if (KIND == CRB_SCORE_SQDIST) {
                const float s = a.pnorm[g] + qmax;
                const float eps = a.cbound * s * s;
                ok = ek < (a.psq[g] - theta) - eps;      // unlisted: dist >= |p|^2 - theta - eps
            } else {
                float eps = a.cbound * a.pnorm[g] * qmax;
                if (KIND == CRB_SCORE_DOT_BIAS) eps += a.cbound * __uint_as_float(a.maxbits[1]);
                ok = ek > theta + eps;                   // unlisted: score <= theta + eps
            }
        }
        if (!ok && lane == 0) a.todo[atomicAdd(a.counters, 1u)] = (int32_t)g;
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap* map, void* base, int64_t rows, int d_pad) {
    static encode_fn_t encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CRB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { crb_set_error("cuTensorMapEncodeTiled not available"); return CRB_ERR_CUDA; }
        encode = (encode_fn_t)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)d_pad * 2};
    const cuuint32_t box[2] = {64, 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { crb_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return CRB_ERR_CUDA; }
    return CRB_OK;
}

static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

int crb_score_topk_tc(crb_handle* h, int32_t kind, const float* P, const float* Q, const float* hvec, int64_t n_items, int32_t dim,
                      const int32_t* users, const int32_t* hist_users, int64_t n_users, int32_t K, int32_t* topk_items,
                      float* topk_scores, cudaStream_t s) {
    CRB_CHECK_ARG(K >= 1, "K");
    const bool aug = kind == CRB_SCORE_SQDIST || kind == CRB_SCORE_DOT_BIAS;
    const int d_pad = (int)round_up(dim + (aug ? 2 : 0), 64);
    const int kb = d_pad / 64;
    if (kb > TC_MAX_KB || K > 32) {
        // operand tile would not fit in shared memory, or K is beyond the candidate lists' certificate (the reference's argsort takes
        // any K, e.g. topk=[10,20,50]): this call runs on the exact kernel, K <= 256 (documented in DESIGN.md)
        h->topk_stats[0] = 0; h->topk_stats[1] = n_users; h->topk_stats[2] = 0;
        return crb_launch_fullrank_exact(h, kind, P, Q, hvec, n_items, dim, users, hist_users, nullptr, n_users, K, topk_items, topk_scores, s);
    }
    CRB_CHECK_ARG(n_items < 0x7fffff00LL, "catalogue too large");
    const int64_t n_items_pad = round_up(n_items, TC_BN);
    const int n_tiles = (int)(n_items_pad / TC_BN);
    const int64_t pass_users = n_users < (1 << 18) ? n_users : (1 << 18);   // bounds the candidate-list workspace (1 GB)
    const int64_t pass_pad = round_up(pass_users, TC_BM);
    const int64_t m_tiles_max = pass_pad / TC_BM;
    // Item splits: a work item is (256-user tile, item range); every extra range costs a list start-up (until a list has seen ~5e4
    // items almost every chunk has a score above the threshold of one of the warp's 32 users and takes the divergent path) and more
    // candidates to re-score: measured ~0.9 ms per work item at 4096 users x 2M items, i.e. 800 tiles' worth.  Pick the split count
    // that minimises waves x (tiles per range + overhead) on this machine.
    int n_splits = 1;
    {
        double split_overhead = 800.0;
        if (const char* e = getenv("CRB_TC_SPLIT_OVERHEAD")) split_overhead = atof(e);
        double best = 0.0;
        for (int n = 1; n <= 64 && n <= n_tiles; ++n) {
            const int64_t waves = (m_tiles_max * n + h->sm_count - 1) / h->sm_count;
            const double cost = (double)waves * ((double)((n_tiles + n - 1) / n) + split_overhead);
            if (n == 1 || cost < best * 0.97) { if (n == 1 || cost < best) { best = cost; n_splits = n; } }
        }
    }
    if (const char* fs = getenv("CRB_TC_SPLITS")) {   // experiments: force the split count
        const int v = atoi(fs);
        if (v >= 1 && v <= 64 && v <= n_tiles) n_splits = v;
    }
    const int tiles_per_split = (n_tiles + n_splits - 1) / n_splits;
    n_splits = (n_tiles + tiles_per_split - 1) / tiles_per_split;
    // workspace.  The bf16 copy of the item table and its norms' maxima come first, at offsets that depend only on (n_items, d_pad):
    // they are kept across calls (h->evq_*) as long as no library call has written a table or reused the workspace since -- a
    // caller that ranks users in many small calls (the reference's test.batch_size loop) converts Q once, not once per call.
    const int64_t slots = (int64_t)n_splits * TC_CH * pass_pad;   // one candidate list per (split, column half, user)
    const int64_t n_passes = (n_users + pass_users - 1) / pass_users;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { int64_t o = off; off += (bytes + 1023) & ~(int64_t)1023; return o; };
    const int64_t o_qb = take(n_items_pad * d_pad * 2), o_misc = take(64);
    const int64_t o_pb = take(pass_pad * d_pad * 2), o_cand = take(slots * TC_C * 8);
    const int64_t o_cnt = take(slots * 4), o_thr = take(slots * 4), o_pn = take(pass_pad * 4), o_psq = take(pass_pad * 4);
    const int64_t o_todo = take(n_passes * pass_pad * 4), o_ctr = take(n_passes * 8);
    void* ws_before = h->eval_ws;
    int rc = crb_eval_ws_reserve(h, off + 1024);
    if (rc) return rc;
    if (h->eval_ws != ws_before) h->evq_valid = 0;
    char* ws = (char*)(((uintptr_t)h->eval_ws + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* qb = (__nv_bfloat16*)(ws + o_qb);
    __nv_bfloat16* pb = (__nv_bfloat16*)(ws + o_pb);
    unsigned int* misc = (unsigned int*)(ws + o_misc);  // [0..1] max bits of the item norms / biases (part of the cached state)
    unsigned int* ctrs = (unsigned int*)(ws + o_ctr);   // per pass: [0] uncertified users, [1] max candidates
    const int prep_grid = h->sm_count * 8;
    const bool cached = h->evq_valid && h->evq_q == Q && h->evq_hvec == hvec && h->evq_items == n_items && h->evq_dim == dim && h->evq_kind == kind;
    if (!cached) {
        CRB_CUDA(cudaMemsetAsync(misc, 0, 64, s));
        PrepArgs pq = {Q, hvec, nullptr, n_items, n_items_pad, dim, d_pad, kind, qb, nullptr, nullptr, misc};
        prep_kernel<false><<<prep_grid, 256, 0, s>>>(pq);
        h->launches++;
        CRB_CUDA(cudaGetLastError());
        h->tc_robust = 0;
        h->evq_valid = 1; h->evq_q = Q; h->evq_hvec = hvec; h->evq_items = n_items; h->evq_dim = dim; h->evq_kind = kind;
    }
    CRB_CUDA(cudaMemsetAsync(ctrs, 0, n_passes * 8, s));
    CUtensorMap map_b;
    if ((rc = make_map(&map_b, qb, n_items_pad, d_pad))) return rc;
    // shared memory plan
    const size_t a_bytes = (size_t)2 * kb * 16384, b_bytes = (size_t)kb * 16384;
    int stages = (int)((size_t)(220 * 1024 - a_bytes) / b_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) { crb_set_error("internal: shared memory plan"); return CRB_ERR_UNSUPPORTED; }
    const size_t smem_bytes = a_bytes + stages * b_bytes + 256 + 1024;
    CRB_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    const float cbound = 1.0f / 128.0f + (float)d_pad * (1.0f / 4194304.0f);
    int64_t certified = 0, rerun = 0, maxcand = 0;
    const char* lo_env = getenv("CRB_RESCORE_LISTS");
    const int lists_only = (lo_env && atoi(lo_env)) || n_splits * TC_CH > RS_MAXL;
    int64_t pass = 0;
    for (int64_t u0 = 0; u0 < n_users; u0 += pass_users, ++pass) {
        const int64_t nu = (n_users - u0) < pass_users ? (n_users - u0) : pass_users;
        const int64_t nu_pad = round_up(nu, TC_BM);
        PrepArgs pu = {P, hvec, users + u0, nu, nu_pad, dim, d_pad, kind, pb, (float*)(ws + o_pn), (float*)(ws + o_psq), nullptr};
        prep_kernel<true><<<prep_grid, 256, 0, s>>>(pu);
        CUtensorMap map_a;
        if ((rc = make_map(&map_a, pb, nu_pad, d_pad))) return rc;
        TcArgs ta;
        ta.n_users = nu; ta.n_users_pad = nu_pad; ta.n_items = n_items; ta.n_tiles = n_tiles; ta.n_splits = n_splits;
        ta.tiles_per_split = tiles_per_split; ta.kb = kb; ta.stages = stages; ta.users = users + u0;
        ta.hist_users = hist_users ? hist_users + u0 : nullptr; ta.seen_rowptr = h->seen_rowptr; ta.seen_cols = h->seen_cols;
#ifdef TC_DEBUG_SWITCHES
        ta.debug = getenv("CRB_TC_DEBUG") ? atoi(getenv("CRB_TC_DEBUG")) : 0;
        ta.dbg = nullptr;
        if (getenv("CRB_TC_DEBUG_CYCLES")) { cudaMalloc(&ta.dbg, 64); cudaMemsetAsync(ta.dbg, 0, 64, s); }
#else
        ta.debug = 0;
        ta.dbg = nullptr;
#endif
        // Compaction policy (see compact_list).  The certificate needs the K-th best canonical score of the merged lists to clear the
        // largest list threshold by the bf16 error bound, i.e. the thresholds to sit around rank 1.5 K of the user's scores or lower:
        // a list holds a 1 / n_lists share of those on average, so it keeps that share plus four standard deviations (never less than
        // 16) -- 31 entries for K = 20 with two lists, 16 from eight lists on -- and is compacted again half a range later.  A failed
        // certificate costs an exact re-run of the user (~0.25 ms), so the margin errs on the safe side.
        // The share assumes that a user's best items are spread over the item splits like the items themselves.  On a catalogue whose
        // ids are sorted by popularity they may all sit in one split: the certificates of many users then fail, results stay exact
        // (re-runs) but slow, and the handle switches to the margin of two lists -- the column halves interleave inside every tile and
        // cannot be skewed -- until the item table changes (h->tc_robust).
        {
            const double share = 1.5 * K / (double)((h->tc_robust ? 1 : n_splits) * TC_CH);
            int lo = (int)ceil(share + 4.0 * sqrt(share));
            ta.keep_lo = lo < 16 ? 16 : lo > 64 ? 64 : lo;
        }
        if (const char* e = getenv("CRB_TC_KEEP_LO")) { const int v = atoi(e); if (v >= 8 && v <= 64) ta.keep_lo = v; }
        ta.keep_hi = 2 * ta.keep_lo;
        ta.trig = ta.keep_hi + (ta.keep_lo / 2 > 16 ? ta.keep_lo / 2 : 16);
        if (const char* e = getenv("CRB_TC_TRIG")) { const int v = atoi(e); if (v > ta.keep_hi) ta.trig = v; }
        if (ta.trig > TC_C - 32 * (TC_BN / 32 / TC_CH)) ta.trig = TC_C - 32 * (TC_BN / 32 / TC_CH);   // the next tile must fit (32 appends per chunk at worst)
        ta.cand = (unsigned long long*)(ws + o_cand); ta.cand_cnt = (int32_t*)(ws + o_cnt); ta.cand_thr = (float*)(ws + o_thr);
        const int64_t n_work = (nu_pad / TC_BM) * n_splits;
        const int grid = (int)(n_work < h->sm_count ? n_work : h->sm_count);
        score_tc_kernel<<<grid, TC_THREADS, smem_bytes, s>>>(map_a, map_b, ta);
        CRB_CUDA(cudaGetLastError());
#ifdef TC_DEBUG_SWITCHES
        if (ta.dbg) {   // sums over all CTAs (MMA warp: [0] wait for B tiles, [1] wait for a drained accumulator, [5] total) and over all
                        // epilogue warps ([2] wait for an accumulator, [3] scan + appends, [4] compactions, [6] number of compactions)
            unsigned long long c[8];
            cudaStreamSynchronize(s);
            cudaMemcpy(c, ta.dbg, 64, cudaMemcpyDeviceToHost);
            cudaFree(ta.dbg);
            const double ctas = (double)grid, ew = ctas * 8 * TC_CH;
            fprintf(stderr, "tc cycles per CTA: mma total %.0f  wait B %.0f  wait acc %.0f | per epilogue warp: wait %.0f scan %.0f compact %.0f (n %.1f)\n",
                    c[5] / ctas, c[0] / ctas, c[1] / ctas, c[2] / ew, c[3] / ew, c[4] / ew, c[6] / ew);
        }
#endif
        RescoreArgs ra;
        ra.kind = kind; ra.dim = dim; ra.K = K; ra.n_splits = n_splits * TC_CH; ra.n_users = nu; ra.n_users_pad = nu_pad;
        ra.P = P; ra.Q = Q; ra.hvec = hvec; ra.users = users + u0; ra.cand = ta.cand; ra.cand_cnt = ta.cand_cnt; ra.cand_thr = ta.cand_thr;
        ra.pnorm = (const float*)(ws + o_pn); ra.psq = (const float*)(ws + o_psq); ra.maxbits = misc; ra.cbound = cbound;
        ra.out_items = topk_items + u0 * K; ra.out_scores = topk_scores ? topk_scores + u0 * K : nullptr;
        ra.todo = (int32_t*)(ws + o_todo) + pass * pass_pad; ra.counters = ctrs + 2 * pass;
        ra.lists_only = lists_only;
        const int rgrid = (int)((nu + 7) / 8 < (int64_t)h->sm_count * 8 ? (nu + 7) / 8 : (int64_t)h->sm_count * 8);
        const size_t rsm = sizeof(float) * 8 * RS_WARP_WORDS;
#define CRB_RESCORE(KK) CRB_CUDA(cudaFuncSetAttribute(rescore_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm)); \
                        rescore_kernel<KK><<<rgrid, 256, rsm, s>>>(ra);
        switch (kind) {
            case CRB_SCORE_DOT: { CRB_RESCORE(CRB_SCORE_DOT) } break;
            case CRB_SCORE_GMF: { CRB_RESCORE(CRB_SCORE_GMF) } break;
            case CRB_SCORE_SQDIST: { CRB_RESCORE(CRB_SCORE_SQDIST) } break;
            case CRB_SCORE_DOT_BIAS: { CRB_RESCORE(CRB_SCORE_DOT_BIAS) } break;
            default: crb_set_error("unknown score kind %d", kind); return CRB_ERR_ARG;
        }
        h->launches += 3;
        CRB_CUDA(cudaGetLastError());
    }
    // one host read for the whole call: users whose certificate failed (none on realistic tables) are re-run exactly, pass by pass
    unsigned int* cnts = (unsigned int*)malloc(sizeof(unsigned int) * 2 * n_passes);
    if (!cnts) { crb_set_error("out of host memory"); return CRB_ERR_ARG; }
    cudaError_t ce = cudaMemcpyAsync(cnts, ctrs, sizeof(unsigned int) * 2 * n_passes, cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess) { free(cnts); CRB_CUDA(ce); }
    pass = 0;
    for (int64_t u0 = 0; u0 < n_users; u0 += pass_users, ++pass) {
        const int64_t nu = (n_users - u0) < pass_users ? (n_users - u0) : pass_users;
        unsigned int bad = cnts[2 * pass];
#ifdef TC_DEBUG_SWITCHES
        if (getenv("CRB_TC_DEBUG") && atoi(getenv("CRB_TC_DEBUG"))) bad = 0;   // experiments: results are meaningless, do not re-run anyone
#endif
        if (bad) {
            // todo holds pass-local user slots: the exact kernel indexes users/outputs of this pass
            rc = crb_launch_fullrank_exact(h, kind, P, Q, hvec, n_items, dim, users + u0, hist_users ? hist_users + u0 : nullptr,
                                           (const int32_t*)(ws + o_todo) + pass * pass_pad, bad, K, topk_items + u0 * K,
                                           topk_scores ? topk_scores + u0 * K : nullptr, s);
            if (rc) { free(cnts); return rc; }
        }
        certified += nu - bad;
        rerun += bad;
        if ((int64_t)cnts[2 * pass + 1] > maxcand) maxcand = cnts[2 * pass + 1];
    }
    free(cnts);
    if (rerun * 64 > n_users) h->tc_robust = 1;
    h->topk_stats[0] = certified; h->topk_stats[1] = rerun; h->topk_stats[2] = maxcand; h->topk_stats[3] = n_splits;
    return CRB_OK;
}
