// Shared internals of libcleverrec_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cleverrec_b200.h"

#define CRB_SAMPLER_MAX_BLOCKS 4096u  // attempt guard (oracle/philox.py MAX_BLOCKS)
#define CRB_LRT_TABLE 65536           // Adam lr_t table length; beyond it lr_t == lr in fp32
#define CRB_DUP_CHUNK 256             // slots per work item of the duplicate-row reduction
#define CRB_PROF_CAP 1024
#define CRB_PROF_TAGS 4               // profiling hook: 0 fused step kernel, 1 staged fetch, 2 duplicate reduce / send, 3 owner's inbox pass

void crb_set_error(const char* fmt, ...);

#define CRB_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            crb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return CRB_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define CRB_CHECK_ARG(cond, msg)             \
    do {                                     \
        if (!(cond)) {                       \
            crb_set_error("bad argument: %s", msg); \
            return CRB_ERR_ARG;              \
        }                                    \
    } while (0)

// Device-side counters of one training step (zeroed by the step's first kernel's predecessor memset).
struct crb_step_ctr {
    unsigned int dup_slots;   // total gradient slots handed to duplicate rows
    unsigned int dup_rows;    // number of duplicate rows
    unsigned int work_items;  // chunk work items of the duplicate reduction
    unsigned int multi_rows;  // duplicate rows with more than one chunk
    unsigned int partial_slots; // partial-sum rows handed to multi-chunk duplicate rows
    unsigned int sampler_err; // a positive ran out of attempts (sticky: everything BEFORE this word is zeroed between steps)
    unsigned int tail_done;   // dup_tail_kernel: blocks that have finished their share (reset by the last one)
    unsigned int pad[1];
};

struct crb_dup_row {   // one row that occurs >= 2 times in the batch
    int32_t row;
    int32_t table;     // 0 = user-side table, 1 = item-side table
    uint32_t base;     // first gradient slot
    uint32_t cnt;      // occurrences
    uint32_t wbase;    // first work item
    uint32_t nchunk;
    uint32_t pbase;    // first partial-sum row (multi-chunk rows only)
    uint32_t pad;
};

struct crb_work {      // one chunk of one duplicate row
    uint32_t dup;      // index into dup rows
    uint32_t chunk;
};

struct crb_handle {
    int device;
    int sm_count;
    // history (borrowed device pointers)
    int64_t n_users, n_items, n_pos;
    const int32_t* pos_user;
    const int32_t* pos_item;
    const int64_t* seen_rowptr;
    const int32_t* seen_cols;
    // Per-user Bloom filter over the seen items (library-owned, built by crb_set_history): 2^bloom_shift words per user, one hash.
    // A clear bit proves "not in the history" with ONE 32-byte probe whose address needs only (u, item) -- the sampler's rejection
    // test (utils/sampler.py:58-59) falls back to the exact binary search over seen_cols only on a set bit (~9 % at 8 bits/entry).
    uint32_t* bloom;
    int bloom_shift;
    int64_t bloom_words;        // allocated words
    int64_t bloom_stride;       // words per user (1 << bloom_shift, or ceil(n_items / 32) in exact mode)
    int bloom_exact;            // 1 = small catalogue: the filter is the user's exact seen-item bitmap (bit = item id), no search needed
    const int64_t* list_start;  // per-user offset / length of the interaction list inside pos_item (FISM / NAIS)
    const int32_t* list_len;
    const int64_t* ilist_start; // item-side lists (TransCF's iu_sp_mat, utils/tools.py:100-113): users of each item inside ipos_user
    const int32_t* ilist_len;
    const int32_t* ipos_user;
    double* dense_loss;         // [4 * loss_blocks] per-block partials of the dense loss terms
    // social structures of SBPR (crb_set_social; borrowed device pointers): the positives of users that have a non-empty SPu in
    // the sampler's enumeration order, every user's SPu list with the social coefficient of each entry, and the sorted-unique
    // union of the user's own and social items (the rejection set of utils/sampler.py:118-120)
    int64_t sp_n_pos;
    const int32_t* sp_pos_user;
    const int32_t* sp_pos_item;
    const int64_t* spu_start;
    const int32_t* spu_items;
    const float* spu_suk;
    const int64_t* excl_rowptr;
    const int32_t* excl_cols;
    // per-row batch multiplicity words: low 32 bits = count, high 32 bits = slot base
    unsigned long long* meta[2];
    int64_t meta_rows[2];
    // per-step workspace (grown on demand)
    int64_t cap_batch;     // triplets
    int32_t cap_dim;
    int32_t* idx[4];       // sampled u, i, j, nbr (or staged host feeds)
    float* yv;             // sampled / staged labels
    uint32_t* rank[3];     // occurrence rank of u, i, j inside their row
    int32_t* sb[2];        // multi-GPU step: per triplet, first gradient slot of the i / j item row if it repeats in the batch, else -1
    float* stage;          // multi-GPU step: [stage_rows, stage_dim] local copies of repeated item rows (row = first gradient slot)
    int64_t stage_rows;
    int32_t stage_dim;
    float* dup_grad;       // [3*cap_batch, dim] gradient slots of duplicate occurrences
    uint32_t* dup_t;       // [3*cap_batch] triplet index of each slot (deterministic order)
    uint32_t* dup_src;     // [3*cap_batch] source row of each slot when gradients are summed in place (DupArgs::dup_src)
    crb_dup_row* dup_rows; // [3*cap_batch]
    crb_work* work;        // [3*cap_batch]
    unsigned int* multi;   // [3*cap_batch / CRB_DUP_CHUNK + 1] indices of multi-chunk duplicate rows
    int64_t cap_partial;   // rows of `partial`
    float* partial;        // [3*cap_batch / CRB_DUP_CHUNK + rows] * dim  chunk partial sums
    crb_step_ctr* ctr;     // device
    double* block_loss;    // [loss_blocks]
    double* loss_dev;      // [cap_steps] when the caller's loss buffer is on the host
    int64_t cap_steps;
    int loss_blocks;
    int step_grid;         // CTAs of the last fused step launch (entries of block_loss that are valid)
    float* lrt;            // Adam lr_t table (device), built for (lr, beta1, beta2)
    double lrt_lr, lrt_b1, lrt_b2;
    float* dense_grad;     // small dense-variable gradient accumulator (h_gmf, ...)
    int64_t cap_dense;
    // evaluation workspace
    void* eval_ws;
    int64_t eval_ws_bytes;
    int64_t topk_stats[4];
    // bf16 copy of the item table kept in the evaluation workspace between crb_score_topk calls (score_tc.cu): valid for exactly
    // these arguments until a library call writes a table (crb_opt_to_dev clears it) or reuses the workspace
    int evq_valid;
    int tc_robust;      // score_topk_tc: the last call on this item table re-ran > 1/64 of its users exactly -> widest compaction margin
    const void* evq_q;
    const void* evq_hvec;
    int64_t evq_items;
    int32_t evq_dim, evq_kind;
    int64_t launches;
    // numpy_stream sampler mode (sampler_np.cu)
    void* np_state;        // device RandomState (624 words + pos)
    int np_seeded;
    void* np_raw;          // raw 32-bit outputs scratch
    int64_t np_raw_cap;
    void* np_scratch;
    int64_t np_scratch_cap;
    void* np_sort_tmp;
    int64_t np_sort_cap;
    int64_t* np_result;
    // Second copy of the sampling / assignment state of a step.  crb_train_epoch_bpr prepares step k+1 (K1 sampler+count, K2
    // assign) on `aux_stream` into one copy while step k's K3/K4 consume the other; crb_alt_swap exchanges the pointers below
    // with the handle's own (kernels receive the pointers by value at launch, so swapping between launches is safe).
    struct {
        int32_t* idx[3];
        uint32_t* rank[3];
        int32_t* sb[2];
        unsigned long long* meta[2];
        int64_t meta_rows[2];
        crb_step_ctr* ctr;
        crb_dup_row* dup_rows;
        crb_work* work;
        unsigned int* multi;
        int64_t cap_batch;
        int ctr_zeroed;
    } alt;
    int ctr_zeroed;          // 1 = the step counters of the current copy are known to be zero (the last step on it ended in dup_tail_kernel)
    int alt_active;          // 1 while the handle's fields hold the alternate copy
    // Whole-epoch CUDA graph of crb_train_epoch_bpr for launch-bound shapes (train.cu): the epoch's kernels are captured once with
    // everything that changes from epoch to epoch (the sampler's permutation keys and epoch word, the optimizer's step base) read
    // from `dyn_dev` on the device, and replayed with one launch per epoch.
    uint32_t* dyn_dev;       // [8]: keys[6], epoch, step_base
    int dyn_mode;            // 1 while capturing: kernels get the dyn pointers
    void* epoch_graph;       // cudaGraphExec_t
    unsigned char epoch_graph_key[256];
    int epoch_graph_key_len;
    int64_t ws_generation;   // bumped whenever a workspace pointer baked into the graph may have changed
    // a sharded step prepared ahead of time in the alternate copy (crb_shard_step_prepare)
    int prep_valid;
    uint64_t prep_seed;
    uint32_t prep_epoch;
    int64_t prep_first, prep_batch;
    int32_t prep_neg_ratio;
    cudaStream_t aux_stream;
    cudaEvent_t ev_entry, ev_prep[2], ev_done[2];
    // profiling hook
    int prof_on;
    int prof_n[CRB_PROF_TAGS];          // event pairs recorded since last read, per tag (0 = the fused step kernel)
    cudaEvent_t* prof_ev[CRB_PROF_TAGS]; // 2 * CRB_PROF_CAP events each
    double prof_ms[CRB_PROF_TAGS];
    int64_t prof_launches[CRB_PROF_TAGS];
};
int crb_prof_begin(crb_handle* h, cudaStream_t s, int tag = 0);
int crb_prof_end(crb_handle* h, cudaStream_t s, int tag = 0);

static inline bool crb_is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float dot4(float4 a, float4 b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

// numerically stable softplus(-x) = -log(sigmoid(x)) and sigmoid(x) - 1 = -sigmoid(-x)
// (SFU exp/log/rcp: absolute error ~1e-7 on a per-sample loss of order 1, relative ~1e-7 on the gradient scale)
__device__ __forceinline__ float softplus_neg(float x) { return fmaxf(-x, 0.f) + __logf(1.f + __expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) {
    const float e = __expf(-fabsf(x));
    const float s = __fdividef(1.f, 1.f + e);
    return x >= 0.f ? s : e * s;
}

// Philox4x32-10 (same constants as oracle/philox.py)
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// bit of `item` inside a user's Bloom filter of 32 << shift bits (one multiplicative hash, top bits)
__host__ __device__ __forceinline__ uint32_t crb_bloom_bit(uint32_t item, int shift) { return (item * 0x9E3779B1u) >> (27 - shift); }

// internal entry points shared between translation units
int crb_bloom_build(crb_handle* h, cudaStream_t s);
int crb_ws_reserve(crb_handle* h, int64_t batch, int32_t dim, int64_t steps, cudaStream_t s);
int crb_meta_reserve(crb_handle* h, int which, int64_t rows, cudaStream_t s);
int crb_eval_ws_reserve(crb_handle* h, int64_t bytes);
int crb_alt_reserve(crb_handle* h, cudaStream_t s);   // allocate / resize the alternate step state to match the primary
void crb_alt_swap(crb_handle* h);
int crb_lrt_prepare(crb_handle* h, const crb_opt* opt, cudaStream_t s);
int crb_launch_sample_pointwise(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count, int32_t neg_ratio, int32_t* u,
                                int32_t* i, float* y, bool count_rows, cudaStream_t s);
int crb_launch_sample_pairwise(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count, int32_t neg_ratio,
                               int32_t* u, int32_t* i, int32_t* j, int32_t* nbr, bool count_rows, cudaStream_t s);
