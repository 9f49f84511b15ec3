/*
 * cleverrec_b200 -- C ABI of the B200-native hot path (libcleverrec_b200.so).
 *
 * This is the drop-in boundary.  In the reference every numeric operation of the hot path is
 * one `self.sess.run(fetches, feed_dict)` into TensorFlow-1 (model/RankingRecommender.py:46,59,
 * 85,98,221,278) preceded by a Python sampler call (utils/sampler.py:10-99).  Each entry point
 * below replaces one of those calls; the comment on it cites the reference interface it stands
 * for.  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success or a negative crb_status and
 *    leaves a thread-local message readable through crb_last_error().
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All work is
 *    enqueued asynchronously on it unless a HOST output pointer forces a synchronisation.
 *  - table pointers (crb_table) are DEVICE pointers owned by the caller (torch tensors).
 *  - index / label / output buffers may be DEVICE or HOST pointers: the library inspects them
 *    with cudaPointerGetAttributes and stages host buffers itself (that is the path `e2e` in
 *    bench.py times: host feed in, host loss/ranks out, exactly like a TF feed_dict).
 *  - there is NO CPU implementation behind any entry point: without a CUDA device every call
 *    fails with CRB_ERR_CUDA.
 */
#ifndef CLEVERREC_B200_H
#define CLEVERREC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRB_ABI_VERSION 2

typedef enum {
    CRB_OK = 0,
    CRB_ERR_ARG = -1,      /* bad argument (null pointer, dim not multiple of 4, ...) */
    CRB_ERR_CUDA = -2,     /* a CUDA runtime call failed; message holds cudaGetErrorString */
    CRB_ERR_STATE = -3,    /* call order (e.g. sampling before crb_set_history) */
    CRB_ERR_SAMPLER = -4,  /* a user has fewer than neg_ratio unseen items (the reference loops forever) */
    CRB_ERR_UNSUPPORTED = -5
} crb_status;

typedef struct crb_handle crb_handle;

/* One embedding table with its optimizer slots (TF variables + slot variables).
 *  w    [rows, dim] fp32 row-major, 16-byte aligned, dim % 4 == 0
 *  s1   Adagrad accumulator (init 0.1) | Adam m   (NULL for SGD)
 *  s2   Adam v                                   (NULL unless Adam)
 *  last int32 [rows]: step of the last Adam update of the row (NULL unless adam_mode == CRB_ADAM_TF1)  */
typedef struct {
    float* w;
    float* s1;
    float* s2;
    int32_t* last;
    int64_t rows;
    int32_t dim;
    int32_t _pad;
} crb_table;

enum { CRB_OPT_SGD = 0, CRB_OPT_ADAGRAD = 1, CRB_OPT_ADAM = 2 };
/* CRB_ADAM_TF1 : tf.train.AdamOptimizer sparse-apply semantics (moments of EVERY row decay every step and
 *                every row moves) realised row-sparsely by replaying a row's missed steps
 *                when it is next touched / at crb_adam_flush -- identical result, sparse traffic.
 * CRB_ADAM_LAZY: LazyAdam (touched rows only).  A documented deviation; throughput mode.            */
enum { CRB_ADAM_TF1 = 0, CRB_ADAM_LAZY = 1 };

/* utils/tools.py:79-87 (get_optimizer) with TF-1 defaults. `step` is the 1-based index of THIS step. */
typedef struct {
    int32_t kind;
    int32_t adam_mode;
    double lr;     /* doubles: Adam's lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t) is evaluated in double (as TF/Python do) */
    double beta1;  /* then rounded to fp32; the kernels use the fp32 roundings of lr, beta1, beta2, eps         */
    double beta2;
    double eps;
    int64_t step;
} crb_opt;

/* Multi-GPU (one process per GPU, one box): the item table is row-sharded, owner = item % n_ranks, local row = item / n_ranks.
 * Every rank maps every shard, every inbox and every flag array into its own address space (crb_malloc + crb_ipc_export /
 * crb_ipc_open), so q[r] / inbox_*[r] / flags[r] are pointers valid ON THIS DEVICE to rank r's memory (NVLink peer access);
 * index `rank` is local memory.  An inbox is DIRECT-MAPPED: one gradient slot per (source rank, local row) --
 * inbox_grad[r] is float [n_ranks][rows_cap][dim], inbox_stamp[r] is uint32 [n_ranks][rows_cap] (the step whose gradient the slot
 * holds; created zeroed).  A source sums its own duplicates before sending, so a slot is written at most once per step: the inbox
 * cannot overflow and is never cleared.  rows_cap >= q[r].rows for every r.  flags[r] is uint32 [CRB_SHARD_FLAGS]: words
 * [0, n_ranks) are the arrival flags of crb_shard_barrier (word s = last barrier ticket rank s announced to rank r), word
 * CRB_SHARD_ERR is a sticky error word (a barrier that timed out). */
#define CRB_MAX_RANKS 8
#define CRB_SHARD_FLAGS 16
#define CRB_SHARD_ERR 8
typedef struct {
    int32_t n_ranks;
    int32_t rank;
    int64_t rows_cap;
    crb_table q[CRB_MAX_RANKS];
    float* inbox_grad[CRB_MAX_RANKS];
    uint32_t* inbox_stamp[CRB_MAX_RANKS];
    uint32_t* flags[CRB_MAX_RANKS];
    float* dense_inbox[CRB_MAX_RANKS];   /* float [n_ranks][CRB_SHARD_DENSE]: per-source-rank gradient of a replicated dense variable (GMF's h) */
} crb_shard;
#define CRB_SHARD_DENSE 512

/* utils/tools.py:66-76 (get_loss) */
enum { CRB_LOSS_BPR = 0, CRB_LOSS_CROSS_ENTROPY = 1, CRB_LOSS_SQUARE = 2, CRB_LOSS_HINGE = 3 };

/* score kinds for the evaluation entry points (each model's `_predict`) */
enum {
    CRB_SCORE_DOT = 0,      /* BPR.py:49,51  MF               s = p_u . q_i                        */
    CRB_SCORE_GMF = 1,      /* GMF.py:40,43                   s = sum_k (p_uk*q_ik)*h_k  (the logit; sigmoid is monotone) */
    CRB_SCORE_SQDIST = 2,   /* CML.py:82,84                   s = sum_k (p_uk-q_ik)^2   (ascending is better)           */
    CRB_SCORE_DOT_BIAS = 3  /* FISM.py:53,70                  s = p_u . q_i + b_i  (p_u = precomputed user vector)      */
};

int crb_abi_version(void);
const char* crb_last_error(void);

/* main.py:39-45 (tf.Session creation) -> one handle per device. */
int crb_create(int device, crb_handle** out);
int crb_destroy(crb_handle* h);

/* data.ui_train (model/RankingPreprocess.py:117) in device form.
 *  pos_user/pos_item [n_pos]: the positives in the order utils/sampler.py:50-52 enumerates them
 *  seen_rowptr [n_users+1], seen_cols: per-user sorted unique history (the `seen_items` sets)
 * DEVICE pointers, borrowed until the next call / crb_destroy. */
int crb_set_history(crb_handle* h, int64_t n_users, int64_t n_items, int64_t n_pos,
                    const int32_t* pos_user, const int32_t* pos_item,
                    const int64_t* seen_rowptr, const int32_t* seen_cols, void* stream);

/* Device form of model/RankingPreprocess.py -- the producer of every input of the hot path.  All columns are
 * DEVICE arrays; every call synchronises the stream (the counts are host values).
 *
 * crb_prep_filter_reindex: _filter_users / _filter_items (RankingPreprocess.py:70-90; user_min / item_min <= 0 skip a filter) and
 *   re_index (:41-47).  raw_u / raw_i [n]: the log's id columns in file order.  Out: out_u / out_i [n_kept] new ids of the kept
 *   rows (file order), out_row [n_kept] their original row numbers (to gather `time` etc.), user_ids / item_ids [n] (first n_users /
 *   n_items valid): the raw id behind each new id.  New ids are the ranks of the surviving raw ids in ASCENDING order -- the
 *   reference's order (iteration of a Python set of the ids) for dense non-negative id spaces. */
int crb_prep_filter_reindex(crb_handle* h, const int64_t* raw_u, const int64_t* raw_i, int64_t n, int32_t user_min, int32_t item_min,
                            int32_t* out_u, int32_t* out_i, int64_t* out_row, int64_t* n_kept, int64_t* n_users, int64_t* n_items,
                            int64_t* user_ids, int64_t* item_ids, void* stream);
/* The leave-one-out split (RankingPreprocess.py:96-109).  u [n]: user of each row; time [n] or NULL (data.split_by_time).
 * perm [n]: the rows in the order the reference's frame enumerates them (ascending user, then time, ties and the no-time case in
 * file order); is_test [n], aligned with perm: 1 for a user's last row when the user has more than 3 rows. */
int crb_prep_split_loo(crb_handle* h, const int32_t* u, const int64_t* time, int64_t n, int64_t n_users, int64_t* perm,
                       unsigned char* is_test, void* stream);
/* Sampled evaluation negatives (RankingPreprocess.py:120-129): for each test user neg_samples (<= 1024) DISTINCT items outside
 * the user's training items (the installed history), out [n_test, neg_samples].  Same law as np.random.choice(..., replace=False)
 * over the unseen items; a pure function of (seed, user, history) with an integer-exact CPU twin (oracle/philox.py), not NumPy's
 * stream (use the packaged host preprocessing for that). */
int crb_prep_eval_negatives(crb_handle* h, uint64_t seed, const int32_t* test_users, int64_t n_test, int32_t neg_samples,
                            int32_t* out, void* stream);

/* Native builder of the arrays above from the training split's (user, item) rows -- replaces the dict / list / set forms of
 * model/RankingPreprocess.py:117 (`groupby('u_id').i_id.apply(list).to_dict()`), utils/sampler.py:53 and
 * model/RankingRecommender.py:222-240 (`set(ui_train[u])`), which cannot be materialised at 1e9 interactions.
 *  users/items [n]: int32 rows of the training split in file order, ids already re-indexed (HOST or DEVICE)
 *  pos_user/pos_item [n]: grouped by user (ascending id = the groupby order), row order kept inside a user      DEVICE out
 *  seen_rowptr [n_users+1], seen_cols [n] (first *n_seen entries valid): sorted-unique per user                  DEVICE out
 *  list_start [n_users+1], list_len [n_users]: each user's list inside pos_item (crb_set_history_lists), or NULL  DEVICE out
 * Synchronises the stream (n_seen is a host value).  Ids outside their range are an error (CRB_ERR_ARG). */
int crb_build_history(crb_handle* h, const int32_t* users, const int32_t* items, int64_t n, int64_t n_users, int64_t n_items,
                      int32_t* pos_user, int32_t* pos_item, int64_t* seen_rowptr, int32_t* seen_cols, int64_t* n_seen,
                      int64_t* list_start, int32_t* list_len, void* stream);

/* utils/sampler.py:46-74 pairwise_ranking_sampler: rows [first, first+count) of the epoch's shuffled
 * triplet list.  nbr (fism_like's u_neighbors_num) may be NULL.  Outputs: DEVICE int32. */
int crb_sample_pairwise(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count,
                        int32_t neg_ratio, int32_t* u, int32_t* i, int32_t* j, int32_t* nbr, void* stream);
/* utils/sampler.py:10-43 pointwise_ranking_sampler (1 positive + neg_ratio negatives per positive). */
int crb_sample_pointwise(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count,
                         int32_t neg_ratio, int32_t* u, int32_t* i, float* y, int32_t* nbr, void* stream);
/* utils/sampler.py:77-99 ranking_sampler_cml: neg is [count, neg_ratio] row-major. */
int crb_sample_cml(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count,
                   int32_t neg_ratio, int32_t* u, int32_t* i, int32_t* neg, void* stream);
/* ---- `numpy_stream` sampler mode: the reference's samplers bit for bit -------------------------------------------------
 * The device keeps a copy of NumPy's legacy global RandomState (MT19937 key[624] + pos).  crb_np_seed == np.random.seed(int);
 * crb_np_set_state / crb_np_get_state exchange the state with np.random.get_state() / set_state() (key and pos fields), so a
 * run can interleave reference code (e.g. RankingPreprocess) and device sampling on ONE stream exactly like the reference. */
int crb_np_seed(crb_handle* h, uint32_t seed);
int crb_np_set_state(crb_handle* h, const uint32_t* key624, int32_t pos);
int crb_np_get_state(crb_handle* h, uint32_t* key624, int32_t* pos);
/* One whole epoch exactly as the reference sampler returns it (utils/sampler.py), advancing the stream like the reference:
 *  kind 0 pairwise_ranking_sampler   u,i [N], third = j int32 [N], nbr int32 [N] or NULL            N = n_pos * neg_ratio
 *  kind 1 pointwise_ranking_sampler  u,i [N], third = y float [N]                                   N = n_pos * (neg_ratio + 1)
 *  kind 2 ranking_sampler_cml        u,i [N], third = neg int32 [N, neg_ratio]                      N = n_pos
 *  kind 3 negatives only, in draw order (train_model_nais, RankingRecommender.py:64-80): third = int32 [n_pos, neg_ratio]
 * DEVICE outputs.  The call synchronises (the number of random values consumed is data dependent). */
int crb_sample_epoch_numpy(crb_handle* h, int32_t kind, int32_t neg_ratio, int32_t* u, int32_t* i, void* third, int32_t* nbr,
                           void* stream);
/* The same for ranking_sampler_sbpr (utils/sampler.py:102-141; structures of crb_set_social): the social-item draw
 * np.random.randint(len(SPu[u])) -- which consumes NO value when the list has one item -- the rejection-sampled negative and the final
 * permutation, bit for bit.  u, i, i_s, i_neg int32 [sp_n_pos * neg_ratio], suk float or NULL; DEVICE outputs. */
int crb_sample_epoch_numpy_sbpr(crb_handle* h, int32_t neg_ratio, int32_t* u, int32_t* i, int32_t* i_s, int32_t* i_neg, float* suk,
                                void* stream);

/* number of rows in one epoch of each sampler (utils/sampler.py:65 `train_nums`) */
int64_t crb_epoch_rows(crb_handle* h, int32_t neg_ratio, int32_t sampler_kind /*0 pairwise,1 pointwise,2 cml,3 sbpr*/);

/* sess.run([self.train, self.loss], {u_idx, i_idx, j_idx})  for model/ranking/BPR.py:31-44.
 * loss_out (double*, DEVICE or HOST) receives this step's summed loss (`loss_val`). */
int crb_train_step_bpr(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_opt* opt,
                       const int32_t* u, const int32_t* i, const int32_t* j, int64_t batch,
                       float reg, double* loss_out, void* stream);

/* One call = `n_steps` iterations of RankingRecommender.train_model's loop body (:39-46) with the
 * sampler fused in: step k trains on epoch rows [first + k*batch, min(first+(k+1)*batch, epoch_rows)).
 * loss_out[k] (DEVICE or HOST double[n_steps]).  opt->step is the index of the first step. */
int crb_train_epoch_bpr(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_opt* opt,
                        uint64_t seed, uint32_t epoch, int64_t first, int64_t batch, int64_t n_steps,
                        int32_t neg_ratio, float reg, double* loss_out, void* stream);

/* The same loop over an epoch the CALLER sampled -- the reference's own `train_ = pairwise_ranking_sampler(...)` arrays, HOST (or
 * DEVICE) int32 [n_rows]: step k trains on rows [k*batch, min((k+1)*batch, n_rows)), exactly the slices of RankingRecommender.py:40-42.
 * Feeds are staged one step ahead on the library's copy stream, so the host -> device copies overlap the previous step's kernels.
 * loss_out: DEVICE or HOST double [ceil(n_rows / batch)]; page-locked host memory receives each step's loss with its own asynchronous
 * copy.  The call returns after the last loss has landed when loss_out is on the host.  opt->step is the index of the first step. */
int crb_train_epoch_bpr_feeds(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_opt* opt,
                              const int32_t* u, const int32_t* i, const int32_t* j, int64_t n_rows, int64_t batch,
                              float reg, double* loss_out, void* stream);

/* sess.run([train, loss], {u_idx, i_idx, y}) for the pointwise dot-product family:
 *   kind CRB_SCORE_DOT : MF   (BPR.py:39 dot + get_loss('square'|'cross_entropy'))
 *   kind CRB_SCORE_GMF : GMF  (GMF.py:37-49), h/h_s1/h_s2 = the dense variable h_gmf and its slots */
int crb_train_step_pointwise(crb_handle* h, int32_t kind, const crb_table* P, const crb_table* Q,
                             float* hvec, float* h_s1, float* h_s2, const crb_opt* opt, int32_t loss_kind,
                             const int32_t* u, const int32_t* i, const float* y, int64_t batch,
                             float reg, double* loss_out, void* stream);

/* RankingRecommender.train_model's pointwise loop (:48-60) with the sampler fused in, `n_steps` iterations per call: step k trains on
 * rows [first + k*batch, min(first + (k+1)*batch, epoch_rows)) of pointwise_ranking_sampler's epoch (utils/sampler.py:10-43).
 * loss_out[k] (DEVICE or HOST double [n_steps]); opt->step is the index of the first step. */
int crb_train_epoch_pointwise(crb_handle* h, int32_t kind, const crb_table* P, const crb_table* Q, float* hvec, float* h_s1,
                              float* h_s2, const crb_opt* opt, int32_t loss_kind, uint64_t seed, uint32_t epoch, int64_t first,
                              int64_t batch, int64_t n_steps, int32_t neg_ratio, float reg, double* loss_out, void* stream);

/* sess.run([train, loss], {u_idx, i_idx, neg_items}) for model/ranking/CML.py:39-70 (train_model_cml,
 * RankingRecommender.py:90-100).  The covariance regulariser makes both table gradients dense in the reference, so TF
 * applies the DENSE optimizer form to every row: gradP/gradQ are caller-owned zeroed [rows, dim] device buffers (left zeroed).
 * neg: int32 [batch, neg_ratio].  item_nums enters the WARP rank weight (CML.py:52). */
int crb_train_step_cml(crb_handle* h, const crb_table* P, const crb_table* Q, float* gradP, float* gradQ, const crb_opt* opt,
                       const int32_t* u, const int32_t* i, const int32_t* neg, int64_t batch, int32_t neg_ratio,
                       float margin, float reg, int64_t item_nums, double* loss_out, void* stream);

/* TransCF (model/ranking/TransCF.py:38-85).  crb_set_item_lists: the item-side lists of iu_sp_mat (utils/tools.py:100-113): for
 * every item the users that interacted with it (duplicates kept), item_users grouped by item, item_start [I+1] / item_len [I] --
 * crb_build_history on the swapped columns produces exactly these.  DEVICE arrays, borrowed.
 * crb_train_step_transcf: sess.run([train, loss], {u_idx, i_idx, j_idx}) (TransCF.py:38-62 + :65-71); gradP / gradQ zeroed
 * device buffers of the tables' shape (the gradients are dense in the reference: they flow through the two SpMMs).
 * crb_transcf_neighbourhood: which = 0 -> alpha rows (mean of `table` = Q over each user's items) for `rows` users (NULL = 0..n-1),
 * which = 1 -> beta rows (mean of `table` = P over each item's users).  crb_score_pairs_transcf: TransCF._predict's ui_dist for
 * flattened (u, i) pairs given A = alpha of all users [U, dim] and B = beta of all items [I, dim]. */
int crb_set_item_lists(crb_handle* h, const int64_t* item_start, const int32_t* item_len, const int32_t* item_users);
int crb_train_step_transcf(crb_handle* h, const crb_table* P, const crb_table* Q, float* gradP, float* gradQ, const crb_opt* opt,
                           const int32_t* u, const int32_t* i, const int32_t* j, int64_t batch, float margin, float reg1, float reg2,
                           double* loss_out, void* stream);
int crb_transcf_neighbourhood(crb_handle* h, int32_t which, const float* table, int32_t dim, const int32_t* rows, int64_t n, float* out,
                              void* stream);
int crb_score_pairs_transcf(crb_handle* h, const float* P, const float* Q, const float* A, const float* B, int32_t dim, const int32_t* u,
                            const int32_t* i, int64_t n, float* scores, void* stream);

/* Per-user interaction lists inside pos_item (the `items` lists of data.ui_train, order and duplicates kept): what
 * get_ui_sp_mat (utils/tools.py:90-97) encodes.  DEVICE arrays [n_users], borrowed.  Call after crb_set_history. */
int crb_set_history_lists(crb_handle* h, const int64_t* list_start, const int32_t* list_len);

/* sess.run([train, loss], {u_idx, i_idx, j_idx, u_neighbors_num}) for model/ranking/FISM.py:40-63 (pairwise).  P, Q are
 * [(I+1), dim]; B is the bias vector as a crb_table with dim = 1 and rows padded to a multiple of 4.  All three gradients
 * are dense in the reference (L2 over whole tables, FISM.py:57): gradP/gradQ/gradB as for CML.  conf_batch_size is the
 * configured batch_size that divides the L2 term (not the fed batch length). */
int crb_train_step_fism(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_table* B, float* gradP, float* gradQ,
                        float* gradB, const crb_opt* opt, const int32_t* u, const int32_t* i, const int32_t* j,
                        const int32_t* nbr, int64_t batch, float alpha, float reg, float reg_bias, int64_t conf_batch_size,
                        double* loss_out, void* stream);

/* FISM.py:70 `coeff * u_neighbors_embed` for evaluation: out[k] = nbr[k]^-alpha * mean(P[list(users[k])]).  DEVICE buffers. */
int crb_fism_user_vectors(crb_handle* h, const float* P, int32_t dim, const int32_t* users, const int32_t* nbr, int64_t n,
                          float alpha, float* out, void* stream);

/* tf.clip_by_norm(rows, max_norm, axes=[1]) (CML.py:72-78): dst may alias src. */
int crb_clip_rows(crb_handle* h, const float* src, float* dst, int64_t rows, int32_t dim, float max_norm, void* stream);

/* sess.run([train, loss], {u_idx, i_idx, y}) for model/ranking/NeuMF.py:58-95.  Tables: GMF pair [.,E], MLP pair [.,L0/2].
 * `dense` packs the dense variables in this order: for k in layers: W_k [layers[k], layers[k]/2] row-major, b_k; then
 * h_neumf [E + layers[-1]/2]  (layers[k+1] == layers[k]/2, n_layers <= 4, L0 <= 512).  g* are zeroed dense gradient
 * buffers of the tables (left zeroed); dense_s1/s2 the optimizer slots of `dense`.  * Pg == Qg == gPg == gQg == NULL selects model/ranking/MLP.py:44-70: the tower alone (logit = h_mlp . tower, h_neumf then holds
 * h_mlp only; reg1 unused, reg2 = MLP.py's reg). */
int crb_train_step_neumf(crb_handle* h, const crb_table* Pg, const crb_table* Qg, const crb_table* Pm, const crb_table* Qm,
                         float* gPg, float* gQg, float* gPm, float* gQm, float* dense, float* dense_s1, float* dense_s2,
                         int32_t n_layers, const crb_opt* opt, int32_t loss_kind, const int32_t* u, const int32_t* i,
                         const float* y, int64_t batch, float reg1, float reg2, double* loss_out, void* stream);

/* NeuMF._predict (NeuMF.py:97-105) on flattened pairs: the logit (sigmoid is monotone) in canonical order.  DEVICE buffers. */
int crb_score_pairs_neumf(crb_handle* h, const float* Pg, const float* Qg, const float* Pm, const float* Qm, const float* dense,
                          int32_t E, int32_t Em, int32_t n_layers, const int32_t* u, const int32_t* i, int64_t n,
                          float* scores, void* stream);

/* SBPR (model/ranking/SBPR.py:38-66; sampler utils/sampler.py:102-141; SPu utils/tools.py:115-127).
 * crb_set_social installs what get_SPu and the sampler's inner loops derive from data.user_friends, as DEVICE arrays (borrowed):
 *   sp_pos_user / sp_pos_item [n_sp_pos]  the positives of the users that have a non-empty SPu, in the sampler's enumeration order
 *                                         (utils/sampler.py:105-110: dict order, users without SPu skipped)
 *   spu_start [U+1], spu_items [n_spu]    SPu[u] as a list (the order `s = randint(len(spu)); SPu[u][s]` indexes, :114-115)
 *   spu_suk [n_spu]                       the social coefficient of each entry: how many of u's friends consumed it (:124-131)
 *   excl_rowptr [U+1], excl_cols          sorted unique union of u's own and social items: the rejection set of :118-120
 * crb_sample_sbpr: rows [first, first+count) of the shuffled epoch -> (u, i, i_s, i_neg, suk) DEVICE outputs (suk may be NULL:
 * is_suk=False).  crb_train_step_sbpr: sess.run([train, loss], {u_idx, i_idx, i_s_idx, i_neg_idx, suk}) (RankingRecommender.py:103-117);
 * B is the bias vector [item_nums + 1] as a crb_table with dim = 1, rows padded to a multiple of 4; gradP / gradQ / gradB zeroed
 * dense gradient buffers (left zeroed); feeds HOST or DEVICE. */
int crb_set_social(crb_handle* h, int64_t n_sp_pos, const int32_t* sp_pos_user, const int32_t* sp_pos_item, const int64_t* spu_start,
                   const int32_t* spu_items, const float* spu_suk, const int64_t* excl_rowptr, const int32_t* excl_cols);
int crb_sample_sbpr(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t first, int64_t count, int32_t neg_ratio, int32_t* u,
                    int32_t* i, int32_t* k, int32_t* j, float* suk, void* stream);
int crb_train_step_sbpr(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_table* B, float* gradP, float* gradQ, float* gradB,
                        const crb_opt* opt, const int32_t* u, const int32_t* i, const int32_t* k, const int32_t* j, const float* suk,
                        int64_t batch, float reg, double* loss_out, void* stream);

/* LRML (model/ranking/LRML.py:42-78).  crb_train_step_lrml: sess.run([train, loss], {u_idx, i_idx, j_idx}) -- the LRAM memory
 * module (x = p*q, softmax(x K) M, :42-51), translated distances (:59-60), hinge + reg * l2 of the three gathered rows (:62-63).
 * `dense` packs K [embed_size, mem_size] row-major then M [mem_size, embed_size] row-major; dense_s1 / dense_s2 its optimizer
 * slots; gradP / gradQ zeroed dense gradient buffers of the tables (left zeroed).  embed_size % 4 == 0, <= 256; mem_size <= 128.
 * crb_score_pairs_lrml: LRML._predict's ui_dist (:73 for fed pairs; :75-76 is the same distance for every (user, item) pair) on
 * flattened pairs, same forward arithmetic as the training step.  DEVICE buffers. */
int crb_train_step_lrml(crb_handle* h, const crb_table* P, const crb_table* Q, float* gradP, float* gradQ, float* dense,
                        float* dense_s1, float* dense_s2, int32_t mem_size, const crb_opt* opt, const int32_t* u, const int32_t* i,
                        const int32_t* j, int64_t batch, float margin, float reg, double* loss_out, void* stream);
int crb_score_pairs_lrml(crb_handle* h, const float* P, const float* Q, const float* dense, int32_t dim, int32_t mem_size,
                         const int32_t* u, const int32_t* i, int64_t n, float* scores, void* stream);

/* RankingRecommender.py:235-240 as a mask: scores[k, item] = value for every item users[k] has seen.  DEVICE buffers. */
int crb_mask_seen(crb_handle* h, float* scores, const int32_t* users, int64_t n_users, int64_t n_items, float value, void* stream);

/* The per-user batch of train_model_nais (RankingRecommender.py:64-80): targets [(1+neg_ratio)*n] = each positive followed by
 * its negatives, y the labels; the user's positives are pos_item[pos_first .. pos_first+n).  DEVICE outputs. */
int crb_sample_nais(crb_handle* h, uint64_t seed, uint32_t epoch, int64_t pos_first, int32_t n_pos_user, int32_t neg_ratio,
                    int32_t* targets, float* y, void* stream);

/* sess.run([train, loss], {u_idx: history, i_idx: targets, y}) for model/ranking/NAIS_single.py:59-90 (one step per user).
 * atten_concat = 0: atten_type 'prod', joint = q * p, W [dim, atten_size]; 1: atten_type 'concat', joint = [p ; q],
 * W [2*dim, atten_size] (NAIS_single.py:52-55, 67-71).  P, Q [(I+1), dim]; B the item bias as a dim-1 table padded to a multiple of
 * 4 rows; `dense` packs W row-major, b [atten_size], h [atten_size].  hist / targets / y are DEVICE arrays. */
int crb_train_step_nais(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_table* B, float* gradP, float* gradQ,
                        float* gradB, float* dense, float* dense_s1, float* dense_s2, int32_t atten_size, int32_t atten_concat,
                        const crb_opt* opt, const int32_t* hist, int32_t n_hist, const int32_t* targets, const float* y,
                        int32_t n_targets, float beta, float reg, double* loss_out, void* stream);

/* train_model_nais (RankingRecommender.py:64-87) for n_users users in one call: sampler + one step per user.
 * list_start / list_len: HOST arrays (offset and length of each user's interaction list inside pos_item, in the order the
 * reference iterates data.ui_train); loss_out: DEVICE double [n_users]; opt->step = index of the first step. */
int crb_train_epoch_nais(crb_handle* h, const crb_table* P, const crb_table* Q, const crb_table* B, float* gradP, float* gradQ,
                         float* gradB, float* dense, float* dense_s1, float* dense_s2, int32_t atten_size, int32_t atten_concat,
                         const crb_opt* opt, uint64_t seed, uint32_t epoch, const int64_t* list_start, const int32_t* list_len, int64_t n_users,
                         int32_t neg_ratio, float beta, float reg, double* loss_out, void* stream);

/* NAIS_single._predict (NAIS_single.py:92-97) for one user: scores[t] = s_t . q_t + bias_t over `targets`.  DEVICE buffers. */
int crb_score_nais(crb_handle* h, const float* P, const float* Q, const float* bias, const float* dense, int32_t dim,
                   int32_t atten_size, int32_t atten_concat, const int32_t* hist, int32_t n_hist, const int32_t* targets,
                   int32_t n_targets, float beta, float* scores, void* stream);

/* Multi-GPU BPR step, phase 1 (every rank, same step index): sess.run([train, loss]) on this rank's slice of the union batch
 * (model/ranking/BPR.py:31-44; the slicing of RankingRecommender.py:38-46).
 * P: this rank's user rows; u = local user rows, i/j = GLOBAL item ids (DEVICE or HOST), or u == NULL to sample rows
 * [first, first+batch) of this rank's epoch (crb_set_history holds the rank's own users, item ids global).
 * Kernels: item rows that occur more than once in this rank's batch are fetched from their owners ONCE into a local staging
 * buffer (item_fetch_kernel); the fused step (shard_step_kernel) reads the remaining item rows straight from the owner's HBM over
 * NVLink, applies user rows locally and sends the gradient of a once-occurring item row straight into its owner's direct-mapped
 * inbox; gradients of repeated item rows are summed locally in triplet order and sent once (dup_reduce_kernel<SHARD>).  The caller
 * then runs crb_shard_barrier, crb_shard_apply_inbox and a second crb_shard_barrier on the same stream. */
int crb_shard_step_compute(crb_handle* h, const crb_table* P, const crb_shard* shard, const crb_opt* opt, const int32_t* u,
                           const int32_t* i, const int32_t* j, uint64_t seed, uint32_t epoch, int64_t first, int32_t neg_ratio,
                           int64_t batch, float reg, double* loss_out, void* stream);
/* Optional phase 0 of the NEXT step on the handle's auxiliary stream: sample rows [first, first+batch), count the occurrences of
 * every user and item row, give repeated rows their gradient slots and resolve every triplet's item sources -- the index-only work
 * (utils/sampler.py:46-74 + the feed slicing of RankingRecommender.py:38-46) then overlaps the current step's barriers and
 * inbox phase.  Consumed by the next crb_shard_step_compute with u == NULL and the same (seed, epoch, first, neg_ratio, batch).
 * n_items: rows of the GLOBAL item table.  feed_u / feed_i / feed_j (HOST or DEVICE int32 [batch], or all NULL): stage the caller's
 * own triplets (the reference's sampler output sliced as RankingRecommender.py:40-42; local user rows, global item ids) instead
 * of sampling -- the copies then overlap the current step; the five scalars only serve as the ticket the consuming
 * crb_shard_step_compute presents. */
int crb_shard_step_prepare(crb_handle* h, const crb_table* P, uint64_t seed, uint32_t epoch, int64_t first, int32_t neg_ratio,
                           int64_t batch, int64_t n_items, const int32_t* feed_u, const int32_t* feed_i, const int32_t* feed_j,
                           void* stream);
/* The same multi-GPU step for the POINTWISE dot-product family (MF: kind CRB_SCORE_DOT, GMF: CRB_SCORE_GMF; GMF.py:37-49, the loop of
 * RankingRecommender.py:48-60): u = local user rows, i = GLOBAL item ids, y = labels (DEVICE or HOST), or u == NULL to sample rows
 * [first, first+batch) of this rank's pointwise epoch.  Item rows and gradients travel exactly as in crb_shard_step_compute.  GMF's
 * h is REPLICATED: every rank writes the sum of its gradient partials into slot `rank` of every peer's dense_inbox; after the
 * barrier crb_shard_apply_dense adds the n_ranks vectors in rank order and applies TF's dense optimizer -- the same update on
 * every rank, so the copies stay bit-identical. */
int crb_shard_step_compute_pointwise(crb_handle* h, int32_t kind, const crb_table* P, const crb_shard* shard, const float* hvec,
                                     const crb_opt* opt, int32_t loss_kind, const int32_t* u, const int32_t* i, const float* y,
                                     uint64_t seed, uint32_t epoch, int64_t first, int32_t neg_ratio, int64_t batch, float reg,
                                     double* loss_out, void* stream);
int crb_shard_apply_dense(crb_handle* h, const crb_shard* shard, const crb_opt* opt, float* hvec, float* h_s1, float* h_s2, int32_t dim,
                          void* stream);
/* phase 2 (after the barrier): one pass over this rank's item rows -- the gradients that arrived for a row are summed in source-rank
 * order and applied once (TF's de-duplicated sparse apply on the union batch); with CRB_ADAM_TF1 a row that received nothing
 * takes its decay-only step, so remote readers always find current weights. */
int crb_shard_apply_inbox(crb_handle* h, const crb_shard* shard, const crb_opt* opt, void* stream);
/* Cross-rank barrier ON THE STREAM, no host synchronisation and no collective library: a one-block kernel stores `ticket` into
 * word `rank` of every peer's flag array (system-scope release) and spins until every word of its own array has reached `ticket`.
 * Every rank must call it with the same strictly increasing tickets (>= 1).  A wait longer than timeout_ms sets the sticky error
 * word instead of hanging; crb_shard_check reports it. */
int crb_shard_barrier(crb_handle* h, const crb_shard* shard, uint32_t ticket, int32_t timeout_ms, void* stream);
/* Synchronises the stream and returns CRB_ERR_STATE if a barrier timed out or the sampler found no admissible negative for some
 * row since the last check (both sticky until read here). */
int crb_shard_check(crb_handle* h, const crb_shard* shard, void* stream);
/* Sampler attempt overflows (utils/sampler.py:58-61 has no such bound; ours is CRB_SAMPLER_MAX_BLOCKS Philox blocks per row)
 * since the last call; synchronises the stream and clears the counter. */
int crb_sampler_errors(crb_handle* h, uint32_t* n_rows, void* stream);

/* Peer-memory plumbing: cudaMalloc'd (zeroed) buffers that can be exported to the other ranks of the box through CUDA IPC. */
int crb_malloc(crb_handle* h, int64_t bytes, void** out);
int crb_free(crb_handle* h, void* p);
int crb_ipc_export(crb_handle* h, void* dev_ptr, unsigned char handle64[64]);
int crb_ipc_open(crb_handle* h, const unsigned char handle64[64], void** out);
int crb_ipc_close(crb_handle* h, void* p);

/* Bring every row of a CRB_ADAM_TF1 table up to `step` (call before reading w: evaluation, checkpoint). */
int crb_adam_flush(crb_handle* h, const crb_table* T, const crb_opt* opt, void* stream);

/* sess.run(self.pre_scores, {u_idx, i_idx}) on flattened (user, candidate) pairs -- test_model_loo,
 * model/RankingRecommender.py:257-278.  Canonical arithmetic: one sequential fp32 fma chain over k.
 * scores: float[n] DEVICE or HOST.  hvec: GMF h (kind GMF) or item bias (kind DOT_BIAS), else NULL. */
int crb_score_pairs(crb_handle* h, int32_t kind, const float* P, const float* Q, const float* hvec, int32_t dim,
                    const int32_t* u, const int32_t* i, int64_t n, float* scores, void* stream);

/* test_model_loo's scoring AND ranking in one kernel (model/RankingRecommender.py:257-288): for user segment k -- the pairs
 * (seg_user[k], items[offsets[k] .. offsets[k+1])) -- the K best positions inside the segment, np.argsort(-scores_u)[:K] under the
 * documented tie rule (ascending != 0: distance models), -1 padded.  A warp owns a user and keeps the running top-K in registers;
 * scores are the canonical chain of crb_score_pairs bit for bit but never reach HBM.  Buffers DEVICE or HOST.  Needs K <= 32,
 * dim % 4 == 0, dim <= 512 (else CRB_ERR_UNSUPPORTED: use crb_score_pairs + crb_topk_segments). */
int crb_score_pairs_topk(crb_handle* h, int32_t kind, const float* P, const float* Q, const float* hvec, int32_t dim,
                         const int32_t* seg_user, const int32_t* items, const int64_t* offsets, int64_t n_users, int32_t K,
                         int32_t ascending, int32_t* topk_pos, void* stream);

/* Lines :281-288 of test_model_loo for a ragged batch: user k owns scores[offsets[k]..offsets[k+1]); writes the
 * positions of its best K candidates (score descending -- ascending when `ascending` --, ties by position
 * ascending; -1 padded) to topk_pos[k*K..].  DEVICE or HOST buffers. */
int crb_topk_segments(crb_handle* h, const float* scores, const int64_t* offsets, int64_t n_users, int32_t K,
                      int32_t ascending, int32_t* topk_pos, void* stream);

/* test_model_rs, model/RankingRecommender.py:203-240: all-item scores of `users`, seen items (crb_set_history)
 * skipped, best K item ids per user.  `exact`=1: fp32 canonical arithmetic on CUDA cores.  `exact`=0: bf16
 * tcgen05 tensor-core pass produces certified candidate sets that are re-scored in canonical fp32, so the
 * returned ids are identical to exact=1 (users whose certificate fails are re-run exactly, on the GPU).
 * P: [*, dim] user vectors indexed by `users` (for FISM pass precomputed vectors and users = 0..n-1 with
 * hist_users = the real ids).  topk_items int32 [n_users*K], topk_scores float [n_users*K] (may be NULL). */
int crb_score_topk(crb_handle* h, int32_t kind, const float* P, const float* Q, const float* hvec,
                   int64_t n_items, int32_t dim, const int32_t* users, const int32_t* hist_users, int64_t n_users,
                   int32_t K, int32_t exact, int32_t* topk_items, float* topk_scores, void* stream);

/* statistics of the last crb_score_topk(exact=0) call: [0] users certified by the tensor-core pass,
 * [1] users re-run exactly, [2] candidate slots used (max over users). */
int crb_score_topk_stats(crb_handle* h, int64_t stats[4]);
/* crb_score_topk keeps the tensor-core path's bf16 copy of the item table between calls with the same (Q, hvec, n_items, dim,
 * kind); every library call that writes a table drops it.  A caller that writes Q / hvec ITSELF between two crb_score_topk calls
 * (e.g. restores a checkpoint into the same buffer) must call this. */
int crb_eval_cache_invalidate(crb_handle* h);

/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
int64_t crb_launch_count(crb_handle* h);

/* Measurement hook (bench.py's roofline): while enabled, the fused step kernel (K3, the dominant kernel) of every
 * training step is bracketed by CUDA events on its own stream.  crb_profile_read synchronises, returns the summed
 * kernel milliseconds and the number of launches since the last read, and resets the accumulators. */
int crb_profile_enable(crb_handle* h, int32_t on);
int crb_profile_read(crb_handle* h, double* step_kernel_ms, int64_t* n_launches);
/* The same for the other bracketed kernels of a step: tag 0 = fused step kernel, 1 = item_fetch_kernel (multi-GPU),
 * 2 = duplicate reduce / send pipeline, 3 = inbox_apply_kernel (multi-GPU). */
int crb_profile_read_tag(crb_handle* h, int32_t tag, double* ms, int64_t* n_launches);

#ifdef __cplusplus
}
#endif
#endif /* CLEVERREC_B200_H */
