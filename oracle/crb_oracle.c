/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
 *
 * Plain-C restatement of the evaluation arithmetic of the hot path, in the "canonical" order the CUDA kernels
 * use (one sequential fp32 fma chain over k per (user,item) pair), so ranks and top-K ids can be compared
 * bit-for-bit.  Built by oracle/build_oracle.py with `gcc -O2 -ffp-contract=off` (fmaf is correctly rounded
 * whether glibc dispatches to the FMA instruction or to its software path).
 *
 *   oracle_score_pairs    <- each model's `_predict` on flattened pairs: BPR.py:49, GMF.py:43 (the logit),
 *                            CML.py:82, FISM.py:53; fed as in model/RankingRecommender.py:257-278
 *   oracle_topk_segments  <- np.argsort(-pre_scores_u)[:topk[-1]]            RankingRecommender.py:281-288
 *   oracle_score_pairs_neumf <- NeuMF._get_logits (NeuMF.py:63-85; MLP.py:44-53 when E = 0) on flattened pairs, the logit
 *   oracle_fullrank_topk  <- matmul + np.argsort + skip ui_train[u] + first K RankingRecommender.py:203-240
 * Tie rule: score descending (ascending for distance models), index ascending; np.argsort's
 * default introsort gives the same order whenever scores are distinct.
 * PARITY NOTE: the TF arithmetic these lines stand for is unpinned by the reference (no tests, TF absent).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { SCORE_DOT = 0, SCORE_GMF = 1, SCORE_SQDIST = 2, SCORE_DOT_BIAS = 3 };

static float canonical_score(int kind, const float* p, const float* q, const float* hvec, int32_t item, int dim) {
    float acc = 0.f;
    for (int k = 0; k < dim; ++k) {
        if (kind == SCORE_DOT || kind == SCORE_DOT_BIAS) {
            acc = fmaf(p[k], q[k], acc);
        } else if (kind == SCORE_GMF) {
            volatile float pq = p[k] * q[k]; /* rounded product first: einsum('ab,b->a', u*i, h) */
            acc = fmaf(pq, hvec[k], acc);
        } else {
            volatile float d = p[k] - q[k];
            acc = fmaf(d, d, acc);
        }
    }
    if (kind == SCORE_DOT_BIAS) acc = acc + hvec[item];
    return acc;
}

void oracle_score_pairs(int kind, const float* P, const float* Q, const float* hvec, int dim, const int32_t* u,
                        const int32_t* it, int64_t n, float* out) {
    for (int64_t k = 0; k < n; ++k)
        out[k] = canonical_score(kind, P + (int64_t)u[k] * dim, Q + (int64_t)it[k] * dim, hvec, it[k], dim);
}

typedef struct { float s; int32_t idx; } cand_t;
static int g_asc = 0;
static int cmp_cand(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a;
    const cand_t* y = (const cand_t*)b;
    if (x->s != y->s) {
        if (g_asc) return x->s < y->s ? -1 : 1;
        return x->s > y->s ? -1 : 1;
    }
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

void oracle_topk_segments(const float* scores, const int64_t* offsets, int64_t n_users, int K, int ascending, int32_t* out) {
    g_asc = ascending;
    for (int64_t usr = 0; usr < n_users; ++usr) {
        const int64_t lo = offsets[usr], n = offsets[usr + 1] - lo;
        cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (n > 0 ? n : 1));
        for (int64_t p = 0; p < n; ++p) { c[p].s = scores[lo + p]; c[p].idx = (int32_t)p; }
        qsort(c, (size_t)n, sizeof(cand_t), cmp_cand);
        for (int r = 0; r < K; ++r) out[usr * K + r] = r < n ? c[r].idx : -1;
        free(c);
    }
}

static int is_seen(const int32_t* cols, int64_t lo, int64_t hi, int32_t v) {
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cols[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo < end && cols[lo] == v;
}

/* users[k]: row of P; hist_users[k] (or users[k] when NULL): row of the seen CSR, <0 = no history */
void oracle_fullrank_topk(int kind, const float* P, const float* Q, const float* hvec, int64_t n_items, int dim,
                          const int32_t* users, const int32_t* hist_users, int64_t n_users, const int64_t* seen_rowptr,
                          const int32_t* seen_cols, int K, int32_t* out_items, float* out_scores) {
    g_asc = kind == SCORE_SQDIST;
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)n_items);
    for (int64_t k = 0; k < n_users; ++k) {
        const int32_t hu = hist_users ? hist_users[k] : users[k];
        const int64_t lo = hu >= 0 ? seen_rowptr[hu] : 0, hi = hu >= 0 ? seen_rowptr[hu + 1] : 0;
        int64_t n = 0;
        for (int64_t it = 0; it < n_items; ++it) {
            if (is_seen(seen_cols, lo, hi, (int32_t)it)) continue;
            c[n].s = canonical_score(kind, P + (int64_t)users[k] * dim, Q + it * dim, hvec, (int32_t)it, dim);
            if (c[n].s == 0.f) c[n].s = 0.f; /* -0 == +0 */
            c[n].idx = (int32_t)it;
            ++n;
        }
        qsort(c, (size_t)n, sizeof(cand_t), cmp_cand);
        for (int r = 0; r < K; ++r) {
            out_items[k * K + r] = r < n ? c[r].idx : -1;
            if (out_scores) out_scores[k * K + r] = r < n ? c[r].s : 0.f;
        }
    }
    free(c);
}

/* NeuMF / MLP logit of flattened pairs in the canonical order of csrc/train_neumf.cu: per layer acc = b[o], then
 * acc = fma(x[k], W[k][o], acc) for k ascending, ReLU; logit = one chain over the rounded GMF products p*q times h[0..E), then
 * over the tower output times h[E..).  `dense` packs W_l [n_in, n_in/2] row-major, b_l per layer, then h.  E = 0: the MLP model. */
void oracle_score_pairs_neumf(const float* Pg, const float* Qg, const float* Pm, const float* Qm, const float* dense, int E, int Em,
                              int n_layers, const int32_t* u, const int32_t* it, int64_t n, float* out) {
    const int L0 = 2 * Em;
    float* x = (float*)malloc(sizeof(float) * (size_t)L0);
    float* z = (float*)malloc(sizeof(float) * (size_t)L0);
    for (int64_t t = 0; t < n; ++t) {
        const int64_t uu = u[t], ii = it[t];
        for (int k = 0; k < L0; ++k) x[k] = k < Em ? Pm[uu * Em + k] : Qm[ii * Em + (k - Em)];
        int off = 0, ni = L0;
        for (int l = 0; l < n_layers; ++l) {
            const int no = ni / 2;
            const float* W = dense + off;
            const float* b = W + (size_t)ni * no;
            for (int o = 0; o < no; ++o) {
                float acc = b[o];
                for (int k = 0; k < ni; ++k) acc = fmaf(x[k], W[(size_t)k * no + o], acc);
                z[o] = acc > 0.f ? acc : 0.f;
            }
            memcpy(x, z, sizeof(float) * (size_t)no);
            off += ni * no + no;
            ni = no;
        }
        const float* h = dense + off;
        float acc = 0.f;
        for (int k = 0; k < E; ++k) {
            volatile float pq = Pg[uu * E + k] * Qg[ii * E + k];
            acc = fmaf(pq, h[k], acc);
        }
        for (int k = 0; k < ni; ++k) acc = fmaf(x[k], h[E + k], acc);
        out[t] = acc;
    }
    free(x);
    free(z);
}
