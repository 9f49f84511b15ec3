"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Literal restatement (pure Python + NumPy, same control flow, same RNG calls) of the
reference's host-side hot-path code, so that it can travel to the GPU box where
``/root/reference`` does not exist.  Pinned against the genuine reference functions
(imported through ``oracle/refimport.py``) by ``tests/test_oracle_vs_reference.py`` in the
build container and against the committed fixtures in ``tests/golden/`` everywhere.

  pairwise_ranking_sampler   <- utils/sampler.py:46-74
  pointwise_ranking_sampler  <- utils/sampler.py:10-43
  ranking_sampler_cml        <- utils/sampler.py:77-99
  ranking_sampler_sbpr       <- utils/sampler.py:102-141
  get_SPu                    <- utils/tools.py:115-127
  nais_user_batches          <- model/RankingRecommender.py:64-87 (sampling part)
  cal_ranking_metrics        <- utils/metrics.py:9-19
  eval_loo / eval_rs         <- model/RankingRecommender.py:250-299 / 198-247 (everything after sess.run)
"""
import math
from collections import defaultdict

import numpy as np


def _negatives(seen_items, item_nums, neg_ratio):
    # utils/sampler.py:53-61
    random_j = set()
    out = []
    for _ in range(neg_ratio):
        j = np.random.randint(item_nums)
        while j in random_j or j in seen_items:
            j = np.random.randint(item_nums)
        random_j.add(j)
        out.append(j)
    return out


def pairwise_ranking_sampler(data, neg_ratio, batch_size, fism_like=False):
    u_f, i_f, j_f, nbr = [], [], [], []
    for u, items in data.ui_train.items():
        seen_items = set(items)
        for i in items:
            for j in _negatives(seen_items, data.item_nums, neg_ratio):
                u_f.append(u)
                i_f.append(i)
                j_f.append(j)
                nbr.append(len(seen_items))
    n = len(u_f)
    train_batches = math.ceil(n / batch_size)
    s_idx = np.random.permutation(n)
    out = (train_batches, np.array(u_f)[s_idx], np.array(i_f)[s_idx], np.array(j_f)[s_idx])
    if fism_like:
        out = out + (np.array(nbr)[s_idx],)
    return out


def pointwise_ranking_sampler(data, neg_ratio, batch_size):
    u_f, i_f, y = [], [], []
    for u, items in data.ui_train.items():
        seen_items = set(items)
        for i in items:
            u_f.append(u)
            i_f.append(i)
            y.append(1.0)
            for j in _negatives(seen_items, data.item_nums, neg_ratio):
                u_f.append(u)
                i_f.append(j)
                y.append(0.0)
    n = len(u_f)
    train_batches = math.ceil(n / batch_size)
    s_idx = np.random.permutation(n)
    return train_batches, np.array(u_f)[s_idx], np.array(i_f)[s_idx], np.array(y)[s_idx]


def ranking_sampler_cml(data, neg_ratio, batch_size):
    u_f, i_f, negs = [], [], []
    for u, items in data.ui_train.items():
        seen_items = set(items)
        for i in items:
            u_f.append(u)
            i_f.append(i)
            negs.append(_negatives(seen_items, data.item_nums, neg_ratio))
    n = len(u_f)
    train_batches = math.ceil(n / batch_size)
    s_idx = np.random.permutation(n)
    return train_batches, np.array(u_f)[s_idx], np.array(i_f)[s_idx], np.array(negs)[s_idx]


def get_SPu(data):
    # utils/tools.py:115-127: items of u's friends that u has not consumed; list(set) order is what the sampler indexes
    SPu = {}
    for u in data.ui_train:
        acc = set()
        if u in data.user_friends:
            for friend in data.user_friends[u]:
                if friend not in data.ui_train:
                    continue
                acc = acc.union(set(data.ui_train[friend])).difference(set(data.ui_train[u]))
            if acc:
                SPu[u] = list(acc)
    return SPu


def ranking_sampler_sbpr(data, SPu, neg_ratio, batch_size, is_suk=True):
    u_f, i_f, k_f, j_f, suk_f = [], [], [], [], []
    for u, items in data.ui_train.items():
        if u not in SPu:
            continue
        tru, spu = set(items), set(SPu[u])
        for i in items:
            for _ in range(neg_ratio):
                u_f.append(u)
                i_f.append(i)
                s = np.random.randint(len(spu))
                k_f.append(SPu[u][s])
                neg = np.random.randint(data.item_nums)
                while neg in tru or neg in spu:
                    neg = np.random.randint(data.item_nums)
                j_f.append(neg)
                if is_suk:
                    suk = 0
                    for friend in data.user_friends[u]:
                        if friend not in data.ui_train:
                            continue
                        if SPu[u][s] in data.ui_train[friend]:
                            suk += 1
                    suk_f.append(suk)
    n = len(u_f)
    train_batches = math.ceil(n / batch_size)
    s_idx = np.random.permutation(n)
    out = (train_batches, np.array(u_f)[s_idx], np.array(i_f)[s_idx], np.array(k_f)[s_idx], np.array(j_f)[s_idx])
    if is_suk:
        out = out + (np.array(suk_f)[s_idx],)
    return out


def nais_user_batches(data, neg_ratio):
    """Yields (u, history, i_idx, y) per user exactly as train_model_nais builds its feed."""
    for u, items in data.ui_train.items():
        seen_items = set(items)
        i_idx, y = [], []
        for i in items:
            i_idx.append(i)
            y.append(1.0)
            for j in _negatives(seen_items, data.item_nums, neg_ratio):
                i_idx.append(j)
                y.append(0.0)
        yield u, items, i_idx, y


def cal_ranking_metrics(real_items, rec_items, K):
    hit, mrr, dcg, idcg = 0, 0, 0, 0
    for id_ in range(len(real_items)):
        item = real_items[id_]
        if item in rec_items:
            hit += 1
            idx = np.where(rec_items == item)[0][0]
            mrr += 1.0 / (idx + 1)
            dcg += 1.0 / (np.log2(idx + 2))
        idcg += 1.0 / (np.log2(id_ + 2))
    return hit / min(K, len(real_items)), mrr, dcg / idcg


def argsort_desc_stable(scores):
    """Our documented tie rule: score descending, index ascending.  np.argsort's default
    introsort agrees whenever scores are distinct."""
    return np.argsort(-np.asarray(scores), kind="stable")


def eval_loo(test_users, ui_test, scores_per_user, neg_samples, topk, cml_like=False):
    """scores_per_user[u] = 1-D scores aligned with ui_test[u].  Mirrors RankingRecommender.py:281-298."""
    HR, MRR, NDCG = defaultdict(list), defaultdict(list), defaultdict(list)
    for u in test_users:
        s = np.asarray(scores_per_user[u])
        args_u = (np.argsort(s, kind="stable") if cml_like else argsort_desc_stable(s))[:topk[-1]]
        real_items = ui_test[u][neg_samples:]
        for kid in range(len(topk)):
            rec_items = np.take(ui_test[u], args_u[:topk[kid]])
            h, m, n = cal_ranking_metrics(real_items, rec_items, topk[kid])
            HR[kid].append(h)
            MRR[kid].append(m)
            NDCG[kid].append(n)
    return HR, MRR, NDCG


def topk_unseen(scores_u, seen, K, cml_like=False):
    """RankingRecommender.py:222-240 for one user: full argsort, skip seen, first K (float64 buffer as the reference)."""
    args_u = np.argsort(scores_u, kind="stable") if cml_like else argsort_desc_stable(scores_u)
    topk_items = np.zeros(K)
    if seen is None:
        return args_u[:K]
    count, j = 0, 0
    while count < K:
        if args_u[j] not in seen:
            topk_items[count] = args_u[j]
            count += 1
        j += 1
    return topk_items


def eval_rs(test_users, ui_train, ui_test, score_rows, topk, cml_like=False):
    """score_rows[k] = all-item scores of test_users[k].  Mirrors RankingRecommender.py:227-246."""
    HR, MRR, NDCG = defaultdict(list), defaultdict(list), defaultdict(list)
    for k, u in enumerate(test_users):
        seen = set(ui_train[u]) if u in ui_train else None
        topk_items = topk_unseen(score_rows[k], seen, topk[-1], cml_like)
        for kid in range(len(topk)):
            h, m, n = cal_ranking_metrics(ui_test[u], topk_items[:topk[kid]], topk[kid])
            HR[kid].append(h)
            MRR[kid].append(m)
            NDCG[kid].append(n)
    return HR, MRR, NDCG
