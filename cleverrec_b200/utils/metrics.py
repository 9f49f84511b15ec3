# coding: utf-8
""" Evaluation metrics for ranking (host float64, formula of the reference utils/metrics.py:9-19).

The device returns integer ranks / item ids only; HR/MRR/NDCG arithmetic stays on the host in float64 with the
reference's operation order so the values are bit-identical (SURVEY.md 7 'Metric bit-identity')."""
import math

import numpy as np


# Calculate HR@K, MRR@K and NDCG@K  (same signature and quirks as the reference: MRR sums over all hits, IDCG
# sums over len(real_items) positions, HR divides by min(K, len(real_items)))
def cal_ranking_metrics(real_items, rec_items, K):
    hit, mrr, dcg, idcg = 0, 0, 0, 0
    rec_items = np.asarray(rec_items)
    for id in range(len(real_items)):
        item = real_items[id]
        where = np.where(rec_items == item)[0]
        if where.shape[0]:
            hit += 1
            idx = where[0]  # item's rank in predicted_items (first occurrence)
            mrr += 1.0 / (idx + 1)
            dcg += 1.0 / (np.log2(idx + 2))
        idcg += 1.0 / (np.log2(id + 2))
    return hit / min(K, len(real_items)), mrr, dcg / idcg


def batch_ranking_metrics(real_lists, rec_items, K):
    """Vectorised cal_ranking_metrics over users: real_lists = list of per-user real item lists, rec_items =
    [n_users, >=K] integer array (-1 padded).  Accumulates over the real items in list order, so every per-user
    value is bit-identical to cal_ranking_metrics(real_lists[k], rec_items[k, :K], K)."""
    n = len(real_lists)
    rec = np.asarray(rec_items)[:, :K]
    lens = np.fromiter((len(r) for r in real_lists), dtype=np.int64, count=n)
    max_len = int(lens.max()) if n else 0
    hit = np.zeros(n, dtype=np.int64)
    mrr, dcg, idcg = np.zeros(n), np.zeros(n), np.zeros(n)
    real = np.full((n, max_len), -2, dtype=np.int64)
    for k, r in enumerate(real_lists):
        real[k, :len(r)] = r
    for id in range(max_len):
        valid = lens > id
        eq = rec == real[:, id:id + 1]
        found = eq.any(axis=1) & valid
        idx = eq.argmax(axis=1)  # first occurrence
        hit += found
        mrr = np.where(found, mrr + 1.0 / (idx + 1), mrr)
        dcg = np.where(found, dcg + 1.0 / np.log2(idx + 2), dcg)
        idcg = np.where(valid, idcg + 1.0 / (np.log2(id + 2)), idcg)
    hr = hit / np.minimum(K, lens)
    return hr, mrr, dcg / idcg


# Calculate RMSE, MAE (reference utils/metrics.py:22-29)
def cal_rmse_mae(y, y_pre):
    abs_sum, square_sum = 0, 0
    for id in range(len(y)):
        res = y[id] - y_pre[id]
        abs_sum += abs(res)
        square_sum += res ** 2
    rmse, mae = math.sqrt(square_sum / len(y)), abs_sum / len(y)
    return rmse, mae
