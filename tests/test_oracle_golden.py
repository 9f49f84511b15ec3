"""The oracle's literal restatement of the reference host code (oracle/ref_host.py) against the committed golden
vectors, which were produced by the GENUINE reference functions (oracle/make_golden.py).  Runs everywhere."""
import numpy as np

from conftest import Data
from oracle import ref_host as H


def _sub(split_loo, n):
    users = list(split_loo.ui_train.keys())[:n]
    return Data(split_loo.user_nums, split_loo.item_nums, {u: split_loo.ui_train[u] for u in users}, {})


def test_samplers_match_reference_stream(split_loo, golden):
    g = golden["sampler_seed3"]
    sub = _sub(split_loo, int(g["n_sub_users"]))
    np.random.seed(3)
    pw = H.pairwise_ranking_sampler(sub, 4, 6144, fism_like=True)
    pt = H.pointwise_ranking_sampler(sub, 2, 1000)
    cm = H.ranking_sampler_cml(sub, 5, 512)
    assert pw[0] == int(g["pw_batches"]) and pt[0] == int(g["pt_batches"]) and cm[0] == int(g["cm_batches"])
    for a, b in ((pw[1], "pw_u"), (pw[2], "pw_i"), (pw[3], "pw_j"), (pw[4], "pw_nbr"), (pt[1], "pt_u"), (pt[2], "pt_i"),
                 (cm[1], "cm_u"), (cm[2], "cm_i"), (cm[3], "cm_neg")):
        assert np.array_equal(np.asarray(a), g[b]), b
    assert np.array_equal(pt[3].astype(np.float32), g["pt_y"])


def test_metrics_known_answers(golden):
    g = golden["metrics"]
    for K, rec, real, out in zip(g["K"], g["rec"], g["real"], g["out"]):
        rec, real = rec[:K], [int(x) for x in real if x >= 0]
        got = H.cal_ranking_metrics(real, rec, int(K))
        assert tuple(got) == tuple(out.tolist())  # bit-identical float64


class _Fake(object):
    def __init__(self, seed, item_nums):
        self.rs, self.item_nums = np.random.RandomState(seed), item_nums


def test_eval_loo_matches_reference_loop(split_loo, golden):
    g = golden["eval_loops"]
    topk, bt = g["topk"].tolist(), int(g["batch_size_t"])
    rs = np.random.RandomState(21)
    users = list(split_loo.ui_test.keys())
    scores = {}
    for a in range(0, len(users), bt):  # the reference draws one flattened score vector per test batch
        cur = users[a:a + bt]
        flat = rs.rand(sum(len(split_loo.ui_test[u]) for u in cur)).astype(np.float32)
        p = 0
        for u in cur:
            n = len(split_loo.ui_test[u])
            scores[u] = flat[p:p + n]
            p += n
    HR, MRR, NDCG = H.eval_loo(users, split_loo.ui_test, scores, 99, topk)
    got = np.asarray([[HR[k], MRR[k], NDCG[k]] for k in range(len(topk))])
    assert np.array_equal(got, g["loo"])


def test_eval_rs_matches_reference_loop(split_rs, golden):
    g = golden["eval_loops"]
    topk, bt = g["topk"].tolist(), int(g["batch_size_t"])
    rs = np.random.RandomState(22)
    users = list(split_rs.ui_test.keys())
    rows = []
    for a in range(0, len(users), bt):
        rows.append(rs.rand(len(users[a:a + bt]), split_rs.item_nums).astype(np.float32))
    rows = np.concatenate(rows)
    HR, MRR, NDCG = H.eval_rs(users, split_rs.ui_train, split_rs.ui_test, rows, topk)
    got = np.asarray([[HR[k], MRR[k], NDCG[k]] for k in range(len(topk))])
    assert np.array_equal(got, g["rs"])
