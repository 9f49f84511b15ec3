"""-m gpu: the drop-in classes (cleverrec_b200.model.ranking.BPR behind Recommender / RankingRecommender) on the
golden ml-100k splits produced by the reference's own preprocessing."""
import logging

import numpy as np
import pytest
import torch

from oracle import c_oracle as O
from oracle import philox as X
from oracle import ref_host as H

pytestmark = pytest.mark.gpu

CFG = {'recommender': 'BPR', 'model_type': 'ranking', 'saved_dir': './saved_model', 'data.split_way': 'loo', 'test.neg_samples': '99',
       'test.batch_size': '1024', 'test.interval': '1', 'topk': '[10,20]', 'epoches': '2', 'batch_size': '6144', 'embed_size': '64',
       'reg': '0.01', 'lr': '0.001', 'neg_ratio': '4', 'optimizer': 'Adam', 'is_pairwise': 'True', 'loss_func': 'bpr',
       'init_method': 'normal ', 'stddev': '0.01', 'seed': '5'}


def _model(data, **over):
    from cleverrec_b200.model.ranking.BPR import BPR
    cfg = dict(CFG)
    cfg.update({k: str(v) for k, v in over.items()})
    m = BPR(None, data, cfg, logging.getLogger("test"))
    m.build_model()
    return m


def test_loo_eval_bit_identical_to_reference_loop(split_loo):
    m = _model(split_loo)
    l1 = m.train_model()
    l2 = m.train_model()
    assert np.isfinite(l1) and l2 < l1
    HR, MRR, NDCG = m.test_model_loo()
    # oracle: canonical scores (C) on the trained tables, then the reference's own ranking/metric loop (restated)
    P, Q = m.P.w.cpu().numpy(), m.Q.w.cpu().numpy()
    scores = {}
    for u in m.test_users:
        items = np.asarray(split_loo.ui_test[u])
        scores[u] = O.score_pairs(0, P, Q, np.full(items.shape[0], u), items)
    oHR, oMRR, oNDCG = H.eval_loo(m.test_users, split_loo.ui_test, scores, 99, m.topk)
    for k in range(len(m.topk)):
        assert HR[k] == oHR[k] and MRR[k] == oMRR[k] and NDCG[k] == oNDCG[k]
    assert 0.0 < np.mean(HR[0]) <= 1.0


@pytest.mark.parametrize("exact", ["True", "False"])
def test_rs_eval_bit_identical_to_reference_loop(split_rs, exact):
    m = _model(split_rs, **{'data.split_way': 'rs', 'test.neg_samples': 0, 'optimizer': 'Adagrad', 'lr': 0.05, 'score_exact': exact})
    m.train_model()
    HR, MRR, NDCG = m.test_model_rs()
    P, Q = m.P.w.cpu().numpy(), m.Q.w.cpu().numpy()
    users = np.asarray(m.test_users, dtype=np.int32)
    uu = np.repeat(users, split_rs.item_nums)
    ii = np.tile(np.arange(split_rs.item_nums, dtype=np.int32), users.shape[0])
    rows = O.score_pairs(0, P, Q, uu, ii).reshape(users.shape[0], split_rs.item_nums)
    oHR, oMRR, oNDCG = H.eval_rs(m.test_users, split_rs.ui_train, split_rs.ui_test, rows, m.topk)
    for k in range(len(m.topk)):
        assert HR[k] == oHR[k] and MRR[k] == oMRR[k] and NDCG[k] == oNDCG[k]


def test_run_model_logs_like_the_reference(split_loo, caplog):
    m = _model(split_loo, epoches=2)
    with caplog.at_level(logging.INFO):
        best_epoch, best = m.run_model()
    text = caplog.text
    assert ' epoch 1\n  Training loss: ' in text and '  (k=10) HR=' in text and 'best_epoch: ' in text
    assert best_epoch in (1, 2) and set(best.keys()) == {0, 1}


def test_training_improves_ranking_quality(split_loo):
    m = _model(split_loo, lr=0.01, epoches=1)
    hr0 = np.mean(m.test_model_loo()[0][0])
    for _ in range(8):
        m.train_model()
    hr1 = np.mean(m.test_model_loo()[0][0])
    assert hr1 > hr0 + 0.1  # random init ~0.10 -> learned


def test_native_history_path_trains_identically(split_loo):
    """data.train_rows (packaged RankingPreprocess) -> Engine.build_history gives the same epoch, losses and tables as the dict path."""
    from conftest import Data
    tu = np.repeat(np.asarray(list(split_loo.ui_train.keys()), dtype=np.int32), [len(v) for v in split_loo.ui_train.values()])
    ti = np.asarray([i for v in split_loo.ui_train.values() for i in v], dtype=np.int32)
    with_rows = Data(split_loo.user_nums, split_loo.item_nums, split_loo.ui_train, split_loo.ui_test)
    with_rows.train_rows = (tu, ti)
    a, b = _model(split_loo, optimizer='Adagrad', lr=0.05), _model(with_rows, optimizer='Adagrad', lr=0.05)
    # same sampled epoch (the sampler is integer-exact), so the two runs differ only by the summation order of hub rows with more
    # than 32 occurrences in a batch (slot order, SURVEY 8e 'Determinism'): equal to fp32 round-off
    ua, ia, ja = a.engine.sample_pairwise(5, 0, 0, 4096, 4)
    ub, ib, jb = b.engine.sample_pairwise(5, 0, 0, 4096, 4)
    assert torch.equal(ua, ub) and torch.equal(ia, ib) and torch.equal(ja, jb)
    la, lb = a.train_model(), b.train_model()
    assert abs(la - lb) <= 1e-7 * abs(la)
    assert torch.allclose(a.P.w, b.P.w, rtol=1e-5, atol=1e-7) and torch.allclose(a.Q.w, b.Q.w, rtol=1e-5, atol=1e-7)
    ha, hb = a.test_model_loo(), b.test_model_loo()
    assert abs(np.mean(ha[0][0]) - np.mean(hb[0][0])) < 0.01
