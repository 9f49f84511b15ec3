"""-m gpu: every model class behind the reference's Recommender / RankingRecommender interface runs end to end
(train_model -> test_model_loo / test_model_rs -> run_model) on a synthetic data object with the reference's attribute
surface, and the evaluation of the dot-product-family models is bit-identical to the oracle's reference loop."""
import logging

import numpy as np
import pytest

from conftest import Data, synthetic_data
from oracle import c_oracle as O
from oracle import ref_host as H

pytestmark = pytest.mark.gpu

BASE = {'model_type': 'ranking', 'saved_dir': './saved_model', 'data.split_way': 'loo', 'test.neg_samples': '49', 'test.batch_size': '64',
        'test.interval': '1', 'topk': '[5,10]', 'epoches': '2', 'batch_size': '512', 'lr': '0.01', 'neg_ratio': '3', 'optimizer': 'Adam',
        'init_method': 'normal', 'stddev': '0.05', 'seed': '3'}
CONFS = {
    'BPR': {'embed_size': '32', 'reg': '0.01', 'is_pairwise': 'True', 'loss_func': 'bpr'},
    'MF': {'embed_size': '32', 'reg_mf': '1e-3', 'is_pairwise': 'False', 'loss_function': "'square'", 'loss_func': 'square'},
    'GMF': {'embed_size': '32', 'reg_gmf': '1e-2', 'is_pairwise': 'False', 'loss_func': 'cross_entropy', 'init_method': 'xavier_uniform'},
    'NeuMF': {'embed_size': '16', 'layers': '[64,32,16]', 'reg_gmf': '1e-2', 'reg_mlp': '1e-3', 'is_pairwise': 'False', 'loss_func': 'cross_entropy',
              'init_method': 'xavier_uniform '},
    'CML': {'embed_size': '32', 'margin': '1.0', 'reg': '10.0', 'cml_like': 'True', 'is_pairwise': 'False', 'loss_func': 'hinge', 'init_method': 'xavier ',
            'neg_ratio': '10', 'lr': '0.003'},
    'FISM': {'embed_size': '32', 'alpha': '0.4', 'reg': '1e-3', 'reg_bias': '1e-3', 'fism_like': 'True', 'is_pairwise': 'True', 'loss_func': 'bpr',
             'init_method': 'xavier_uniform'},
    'NAIS_single': {'embed_size': '32', 'atten_size': '16', 'atten_type': "'prod'", 'beta': '0.5', 'reg': '1e-3', 'lr': '0.01', 'optimizer': 'Adagrad',
                    'nais_like': 'True', 'is_pairwise': 'False', 'loss_func': 'cross_entropy', 'init_method': 'xavier_uniform'},
}


def _data(split, neg_samples):
    d = synthetic_data(120, 300, 12, seed=11, test_per_user=2 if split == 'rs' else 1)
    if split == 'loo':  # negatives then the positive, as RankingPreprocess.py:121-129 builds ui_test
        rs = np.random.RandomState(0)
        for u in d.ui_test:
            cand = np.setdiff1d(np.arange(d.item_nums), d.ui_train[u])
            d.ui_test[u] = rs.choice(cand, neg_samples, replace=False).tolist() + d.ui_test[u]
    return d


def _model(name, data, **over):
    import importlib
    cfg = dict(BASE, recommender=name)
    cfg.update(CONFS[name])
    cfg.update({k: str(v) for k, v in over.items()})
    cls = getattr(importlib.import_module('cleverrec_b200.model.ranking.' + name), name)
    m = cls(None, data, cfg, logging.getLogger('test'))
    m.build_model()
    return m


@pytest.mark.parametrize("name", sorted(CONFS))
def test_train_and_loo_eval(name):
    data = _data('loo', 49)
    m = _model(name, data)
    losses = [m.train_model() for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    HR, MRR, NDCG = m.test_model_loo()
    assert len(HR[0]) == len(HR[1]) == len(m.test_users) and set(HR.keys()) == {0, 1}
    assert all(0.0 <= x <= 1.0 for x in HR[1]) and all(0.0 <= x <= 1.0 + 1e-12 for x in NDCG[1])
    # (the synthetic interactions carry no signal; ranking quality is asserted on the real ml-100k split in test_gpu_driver.py)


@pytest.mark.parametrize("name", ["BPR", "GMF", "CML", "FISM", "NeuMF", "NAIS_single"])
def test_train_and_rs_eval(name):
    data = _data('rs', 0)
    m = _model(name, data, **{'data.split_way': 'rs', 'test.neg_samples': 0})
    m.train_model()
    HR, MRR, NDCG = m.test_model_rs()
    assert len(HR[0]) == len(m.test_users) and all(np.isfinite(NDCG[1]))


@pytest.mark.parametrize("name,kind,asc", [("GMF", 1, False), ("MF", 0, False), ("CML", 2, True)])
def test_loo_eval_bit_identical_to_reference_loop(name, kind, asc):
    data = _data('loo', 49)
    m = _model(name, data)
    m.train_model()
    HR, MRR, NDCG = m.test_model_loo()
    P, Q = m.P.w.cpu().numpy(), m.Q.w.cpu().numpy()
    hvec = m.h_gmf.cpu().numpy() if name == "GMF" else None
    scores = {u: O.score_pairs(kind, P, Q, np.full(len(data.ui_test[u]), u), np.asarray(data.ui_test[u]), hvec) for u in m.test_users}
    oHR, oMRR, oNDCG = H.eval_loo(m.test_users, data.ui_test, scores, 49, m.topk, cml_like=asc)
    for k in range(len(m.topk)):
        assert HR[k] == oHR[k] and MRR[k] == oMRR[k] and NDCG[k] == oNDCG[k]


def test_fism_rs_eval_bit_identical(name="FISM"):
    data = _data('rs', 0)
    m = _model(name, data, **{'data.split_way': 'rs', 'test.neg_samples': 0})
    m.train_model()
    HR, MRR, NDCG = m.test_model_rs()
    S, Q, b = m._S.cpu().numpy(), m.Q.w.cpu().numpy(), m.b.cpu().numpy()
    I = data.item_nums
    rows = np.repeat(np.arange(len(m.test_users), dtype=np.int32), I)
    items = np.tile(np.arange(I, dtype=np.int32), len(m.test_users))
    sc = O.score_pairs(3, S, Q, rows, items, b).reshape(len(m.test_users), I)
    oHR, oMRR, oNDCG = H.eval_rs(m.test_users, data.ui_train, data.ui_test, sc, m.topk)
    for k in range(len(m.topk)):
        assert HR[k] == oHR[k] and NDCG[k] == oNDCG[k]


def test_run_model_all_log_lines(caplog):
    data = _data('loo', 49)
    m = _model('GMF', data)
    with caplog.at_level(logging.INFO):
        best_epoch, best = m.run_model()
    assert 'Training loss:' in caplog.text and 'best_epoch:' in caplog.text and best_epoch >= 1


def test_numpy_stream_training_follows_the_reference_triplets():
    """sampler=numpy_stream: after np.random.seed(s) the model trains on exactly the triplet sequence the reference's sampler
    produces, so an epoch equals the restated TF graph fed with the restated reference sampler's batches."""
    import torch
    from oracle import tf1_restatement as T
    data = _data('loo', 49)
    m = _model('BPR', data, sampler='numpy_stream', optimizer='SGD', lr=0.05, batch_size=256)
    P0, Q0 = m.P.w.cpu().clone(), m.Q.w.cpu().clone()
    np.random.seed(77)
    got = m.train_model()
    state_after = np.random.get_state()
    np.random.seed(77)
    tr = H.pairwise_ranking_sampler(data, m.neg_ratio, 256)
    assert np.array_equal(np.random.get_state()[1], state_after[1]) and np.random.get_state()[2] == state_after[2]
    ref, ropt, total = {"P": P0, "Q": Q0}, T.TF1Optimizer("SGD", 0.05), 0.0
    for k in range(tr[0]):
        sl = slice(k * 256, (k + 1) * 256)
        b = {"u": torch.tensor(tr[1][sl]), "i": torch.tensor(tr[2][sl]), "j": torch.tensor(tr[3][sl])}
        total += T.train_step(T.bpr_loss, ref, b, {"reg": m.reg}, ropt, sparse_index={"P": ["u"], "Q": ["i", "j"]})
    assert abs(got - total / tr[0]) <= 1e-5 * abs(total / tr[0])
    np.testing.assert_allclose(m.P.w.cpu().numpy(), ref["P"].numpy(), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(m.Q.w.cpu().numpy(), ref["Q"].numpy(), rtol=2e-5, atol=1e-6)
