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
    'MLP': {'layers': '[64,32,16]', 'reg_mlp': '1e-2', 'is_pairwise': 'False', 'loss_func': 'cross_entropy', 'init_method': 'xavier_uniform'},
    'CML': {'embed_size': '32', 'margin': '1.0', 'reg': '10.0', 'cml_like': 'True', 'is_pairwise': 'False', 'loss_func': 'hinge', 'init_method': 'xavier ',
            'neg_ratio': '10', 'lr': '0.003'},
    'TransCF': {'embed_size': '32', 'margin': '0.5', 'reg1': '0.1', 'reg2': '0.01', 'cml_like': 'True', 'is_pairwise': 'True', 'loss_func': 'hinge',
                'lr': '0.003'},
    'LRML': {'embed_size': '32', 'mem_size': '10', 'margin': '0.2', 'reg': '0.001', 'cml_like': 'True', 'is_pairwise': 'True', 'loss_func': 'hinge',
             'lr': '0.003'},
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


@pytest.mark.parametrize("name", ["BPR", "GMF", "CML", "FISM", "NeuMF", "MLP", "NAIS_single", "TransCF", "LRML"])
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


# ------------------------------------------------------------------------------------------------ checkpoints / pretraining
def test_checkpoint_round_trip_every_model(tmp_path):
    """save_model() writes the reference's Saver variable names; the arrays are the (flushed) tables."""
    from cleverrec_b200.utils.tools import latest_checkpoint, load_checkpoint
    data = _data('loo', 49)
    names = {'BPR': {'BPR_params/P', 'BPR_params/Q'}, 'GMF': {'GMF_params/P', 'GMF_params/Q', 'GMF_params/h_gmf'},
             'CML': {'cml_params/P', 'cml_params/Q'}, 'FISM': {'FISM_paras/P', 'FISM_params/Q', 'FISM_params/b'},
             'NAIS_single': {'NAIS_paras/P', 'NAIS_params/Q', 'NAIS_params/bias', 'NAIS_params/W', 'NAIS_params/b', 'NAIS_params/h'}}
    for name, want in names.items():
        m = _model(name, data, saved_dir=str(tmp_path))
        m.train_model()
        path = m.save_model()
        assert path == latest_checkpoint(str(tmp_path / name))
        v = load_checkpoint(path)
        assert set(v) == want, name
        for key, arr in v.items():
            assert arr.dtype == np.float32 and np.all(np.isfinite(arr))
        first = sorted(want)[0]
        table = m.P.w if first.endswith('/P') else None
        if table is not None:
            assert np.array_equal(v[first], table.cpu().numpy())   # after the flush: what evaluation would read


def test_run_model_saves_on_best_when_asked(tmp_path):
    from cleverrec_b200.utils.tools import latest_checkpoint
    m = _model('BPR', _data('loo', 49), saved_dir=str(tmp_path), save_model='True', epoches=2)
    m.run_model()
    assert latest_checkpoint(str(tmp_path / 'BPR')) is not None
    m2 = _model('BPR', _data('loo', 49), saved_dir=str(tmp_path / 'off'), epoches=1)   # default: the reference never saves
    m2.run_model()
    assert latest_checkpoint(str(tmp_path / 'off' / 'BPR')) is None


def test_nais_starts_from_trained_fism(tmp_path):
    """NAIS_single.py:35-38: P, Q, bias restored from the FISM checkpoint by the reference's variable names."""
    data = _data('loo', 49)
    f = _model('FISM', data, saved_dir=str(tmp_path))
    for _ in range(3):
        f.train_model()
    f.save_model()
    n = _model('NAIS_single', data, fism_pretrain=str(tmp_path / 'FISM'))
    assert np.array_equal(n.P.w.cpu().numpy(), f.P.w.cpu().numpy()) and np.array_equal(n.Q.w.cpu().numpy(), f.Q.w.cpu().numpy())
    assert np.array_equal(n.bias.cpu().numpy(), f.b.cpu().numpy())
    assert np.isfinite(n.train_model())
    fresh = _model('NAIS_single', data, fism_pretrain=str(tmp_path / 'nothing_here'))   # no checkpoint: logged, trains from scratch
    assert not np.array_equal(fresh.P.w.cpu().numpy(), f.P.w.cpu().numpy())


def test_neumf_starts_from_gmf_and_mlp_checkpoints(tmp_path):
    """NeuMF.py:46-56,127-139: branches restored by name, h_neumf = 0.5 * concat(h_gmf, h_mlp).  The MLP branch comes from a NeuMF
    checkpoint re-keyed to the 'MLP_params/*' names (the standalone MLP model is outside the hot-path scope)."""
    from cleverrec_b200.utils.tools import load_checkpoint, save_checkpoint
    data = _data('loo', 49)
    g = _model('GMF', data, saved_dir=str(tmp_path), embed_size=16)
    g.train_model()
    g.save_model()
    donor = _model('NeuMF', data, saved_dir=str(tmp_path))
    donor.train_model()
    v = load_checkpoint(donor.save_model())
    mlp = {'MLP_params/P': v['NeuMF_params/P_mlp'], 'MLP_params/Q': v['NeuMF_params/Q_mlp'], 'MLP_params/h_mlp': v['NeuMF_params/h_mlp']}
    for k in range(3):
        mlp['MLP_params/W_%d' % k], mlp['MLP_params/b_%d' % k] = v['NeuMF_params/W_%d' % k], v['NeuMF_params/b_%d' % k]
    save_checkpoint(str(tmp_path / 'MLP'), 'MLP', mlp)
    m = _model('NeuMF', data, gmf_pretrain=str(tmp_path / 'GMF'), mlp_pretrain=str(tmp_path / 'MLP'))
    assert np.array_equal(m.P_gmf.w.cpu().numpy(), g.P.w.cpu().numpy()) and np.array_equal(m.Q_mlp.w.cpu().numpy(), v['NeuMF_params/Q_mlp'])
    layout, _ = m.dense_layout()
    off, shape = layout['h_neumf']
    want_h = 0.5 * np.concatenate([g.h_gmf.cpu().numpy(), v['NeuMF_params/h_mlp']])
    assert np.array_equal(m.dense[off:off + shape[0]].cpu().numpy(), want_h.astype(np.float32))
    off, shape = layout['W_1']
    assert np.array_equal(m.dense[off:off + shape[0] * shape[1]].cpu().numpy().reshape(shape), v['NeuMF_params/W_1'])
    assert np.isfinite(m.train_model())


def test_neumf_pretraining_chain_gmf_plus_mlp(tmp_path):
    """The NCF recipe the reference's confs describe (conf/NeuMF.properties gmf_pretrain / mlp_pretrain): train GMF and MLP, save,
    start NeuMF from both.  NeuMF's first logits are then 0.5 * (GMF logit + MLP logit) (NeuMF.py:56)."""
    import torch
    data = _data('loo', 49)
    g = _model('GMF', data, saved_dir=str(tmp_path), embed_size=16)
    m = _model('MLP', data, saved_dir=str(tmp_path))
    for _ in range(2):
        g.train_model(); m.train_model()
    g.save_model(); m.save_model()
    n = _model('NeuMF', data, gmf_pretrain=str(tmp_path / 'GMF'), mlp_pretrain=str(tmp_path / 'MLP'))
    u = torch.arange(0, 50, dtype=torch.int32, device='cuda')
    i = torch.arange(100, 150, dtype=torch.int32, device='cuda')
    lg = g.engine.score_pairs(1, g.P.w, g.Q.w, u, i, hvec=g.h_gmf)
    lm = m.engine.score_pairs_neumf(m.tabs, m.dense, len(m.layers), u, i)
    ln = n.engine.score_pairs_neumf(n.tabs, n.dense, len(n.layers), u, i)
    assert torch.allclose(ln, 0.5 * (lg + lm), rtol=1e-5, atol=1e-6)
    assert np.isfinite(n.train_model())


@pytest.mark.parametrize("name", ["BPR", "FISM"])
def test_losses_the_reference_graph_cannot_build_are_rejected(name):
    """BPR.py:42 / FISM.py:59 call get_loss(loss_func, ui - uj) with neither margin nor logits: only 'bpr' builds in the reference, and
    the mirror classes say so at construction instead of silently training another loss."""
    with pytest.raises(ValueError):
        _model(name, _data('loo', 49), loss_func='hinge')
