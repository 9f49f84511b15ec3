"""-m gpu: NAIS_single (csrc/train_nais.cu) against the torch restatement of model/ranking/NAIS_single.py:59-97."""
import numpy as np
import pytest
import torch

from conftest import synthetic_data
from oracle import philox as X
from oracle import tf1_restatement as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cleverrec_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("kind", ["SGD", "Adagrad", "Adam"])
@pytest.mark.parametrize("d,A,atten", [(32, 16, "prod"), (128, 32, "prod"), (32, 16, "concat"), (128, 32, "concat")])
def test_nais_user_steps(eng, kind, d, A, atten):
    """Both attention types of NAIS_single.py:66-71: 'prod' (joint = q * p, W [d, A]) and 'concat' (joint = [p ; q], W [2d, A])."""
    from cleverrec_b200.engine import Optimizer, Table
    I = 60
    n = I + 1
    concat = atten == "concat"
    jd = 2 * d if concat else d
    g = torch.Generator().manual_seed(d)
    rnd = lambda *s: torch.randn(*s, generator=g) * 0.2
    ref = {"P": rnd(n, d), "Q": rnd(n, d), "bias": rnd(n) * 0.5, "W": rnd(jd, A), "b_att": rnd(A) * 0.5, "h": rnd(A)}
    lr = 0.02 if kind != "Adam" else 0.005
    opt, ropt = Optimizer(kind, lr, adam_mode="lazy"), T.TF1Optimizer(kind, lr, adam_mode="tf1")
    P, Q = Table(ref["P"].clone().cuda(), kind, "lazy"), Table(ref["Q"].clone().cuda(), kind, "lazy")
    B = Table(torch.cat([ref["bias"], torch.zeros((-n) % 4)]).reshape(-1, 1).cuda().contiguous(), kind, "lazy")
    dense = torch.cat([ref["W"].reshape(-1), ref["b_att"], ref["h"]]).cuda()
    s1 = torch.full_like(dense, 0.1) if kind == "Adagrad" else (torch.zeros_like(dense) if kind == "Adam" else None)
    s2 = torch.zeros_like(dense) if kind == "Adam" else None
    rs = np.random.RandomState(A)
    hp = {"reg": 1e-3, "beta": 0.5, "atten_type": atten}
    for n_hist in (5, 1, 40):
        hist = rs.choice(I, n_hist, replace=False)
        tg = rs.randint(0, I, n_hist * 3)
        y = (rs.rand(n_hist * 3) < 0.3).astype(np.float32)
        got = eng.train_step_nais(P, Q, B, dense, s1, s2, A, opt, hist, tg, y, 0.5, 1e-3, concat=concat)
        b = {"hist": torch.tensor(hist), "i": torch.tensor(tg), "y": torch.tensor(y)}
        want = T.train_step(T.nais_loss, ref, b, hp, ropt, sparse_index={"P": ["hist"], "Q": ["i"], "bias": ["i"]})
        assert abs(got - want) <= 5e-5 * abs(want), (got, want)
    rtol, atol = (3e-4, 3e-5) if kind == "Adam" else (3e-5, 2e-6)
    cur = {"P": P.w.cpu().numpy(), "Q": Q.w.cpu().numpy(), "bias": B.w.cpu().numpy().reshape(-1)[:n], "W": dense[:jd * A].cpu().numpy().reshape(jd, A),
           "b_att": dense[jd * A:jd * A + A].cpu().numpy(), "h": dense[jd * A + A:].cpu().numpy()}
    for name, got in cur.items():
        want = ref[name].numpy()
        bad = ~np.isclose(got, want, rtol=rtol, atol=atol)
        assert bad.sum() <= max(1, 3e-3 * bad.size), (name, int(bad.sum()), float(np.abs(got - want).max()))
    # scoring (NAIS_single.py:92-97)
    hist, tg = rs.choice(I, 12, replace=False), np.arange(I)
    sc = eng.score_nais(P.w, Q.w, B.w.reshape(-1)[:n].contiguous(), dense, A, hist, tg, 0.5, concat=concat).cpu().numpy()
    p64 = {k: torch.tensor(v).double() for k, v in cur.items()}
    q = p64["Q"][torch.tensor(tg)]
    s = T.nais_user_embed(p64, torch.tensor(hist), q, hp)
    want = ((s * q).sum(1) + p64["bias"][torch.tensor(tg)]).numpy()
    np.testing.assert_allclose(sc, want, rtol=1e-4, atol=2e-6)


def test_nais_sampler_matches_twin(eng):
    d = synthetic_data(20, 90, 7, seed=5)
    pu, pi, rp, sc = X.build_history(d.ui_train, d.user_nums)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    R, first = 4, 0
    for u, items in list(d.ui_train.items())[:5]:
        n = len(items)
        tg, y = eng.sample_nais(3, 1, first, n, R)
        negs = X.group_negatives(np.arange(first, first + n), np.full(n, u), 3, 1, R, d.item_nums, rp, sc)
        want = np.concatenate([np.concatenate([[items[k]], negs[k]]) for k in range(n)])
        assert np.array_equal(tg.cpu().numpy(), want)
        assert np.array_equal(y.cpu().numpy(), np.tile(np.asarray([1.0] + [0.0] * R, dtype=np.float32), n))
        first += n


def test_nais_model_evaluation_equals_the_reference_loops():
    """test_model_loo_nais / test_model_rs_nais (scores written into one device buffer from device-resident histories and candidates,
    ranking on the host / device) against the reference's loops (RankingRecommender.py:301-348, restated in oracle/ref_host.py) fed
    with per-user scores obtained the plain way (host lists in, host scores out)."""
    import logging
    from conftest import synthetic_data
    from oracle import ref_host as H
    from cleverrec_b200.model.ranking.NAIS_single import NAIS_single
    cfg = {'model_type': 'ranking', 'saved_dir': './saved_model', 'data.split_way': 'loo', 'test.neg_samples': '49', 'test.batch_size': '64',
           'test.interval': '1', 'topk': '[5,10]', 'epoches': '1', 'batch_size': '512', 'lr': '0.01', 'neg_ratio': '2', 'optimizer': 'Adagrad',
           'init_method': 'xavier_uniform', 'stddev': '0.05', 'seed': '3', 'recommender': 'NAIS_single', 'embed_size': '32', 'atten_size': '16',
           'atten_type': "'prod'", 'beta': '0.5', 'reg': '1e-3', 'nais_like': 'True', 'is_pairwise': 'False', 'loss_func': 'cross_entropy'}
    data = synthetic_data(120, 300, 12, seed=11, test_per_user=1)
    rs = np.random.RandomState(0)
    for u in data.ui_test:
        cand = np.setdiff1d(np.arange(data.item_nums), data.ui_train[u])
        data.ui_test[u] = rs.choice(cand, 49, replace=False).tolist() + data.ui_test[u]
    data.ui_test[7] = rs.choice(300, 49, replace=False).tolist() + [5]        # a test user without training history (u % 11 == 7)
    mc = NAIS_single(None, data, dict(cfg, atten_type='concat'), logging.getLogger('test'))     # the other attention type, end to end
    mc.build_model()
    assert mc.dense.numel() == 2 * 32 * 16 + 2 * 16 and mc._variables()['NAIS_params/W'].shape == (64, 16)
    lc = [mc.train_model() for _ in range(3)]
    assert np.all(np.isfinite(lc)) and lc[-1] < lc[0] and len(mc.test_model_loo()[0][0]) == len(mc.test_users)
    m = NAIS_single(None, data, cfg, logging.getLogger('test'))
    m.build_model()
    m.train_model()
    HR, MRR, NDCG = m.test_model_loo()
    scores = {u: m._scores(u, data.ui_test[u]) for u in m.test_users}
    oHR, oMRR, oNDCG = H.eval_loo(m.test_users, data.ui_test, scores, 49, m.topk)
    for k in range(len(m.topk)):
        assert HR[k] == oHR[k] and MRR[k] == oMRR[k] and NDCG[k] == oNDCG[k]
    data2 = synthetic_data(120, 300, 12, seed=11, test_per_user=2)
    m2 = NAIS_single(None, data2, dict(cfg, **{'data.split_way': 'rs', 'test.neg_samples': '0'}), logging.getLogger('test'))
    m2.build_model()
    m2.train_model()
    HR, MRR, NDCG = m2.test_model_rs()
    rows = np.stack([m2._scores(u, np.arange(data2.item_nums)) for u in m2.test_users])
    oHR, oMRR, oNDCG = H.eval_rs(m2.test_users, data2.ui_train, data2.ui_test, rows, m2.topk)
    for k in range(len(m2.topk)):
        assert HR[k] == oHR[k] and NDCG[k] == oNDCG[k]


def test_nais_numpy_stream_feeds_are_the_reference_loop_bit_for_bit():
    """sampler=numpy_stream for NAIS_single: the per-user feeds (history, targets = every positive followed by its neg_ratio negatives,
    labels) equal those the reference's train_model_nais loop (RankingRecommender.py:64-87) builds under the same np.random.seed --
    re-executed here literally -- and NumPy's global stream ends where the reference leaves it."""
    import logging
    from conftest import synthetic_data
    from cleverrec_b200.model.ranking.NAIS_single import NAIS_single
    cfg = {'model_type': 'ranking', 'saved_dir': './saved_model', 'data.split_way': 'loo', 'test.neg_samples': '49', 'test.batch_size': '64',
           'test.interval': '1', 'topk': '[5,10]', 'epoches': '1', 'batch_size': '512', 'lr': '0.01', 'neg_ratio': '3', 'optimizer': 'Adagrad',
           'init_method': 'xavier_uniform', 'stddev': '0.05', 'seed': '3', 'recommender': 'NAIS_single', 'embed_size': '32', 'atten_size': '16',
           'atten_type': "'prod'", 'beta': '0.5', 'reg': '1e-3', 'nais_like': 'True', 'is_pairwise': 'False', 'loss_func': 'cross_entropy',
           'sampler': 'numpy_stream'}
    data = synthetic_data(40, 90, 8, seed=2)
    m = NAIS_single(None, data, cfg, logging.getLogger('test'))
    m.build_model()
    fed = []
    real_step = m.train_step

    def spy(u_idx, i_idx, y, loss_out=None):
        fed.append((u_idx.cpu().numpy().tolist(), i_idx.cpu().numpy().tolist(), y.cpu().numpy().tolist()))
        return real_step(u_idx, i_idx, y, loss_out=loss_out)
    m.train_step = spy
    np.random.seed(21)
    loss = m.train_model()
    after = np.random.get_state()
    assert np.isfinite(loss)
    # the reference loop, literally
    np.random.seed(21)
    want = []
    for u, items in data.ui_train.items():
        i_idx, y = [], []
        seen_items = set(data.ui_train[u])
        for i in items:
            i_idx.append(i); y.append(1.0)
            random_j = set()
            for s in range(3):
                j = np.random.randint(data.item_nums)
                while j in random_j or j in seen_items:
                    j = np.random.randint(data.item_nums)
                random_j.add(j)
                i_idx.append(j); y.append(0.0)
        want.append((list(items), i_idx, y))
    assert fed == want
    ref_state = np.random.get_state()
    assert after[2] == ref_state[2] and np.array_equal(after[1], ref_state[1])
