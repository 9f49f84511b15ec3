"""-m gpu: SBPR (sampler csrc/sampler.cu::sample_sbpr_kernel, step csrc/train_dense.cu::sbpr_step_kernel, model/ranking/SBPR.py)
against the integer-exact sampler twin (oracle/philox.py::sample_sbpr), the torch restatement of the TF graph
(oracle/tf1_restatement.py::sbpr_loss) and the reference's evaluation loops."""
import logging
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, Data, synthetic_data, unflat
from oracle import c_oracle as O
from oracle import philox as X
from oracle import ref_host as H
from oracle import tf1_restatement as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cleverrec_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _ciao():
    z = np.load(os.path.join(GOLDEN, "sbpr_ciao.npz"))
    d = Data(int(z["user_nums"]), int(z["item_nums"]), unflat(z["train_keys"], z["train_lens"], z["train_items"]), {})
    d.user_friends = unflat(z["friends_keys"], z["friends_lens"], z["friends_items"])
    return d, unflat(z["spu_keys"], z["spu_lens"], z["spu_items"])


def _social_synthetic(n_users=120, n_items=300, seed=11, test_per_user=1):
    d = synthetic_data(n_users, n_items, 12, seed=seed, test_per_user=test_per_user)
    rs = np.random.RandomState(seed)
    d.user_friends = {}
    for u in range(n_users):
        if u % 5 == 3:
            continue                                   # users without friends: absent from SPu, never sampled
        f = rs.choice(n_users, size=rs.randint(1, 6), replace=False).tolist()
        d.user_friends[u] = f + f[:1] if u % 7 == 0 else f   # a repeated trust row counts twice in suk
    return d


@pytest.mark.parametrize("which", ["ciao", "synthetic"])
def test_sbpr_sampler_bit_exact(eng, which):
    if which == "ciao":
        d, SPu = _ciao()
    else:
        d = _social_synthetic()
        SPu = H.get_SPu(d)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    eng.set_social(d.ui_train, d.user_friends, SPu, d.user_nums)
    social = X.social_history(d.ui_train, d.user_friends, SPu, d.user_nums)
    R = 3
    n = eng.epoch_rows(R, "sbpr")
    assert n == social[0].shape[0] * R
    cnt = min(n, 6000)
    for seed, epoch, first in ((0, 0, 0), (0xFEEDFACECAFE, 3, n - cnt)):
        got = eng.sample_sbpr(seed, epoch, first, cnt, R)
        want = X.sample_sbpr(seed, epoch, first, cnt, R, d.item_nums, social)
        for g, w, name in zip(got, want, "uikjs"):
            assert np.array_equal(g.cpu().numpy(), w), name
    u, i, k, j = (t.cpu().numpy() for t in eng.sample_sbpr(1, 0, 0, cnt, R, is_suk=False))
    for t in range(0, cnt, 61):
        assert k[t] in SPu[u[t]] and j[t] not in d.ui_train[u[t]] and j[t] not in SPu[u[t]] and i[t] in d.ui_train[u[t]]
    # a whole epoch covers every social positive exactly neg_ratio times
    if n <= 200000:
        u, i, k, j = (t.cpu().numpy() for t in eng.sample_sbpr(2, 1, 0, n, R, is_suk=False))
        got_pairs = np.sort(u.astype(np.int64) * d.item_nums + i)
        want_pairs = np.sort(np.repeat(social[0].astype(np.int64) * d.item_nums + social[1], R))
        assert np.array_equal(got_pairs, want_pairs)


@pytest.mark.parametrize("kind", ["SGD", "Adagrad", "Adam"])
@pytest.mark.parametrize("d", [128, 32])
def test_sbpr_steps(eng, kind, d):
    from cleverrec_b200.engine import Optimizer, Table
    U, I = 40, 61
    g = torch.Generator().manual_seed(d)
    ref = {"P": torch.randn(U, d, generator=g) * 0.3, "Q": torch.randn(I, d, generator=g) * 0.3, "bias": torch.randn(I + 1, generator=g) * 0.1}
    lr = 0.02 if kind != "Adam" else 0.005
    opt, ropt = Optimizer(kind, lr, adam_mode="lazy"), T.TF1Optimizer(kind, lr, adam_mode="tf1")
    P, Q = Table(ref["P"].clone().cuda(), kind, "lazy"), Table(ref["Q"].clone().cuda(), kind, "lazy")
    n = I + 1
    B = Table(torch.cat([ref["bias"], torch.zeros((-n) % 4)]).reshape(-1, 1).cuda().contiguous(), kind, "lazy")
    rs = np.random.RandomState(1)
    for Bsz in (128, 1, 77):
        u, i, k, j = rs.randint(0, U, Bsz), rs.randint(0, I, Bsz), rs.randint(0, I, Bsz), rs.randint(0, I, Bsz)
        suk = rs.randint(1, 5, Bsz).astype(np.float32)
        got = eng.train_step_sbpr(P, Q, B, opt, u, i, k, j, suk, 0.05)
        b = {"u": torch.tensor(u), "i": torch.tensor(i), "k": torch.tensor(k), "j": torch.tensor(j), "suk": torch.tensor(suk)}
        want = T.train_step(T.sbpr_loss, ref, b, {"reg": 0.05}, ropt, sparse_index={"P": ["u"], "Q": ["i", "k", "j"], "bias": ["i", "k", "j"]})
        assert abs(got - want) <= 5e-5 * abs(want), (got, want)
    rtol, atol = (3e-4, 3e-5) if kind == "Adam" else (3e-5, 2e-6)
    for got, want, name in ((P.w, ref["P"], "P"), (Q.w, ref["Q"], "Q"), (B.w.reshape(-1)[:n], ref["bias"], "bias")):
        bad = ~np.isclose(got.cpu().numpy(), want.numpy(), rtol=rtol, atol=atol)
        assert bad.sum() <= max(1, 3e-3 * bad.size), (name, int(bad.sum()))
    assert float(B.w.reshape(-1)[n:].abs().max() if (-n) % 4 else 0.0) == 0.0          # the padding of the bias vector never moves
    assert float(P.grad.abs().max()) == 0.0 and float(Q.grad.abs().max()) == 0.0 and float(B.grad.abs().max()) == 0.0


CFG = {'model_type': 'ranking', 'saved_dir': './saved_model', 'data.split_way': 'loo', 'test.neg_samples': '49', 'test.batch_size': '64',
       'test.interval': '1', 'topk': '[5,10]', 'epoches': '2', 'batch_size': '512', 'lr': '0.01', 'neg_ratio': '3', 'optimizer': 'Adam',
       'init_method': 'normal', 'stddev': '0.05', 'seed': '3', 'recommender': 'SBPR', 'embed_size': '32', 'reg': '0.05',
       'is_pairwise': 'True', 'loss_func': 'bpr', 'social_file': 'trusts.csv'}


def test_sbpr_model_trains_and_evaluates_like_the_reference_loops():
    from cleverrec_b200.model.ranking.SBPR import SBPR
    data = _social_synthetic()
    rs = np.random.RandomState(0)
    for u in data.ui_test:
        cand = np.setdiff1d(np.arange(data.item_nums), data.ui_train[u])
        data.ui_test[u] = rs.choice(cand, 49, replace=False).tolist() + data.ui_test[u]
    m = SBPR(None, data, dict(CFG), logging.getLogger('test'))
    m.build_model()
    assert m.SPu == H.get_SPu(data)
    losses = [m.train_model() for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    HR, MRR, NDCG = m.test_model_loo()
    P, Q, b = m.P.w.cpu().numpy(), m.Q.w.cpu().numpy(), m.bias.cpu().numpy()
    scores = {u: O.score_pairs(3, P, Q, np.full(len(data.ui_test[u]), u), np.asarray(data.ui_test[u]), b) for u in m.test_users}
    oHR, oMRR, oNDCG = H.eval_loo(m.test_users, data.ui_test, scores, 49, m.topk)
    for k in range(len(m.topk)):
        assert HR[k] == oHR[k] and MRR[k] == oMRR[k] and NDCG[k] == oNDCG[k]
    # full ranking: SBPR.py:63 scores with the plain matmul (no bias)
    data2 = _social_synthetic(test_per_user=2)
    m2 = SBPR(None, data2, dict(CFG, **{'data.split_way': 'rs', 'test.neg_samples': '0', 'score_exact': 'True'}), logging.getLogger('test'))
    m2.build_model()
    m2.train_model()
    HR, MRR, NDCG = m2.test_model_rs()
    P, Q = m2.P.w.cpu().numpy(), m2.Q.w.cpu().numpy()
    I = data2.item_nums
    users = np.asarray(m2.test_users)
    sc = O.score_pairs(0, P, Q, np.repeat(users, I), np.tile(np.arange(I), len(users)), None).reshape(len(users), I)
    oHR, oMRR, oNDCG = H.eval_rs(m2.test_users, data2.ui_train, data2.ui_test, sc, m2.topk)
    for k in range(len(m2.topk)):
        assert HR[k] == oHR[k] and NDCG[k] == oNDCG[k]


def test_sampler_api_tuple(eng):
    """The reference's free function: (train_batches, u, i, i_s, i_neg, suk) as int64 NumPy arrays (utils/sampler.py:133-141)."""
    from cleverrec_b200.utils import sampler as S
    d = _social_synthetic()
    SPu = H.get_SPu(d)
    S.set_mode("philox")
    out = S.ranking_sampler_sbpr(d, SPu, 2, 100)
    n = 2 * sum(len(d.ui_train[u]) for u in d.ui_train if u in SPu)
    assert out[0] == -(-n // 100) and len(out) == 6 and all(a.shape == (n,) and a.dtype == np.int64 for a in out[1:])
    assert len(S.ranking_sampler_sbpr(d, SPu, 2, 100, is_suk=False)) == 5


def test_sbpr_without_social_data_is_an_error():
    from cleverrec_b200.model.ranking.SBPR import SBPR
    d = synthetic_data(30, 50, 5, seed=1)
    d.user_friends = {}
    with pytest.raises(ValueError):
        SBPR(None, d, dict(CFG), logging.getLogger('test'))


def test_sbpr_numpy_stream_is_the_reference_sampler_bit_for_bit(eng):
    """sampler=numpy_stream for SBPR: the device replays np.random's MT19937 stream through ranking_sampler_sbpr's draws (the
    social-item randint -- which consumes nothing for a one-item list --, the rejection-sampled negative, the final permutation).
    Golden: the GENUINE reference function on Ciao under np.random.seed(3) (oracle/make_golden.py section F), and the stream
    position afterwards equals NumPy's own."""
    d, SPu = _ciao()
    z = np.load(os.path.join(GOLDEN, "sbpr_ciao.npz"))
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    eng.set_social(d.ui_train, d.user_friends, SPu, d.user_nums)
    eng.np_seed(3)
    u, i, k, j, suk = (t.cpu().numpy() for t in eng.sample_epoch_numpy_sbpr(2))
    for got, key in ((u, "sb_u"), (i, "sb_i"), (k, "sb_k"), (j, "sb_j"), (suk, "sb_suk")):
        assert np.array_equal(got.astype(np.int64), z[key].astype(np.int64)), key
    # the stream left behind is NumPy's after the same calls
    np.random.seed(3)
    H.ranking_sampler_sbpr(d, SPu, 2, 4096)
    want = np.random.get_state()
    got = eng.np_get_state()
    assert got[2] == want[2] and np.array_equal(got[1], want[1])
    # a hand-made case with ONE-item social lists (np.random.randint(1) consumes no random value) next to longer ones
    d2 = Data(6, 14, {0: [0, 1, 2], 1: [0, 1, 2, 3], 2: [4, 5], 3: [4, 5, 6, 7, 8], 4: [9], 5: [0, 9, 10, 11, 12, 13]}, {})
    d2.user_friends = {0: [1], 1: [0, 3], 2: [3, 5], 3: [2], 4: [5, 0], 5: [4]}
    SPu2 = H.get_SPu(d2)
    assert sorted(len(v) for v in SPu2.values())[0] == 1 and max(len(v) for v in SPu2.values()) > 2
    eng.set_history(d2.ui_train, d2.user_nums, d2.item_nums)
    eng.set_social(d2.ui_train, d2.user_friends, SPu2, d2.user_nums)
    np.random.seed(11)
    want = H.ranking_sampler_sbpr(d2, SPu2, 3, 512)
    eng.np_seed(11)
    got = eng.sample_epoch_numpy_sbpr(3)
    for g, w in zip(got, want[1:]):
        assert np.array_equal(g.cpu().numpy().astype(np.int64), np.asarray(w).astype(np.int64))
