"""-m gpu: TransCF (csrc/train_dense.cu transcf_*; model/ranking/TransCF.py:38-85) against the torch restatement of the TF graph
(oracle/tf1_restatement.py::transcf_loss with the two SpMMs of utils/tools.py:100-113) and a NumPy restatement of the canonical
score chain.  Tolerances: loss 1e-5 relative, tables 1e-5 relative (SGD / Adagrad) and the Adam criterion of test_gpu_train_bpr."""
import logging

import numpy as np
import pytest
import torch

from conftest import synthetic_data
from oracle import tf1_restatement as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cleverrec_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def sp_mats(ui_train):
    """get_sp_mat (utils/tools.py:100-113): COO (rows, cols, values) of ui_sp_mat and iu_sp_mat."""
    ur, uc, uv, ir, ic = [], [], [], [], []
    cnt = {}
    for u, items in ui_train.items():
        for i in items:
            ur.append(u); uc.append(i); uv.append(1.0 / len(items))
            ir.append(i); ic.append(u)
            cnt[i] = cnt.get(i, 0) + 1
    iv = [1.0 / cnt[i] for i in ir]
    t = lambda a, dt: torch.tensor(np.asarray(a), dtype=dt)
    return (t(ur, torch.int64), t(uc, torch.int64), t(uv, torch.float32)), (t(ir, torch.int64), t(ic, torch.int64), t(iv, torch.float32))


def make(eng, d_data, dim, kind, seed=0):
    from cleverrec_b200.engine import Optimizer, Table
    eng.set_history(d_data.ui_train, d_data.user_nums, d_data.item_nums)
    eng.set_item_lists()
    g = torch.Generator().manual_seed(seed)
    P0, Q0 = torch.randn(d_data.user_nums, dim, generator=g) * 0.3, torch.randn(d_data.item_nums, dim, generator=g) * 0.3
    opt = Optimizer(kind, 0.05 if kind != "Adam" else 0.01, adam_mode="lazy")
    P, Q = Table(P0.cuda(), kind, "lazy"), Table(Q0.cuda(), kind, "lazy")
    return P, Q, opt, {"P": P0.clone(), "Q": Q0.clone()}, T.TF1Optimizer(kind, opt.lr, adam_mode="lazy")


@pytest.mark.parametrize("kind", ["SGD", "Adagrad", "Adam"])
@pytest.mark.parametrize("dim", [16, 64, 100])
def test_step_matches_restated_graph(eng, kind, dim):
    d = synthetic_data(40, 90, 8, seed=dim)
    d.ui_train[3] = d.ui_train[3] + d.ui_train[3][:2]   # a list with repeated items: both matrices count duplicates
    P, Q, opt, ref, ropt = make(eng, d, dim, kind)
    ui, iu = sp_mats(d.ui_train)
    hp = {"reg1": 0.1, "reg2": 0.01, "margin": 0.5, "user_nums": d.user_nums, "item_nums": d.item_nums}
    rs = np.random.RandomState(1)
    users = np.asarray(list(d.ui_train.keys()))
    for step in range(3):
        B = 64 if step != 1 else 7
        u = rs.choice(users, B)
        i = np.asarray([rs.choice(d.ui_train[x]) for x in u])
        j = rs.randint(0, d.item_nums, B)
        j[0] = 89 if 89 not in {it for v in d.ui_train.values() for it in v} else j[0]   # an item nobody interacted with: beta = 0
        got = eng.train_step_transcf(P, Q, opt, u, i, j, hp["margin"], hp["reg1"], hp["reg2"])
        b = {"u": torch.tensor(u.astype(np.int64)), "i": torch.tensor(i.astype(np.int64)), "j": torch.tensor(j.astype(np.int64))}
        want = T.train_step(T.transcf_loss, ref, b, hp, ropt, extra=(ui, iu))   # dense apply of both tables
        assert abs(got - want) <= 1e-5 * abs(want), (step, got, want)
    for name, tab in (("P", P), ("Q", Q)):
        g_, w_ = tab.w.cpu().numpy(), ref[name].numpy()
        if kind == "Adam":
            bad = ~np.isclose(g_, w_, rtol=1e-4, atol=1e-5)
            assert bad.mean() <= 1e-3 and np.abs(g_ - w_).max() <= 0.02 * opt.lr * 3
        else:
            np.testing.assert_allclose(g_, w_, rtol=1e-5, atol=2e-7)


def test_neighbourhoods_and_pair_scores_bit_exact(eng):
    d = synthetic_data(30, 70, 6, seed=5)
    dim = 24
    P, Q, opt, ref, _ = make(eng, d, dim, "SGD")
    A = eng.transcf_neighbourhood(0, Q.w, d.user_nums).cpu().numpy()
    B = eng.transcf_neighbourhood(1, P.w, d.item_nums).cpu().numpy()
    Pn, Qn = ref["P"].numpy(), ref["Q"].numpy()

    def seg_mean(table, members):   # the canonical order: acc = fma(x, 1/n, acc) over the list (fp32; NumPy has no fma, so the
        acc = np.zeros(dim, dtype=np.float64)   # product-sum is formed exactly in fp64 and rounded once, which IS fma)
        if not len(members):
            return acc.astype(np.float32)
        inv = np.float32(1.0) / np.float32(len(members))
        acc = np.zeros(dim, dtype=np.float32)
        for m in members:
            acc = (table[m].astype(np.float64) * np.float64(inv) + acc.astype(np.float64)).astype(np.float32)
        return acc
    users_of = {}
    for u, items in d.ui_train.items():
        for it in items:
            users_of.setdefault(it, []).append(u)
    for u in range(d.user_nums):
        assert np.array_equal(A[u], seg_mean(Qn, d.ui_train.get(u, []))), u
    for it in range(d.item_nums):
        assert np.array_equal(B[it], seg_mean(Pn, users_of.get(it, []))), it
    rs = np.random.RandomState(0)
    u, i = rs.randint(0, d.user_nums, 500), rs.randint(0, d.item_nums, 500)
    got = eng.score_pairs_transcf(P.w, Q.w, torch.tensor(A).cuda(), torch.tensor(B).cuda(), u, i).cpu().numpy()
    want = np.zeros(500, dtype=np.float32)
    for k in range(500):
        acc = np.float32(0)
        for c in range(dim):
            e = np.float32(np.float32(np.float64(A[u[k], c]) * np.float64(B[i[k], c]) + np.float64(Pn[u[k], c])) - Qn[i[k], c])
            acc = np.float32(np.float64(e) * np.float64(e) + np.float64(acc))
        want[k] = acc
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_state_errors(eng):
    from cleverrec_b200._lib import CrbError
    from cleverrec_b200.engine import Engine, Optimizer, Table
    e2 = Engine(0)
    d = synthetic_data(10, 60, 8, seed=1)
    e2.set_history(d.ui_train, d.user_nums, d.item_nums)   # item lists NOT set
    P, Q = Table(torch.zeros(10, 8).cuda(), "SGD"), Table(torch.zeros(60, 8).cuda(), "SGD")
    with pytest.raises(CrbError) as err:
        e2.train_step_transcf(P, Q, Optimizer("SGD", 0.1), [0], [1], [2], 0.5, 0.1, 0.01)
    assert err.value.code == -3
    e2.close()
