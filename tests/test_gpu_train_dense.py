"""-m gpu: CML and FISM steps (csrc/train_dense.cu) against the torch restatement of model/ranking/CML.py:39-70 and
model/ranking/FISM.py:40-63 (dense gradients -> TF dense optimizer apply on every row)."""
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


def _close(got, want, kind, lr, name):
    rtol, atol = (2e-4, 2e-5) if kind == "Adam" else (2e-5, 1e-6)
    bad = ~np.isclose(got, want, rtol=rtol, atol=atol)
    assert bad.sum() <= max(1, 2e-3 * bad.size), (name, int(bad.sum()), float(np.abs(got - want).max()))  # Adam near-cancellation outliers
    assert np.abs(got - want).max() <= 0.05 * lr + 1e-6, name


@pytest.mark.parametrize("kind", ["SGD", "Adagrad", "Adam"])
@pytest.mark.parametrize("d", [32, 128])
def test_cml_steps(eng, kind, d):
    from cleverrec_b200.engine import Optimizer, Table
    U, I, R = 40, 90, 20
    g = torch.Generator().manual_seed(d)
    P0, Q0 = torch.randn(U, d, generator=g) * 0.2, torch.randn(I, d, generator=g) * 0.2
    lr = 0.01 if kind != "Adam" else 0.003
    opt, ropt = Optimizer(kind, lr, adam_mode="lazy"), T.TF1Optimizer(kind, lr)
    P, Q = Table(P0.cuda(), kind, "lazy"), Table(Q0.cuda(), kind, "lazy")
    ref = {"P": P0.clone(), "Q": Q0.clone()}
    hp = {"reg": 10.0, "margin": 1.0, "item_nums": I, "neg_ratio": R}
    rs = np.random.RandomState(d)
    for B in (64, 1, 130):
        u, i = rs.randint(0, U, B), rs.randint(0, I, B)
        neg = rs.randint(0, I, (B, R))
        got = eng.train_step_cml(P, Q, opt, u, i, neg, 1.0, 10.0, I)
        b = {"u": torch.tensor(u), "i": torch.tensor(i), "neg": torch.tensor(neg)}
        want = T.train_step(T.cml_loss, ref, b, hp, ropt)  # both tables dense
        assert abs(got - want) <= 5e-5 * abs(want) + 1e-5, (got, want)
    _close(P.w.cpu().numpy(), ref["P"].numpy(), kind, lr, "P")
    _close(Q.w.cpu().numpy(), ref["Q"].numpy(), kind, lr, "Q")
    assert float(P.grad.abs().max()) == 0.0 and float(Q.grad.abs().max()) == 0.0  # buffers left zeroed


@pytest.mark.parametrize("kind", ["SGD", "Adagrad", "Adam"])
def test_fism_steps(eng, kind):
    from cleverrec_b200.engine import Optimizer, Table
    d = synthetic_data(30, 80, 8, seed=3)
    d.ui_train[2] = d.ui_train[2] + d.ui_train[2][:2]  # duplicated interactions: list length != set size
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    dim, n = 32, d.item_nums + 1
    g = torch.Generator().manual_seed(1)
    P0, Q0, b0 = torch.randn(n, dim, generator=g) * 0.2, torch.randn(n, dim, generator=g) * 0.2, torch.rand(n, generator=g) * 0.2 - 0.1
    pad = (-n) % 4
    lr = 0.05 if kind != "Adam" else 0.01
    opt, ropt = Optimizer(kind, lr, adam_mode="lazy"), T.TF1Optimizer(kind, lr)
    P, Q = Table(P0.cuda(), kind, "lazy"), Table(Q0.cuda(), kind, "lazy")
    Bt = Table(torch.cat([b0, torch.zeros(pad)]).reshape(-1, 1).cuda().contiguous(), kind, "lazy")
    ref = {"P": P0.clone(), "Q": Q0.clone(), "b": b0.clone()}
    rows = torch.tensor([u for u, it in d.ui_train.items() for _ in it])
    cols = torch.tensor([i for u, it in d.ui_train.items() for i in it])
    vals = torch.tensor([1.0 / len(it) for u, it in d.ui_train.items() for _ in it])
    hp = {"reg": 1e-3, "reg_bias": 1e-3, "alpha": 0.4, "batch_size": 64, "user_nums": d.user_nums, "loss_func": "bpr"}
    users = list(d.ui_train.keys())
    rs = np.random.RandomState(0)
    for Bn in (64, 1, 100):
        u = np.asarray([users[k] for k in rs.randint(0, len(users), Bn)])
        i = np.asarray([d.ui_train[x][rs.randint(len(d.ui_train[x]))] for x in u])
        j = rs.randint(0, d.item_nums, Bn)
        nbr = np.asarray([len(set(d.ui_train[x])) for x in u])
        got = eng.train_step_fism(P, Q, Bt, opt, u, i, j, nbr, 0.4, 1e-3, 1e-3, 64)
        b = {"u": torch.tensor(u), "i": torch.tensor(i), "j": torch.tensor(j), "nbr_num": torch.tensor(nbr)}
        want = T.train_step(T.fism_loss, ref, b, hp, ropt, extra=((rows, cols, vals),))
        assert abs(got - want) <= 5e-5 * abs(want), (got, want)
    _close(P.w.cpu().numpy(), ref["P"].numpy(), kind, lr, "P")
    _close(Q.w.cpu().numpy(), ref["Q"].numpy(), kind, lr, "Q")
    _close(Bt.w.cpu().numpy().reshape(-1)[:n], ref["b"].numpy(), kind, lr, "b")
    # evaluation-time user vectors (FISM.py:70) with the list-length neighbour count
    tu = np.asarray(users[:10], dtype=np.int32)
    nb = np.asarray([len(d.ui_train[x]) for x in tu], dtype=np.int32)
    S = eng.fism_user_vectors(P.w, tu, nb, 0.4).cpu().numpy()
    refP = P.w.cpu().numpy()
    for k, x in enumerate(tu):
        want = (float(nb[k]) ** -0.4) * refP[d.ui_train[x]].mean(0)
        np.testing.assert_allclose(S[k], want, rtol=2e-5, atol=1e-7)


def test_clip_rows(eng):
    x = torch.randn(100, 48).cuda() * 0.3
    got = eng.clip_rows(x, 1.0).cpu().numpy()
    n = np.linalg.norm(x.cpu().numpy(), axis=1, keepdims=True)
    np.testing.assert_allclose(got, x.cpu().numpy() / np.maximum(n, 1.0), rtol=1e-6)
