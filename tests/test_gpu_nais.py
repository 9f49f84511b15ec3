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
@pytest.mark.parametrize("d,A", [(32, 16), (128, 32)])
def test_nais_user_steps(eng, kind, d, A):
    from cleverrec_b200.engine import Optimizer, Table
    I = 60
    n = I + 1
    g = torch.Generator().manual_seed(d)
    rnd = lambda *s: torch.randn(*s, generator=g) * 0.2
    ref = {"P": rnd(n, d), "Q": rnd(n, d), "bias": rnd(n) * 0.5, "W": rnd(d, A), "b_att": rnd(A) * 0.5, "h": rnd(A)}
    lr = 0.02 if kind != "Adam" else 0.005
    opt, ropt = Optimizer(kind, lr, adam_mode="lazy"), T.TF1Optimizer(kind, lr, adam_mode="tf1")
    P, Q = Table(ref["P"].clone().cuda(), kind, "lazy"), Table(ref["Q"].clone().cuda(), kind, "lazy")
    B = Table(torch.cat([ref["bias"], torch.zeros((-n) % 4)]).reshape(-1, 1).cuda().contiguous(), kind, "lazy")
    dense = torch.cat([ref["W"].reshape(-1), ref["b_att"], ref["h"]]).cuda()
    s1 = torch.full_like(dense, 0.1) if kind == "Adagrad" else (torch.zeros_like(dense) if kind == "Adam" else None)
    s2 = torch.zeros_like(dense) if kind == "Adam" else None
    rs = np.random.RandomState(A)
    hp = {"reg": 1e-3, "beta": 0.5}
    for n_hist in (5, 1, 40):
        hist = rs.choice(I, n_hist, replace=False)
        tg = rs.randint(0, I, n_hist * 3)
        y = (rs.rand(n_hist * 3) < 0.3).astype(np.float32)
        got = eng.train_step_nais(P, Q, B, dense, s1, s2, A, opt, hist, tg, y, 0.5, 1e-3)
        b = {"hist": torch.tensor(hist), "i": torch.tensor(tg), "y": torch.tensor(y)}
        want = T.train_step(T.nais_loss, ref, b, hp, ropt, sparse_index={"P": ["hist"], "Q": ["i"], "bias": ["i"]})
        assert abs(got - want) <= 5e-5 * abs(want), (got, want)
    rtol, atol = (3e-4, 3e-5) if kind == "Adam" else (3e-5, 2e-6)
    cur = {"P": P.w.cpu().numpy(), "Q": Q.w.cpu().numpy(), "bias": B.w.cpu().numpy().reshape(-1)[:n], "W": dense[:d * A].cpu().numpy().reshape(d, A),
           "b_att": dense[d * A:d * A + A].cpu().numpy(), "h": dense[d * A + A:].cpu().numpy()}
    for name, got in cur.items():
        want = ref[name].numpy()
        bad = ~np.isclose(got, want, rtol=rtol, atol=atol)
        assert bad.sum() <= max(1, 3e-3 * bad.size), (name, int(bad.sum()), float(np.abs(got - want).max()))
    # scoring (NAIS_single.py:92-97)
    hist, tg = rs.choice(I, 12, replace=False), np.arange(I)
    sc = eng.score_nais(P.w, Q.w, B.w.reshape(-1)[:n].contiguous(), dense, A, hist, tg, 0.5).cpu().numpy()
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
