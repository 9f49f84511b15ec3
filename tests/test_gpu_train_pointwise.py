"""-m gpu: pointwise MF / GMF steps (csrc/train.cu pointwise_step_kernel + dense_apply_kernel) against the torch
restatement of model/ranking/GMF.py:37-49 (and the MF specification) with TF-1 optimizer semantics."""
import numpy as np
import pytest
import torch

from oracle import tf1_restatement as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cleverrec_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("kind,mode", [("SGD", "tf1"), ("Adagrad", "tf1"), ("Adam", "tf1"), ("Adam", "lazy")])
@pytest.mark.parametrize("model,loss", [("GMF", "cross_entropy"), ("MF", "square"), ("MF", "cross_entropy")])
@pytest.mark.parametrize("d", [32, 64, 100])
def test_pointwise_steps_match_restatement(eng, kind, mode, model, loss, d):
    from cleverrec_b200 import _lib
    from cleverrec_b200.engine import Optimizer, Table
    U, I = 50, 70
    g = torch.Generator().manual_seed(d)
    P0, Q0, h0 = torch.randn(U, d, generator=g) * 0.3, torch.randn(I, d, generator=g) * 0.3, torch.randn(d, generator=g)
    lr = 0.05 if kind != "Adam" else 0.01
    opt, ropt = Optimizer(kind, lr, adam_mode=mode), T.TF1Optimizer(kind, lr, adam_mode=mode)
    P, Q = Table(P0.cuda(), kind, mode), Table(Q0.cuda(), kind, mode)
    ref = {"P": P0.clone(), "Q": Q0.clone()}
    gmf = model == "GMF"
    hd = s1 = s2 = None
    if gmf:
        ref["h"] = h0.clone()
        hd = h0.cuda()
        s1 = torch.full_like(hd, 0.1) if kind == "Adagrad" else (torch.zeros_like(hd) if kind == "Adam" else None)
        s2 = torch.zeros_like(hd) if kind == "Adam" else None
    loss_kind = {"cross_entropy": _lib.LOSS_CROSS_ENTROPY, "square": _lib.LOSS_SQUARE}[loss]
    rs = np.random.RandomState(d)
    for B in (96, 1, 300):
        u, i = rs.randint(0, U, B), rs.randint(0, I, B)
        y = (rs.rand(B) < 0.3).astype(np.float32)
        got = eng.train_step_pointwise(_lib.SCORE_GMF if gmf else _lib.SCORE_DOT, P, Q, opt, u, i, y, 0.01, loss_kind, hd, s1, s2)
        b = {"u": torch.tensor(u), "i": torch.tensor(i), "y": torch.tensor(y)}
        want = T.train_step(T.gmf_loss if gmf else T.mf_loss, ref, b, {"reg": 0.01, "loss_func": loss}, ropt, sparse_index={"P": ["u"], "Q": ["i"]})
        assert abs(got - want) <= 2e-5 * abs(want)
    eng.adam_flush(P, opt)
    eng.adam_flush(Q, opt)
    rtol, atol = (1e-4, 2e-5) if kind == "Adam" else (1e-5, 1e-6)
    for t_, name in ((P, "P"), (Q, "Q")):
        got, want = t_.w.cpu().numpy(), ref[name].numpy()
        bad = ~np.isclose(got, want, rtol=rtol, atol=atol)
        assert bad.mean() <= 2e-3, (name, bad.sum(), np.abs(got - want).max())
        assert np.abs(got - want).max() <= 0.05 * lr
    if gmf:
        np.testing.assert_allclose(hd.cpu().numpy(), ref["h"].numpy(), rtol=1e-4, atol=2e-6)
