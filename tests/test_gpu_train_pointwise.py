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


@pytest.mark.parametrize("kind_name,opt_kind", [("GMF", "Adam"), ("MF", "SGD"), ("GMF", "Adagrad")])
def test_pointwise_epoch_call_equals_sample_then_step(kind_name, opt_kind):
    """crb_train_epoch_pointwise (sampler fused in, one call per epoch) == crb_sample_pointwise + crb_train_step_pointwise per batch:
    same losses and tables bit for bit, ragged last batch included."""
    import torch
    from conftest import synthetic_data
    from cleverrec_b200 import _lib
    from cleverrec_b200.engine import Engine, Optimizer, Table
    eng = Engine(0)
    d = synthetic_data(90, 400, 15, seed=9)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    kind = _lib.SCORE_GMF if kind_name == "GMF" else _lib.SCORE_DOT
    dim, B, R = 32, 500, 3
    n_rows = eng.epoch_rows(R, "pointwise")
    n_steps = -(-n_rows // B)
    g = torch.Generator().manual_seed(1)
    P0, Q0, h0 = torch.randn(d.user_nums, dim, generator=g) * 0.1, torch.randn(d.item_nums, dim, generator=g) * 0.1, torch.randn(dim, generator=g)

    def fresh():
        P, Q, opt = Table(P0.clone().cuda(), opt_kind), Table(Q0.clone().cuda(), opt_kind), Optimizer(opt_kind, 0.01)
        h = h0.clone().cuda() if kind_name == "GMF" else None
        s1 = (torch.full_like(h, 0.1) if opt_kind == "Adagrad" else torch.zeros_like(h)) if (h is not None and opt_kind != "SGD") else None
        s2 = torch.zeros_like(h) if (h is not None and opt_kind == "Adam") else None
        return P, Q, opt, h, s1, s2
    Pa, Qa, oa, ha, s1a, s2a = fresh()
    la = torch.zeros(n_steps, dtype=torch.float64, device="cuda")
    for k in range(n_steps):
        cnt = min(B, n_rows - k * B)
        u, i, y = eng.sample_pointwise(5, 2, k * B, cnt, R)
        eng.train_step_pointwise(kind, Pa, Qa, oa, u, i, y, 0.01, _lib.LOSS_CROSS_ENTROPY, ha, s1a, s2a, loss_out=la[k:k + 1])
    Pb, Qb, ob, hb, s1b, s2b = fresh()
    lb = torch.zeros(n_steps, dtype=torch.float64, device="cuda")
    eng.train_epoch_pointwise(kind, Pb, Qb, ob, 5, 2, 0, B, n_steps, R, 0.01, _lib.LOSS_CROSS_ENTROPY, hb, s1b, s2b, loss_out=lb)
    assert torch.equal(la, lb) and oa.t == ob.t == n_steps
    for T_ in ((Pa, Pb), (Qa, Qb)):
        eng.adam_flush(T_[0], oa); eng.adam_flush(T_[1], ob)
        assert torch.equal(T_[0].w, T_[1].w)
    if ha is not None:
        assert torch.equal(ha, hb)
    host = np.zeros(2)
    eng.train_epoch_pointwise(kind, Pb, Qb, ob, 5, 3, 0, B, 2, R, 0.01, _lib.LOSS_CROSS_ENTROPY, hb, s1b, s2b, loss_out=host)   # host losses: synchronous
    assert np.all(host > 0)
    eng.close()
