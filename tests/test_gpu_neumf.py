"""-m gpu: NeuMF step and scoring (csrc/train_neumf.cu) against the torch restatement of model/ranking/NeuMF.py:58-105."""
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


def _pack(params, layers, E):
    parts = []
    for k in range(len(layers)):
        parts += [params["W_%d" % k].reshape(-1), params["b_%d" % k].reshape(-1)]
    parts.append(params["h_neumf"].reshape(-1))
    return torch.cat(parts)


@pytest.mark.parametrize("kind", ["SGD", "Adagrad", "Adam"])
@pytest.mark.parametrize("layers,E", [([128, 64, 32], 32), ([16, 8], 8)])
def test_neumf_steps(eng, kind, layers, E):
    from cleverrec_b200 import _lib
    from cleverrec_b200.engine import Optimizer, Table
    U, I = 40, 60
    g = torch.Generator().manual_seed(len(layers))
    rnd = lambda *s: torch.randn(*s, generator=g) * 0.2
    ref = {"P_gmf": rnd(U, E), "Q_gmf": rnd(I, E), "P_mlp": rnd(U, layers[0] // 2), "Q_mlp": rnd(I, layers[0] // 2)}
    for k, n in enumerate(layers):
        ref["W_%d" % k], ref["b_%d" % k] = rnd(n, n // 2), rnd(n // 2)
    ref["h_neumf"] = rnd(E + layers[-1] // 2)
    lr = 0.02 if kind != "Adam" else 0.005
    opt, ropt = Optimizer(kind, lr, adam_mode="lazy"), T.TF1Optimizer(kind, lr, adam_mode="tf1")
    tabs = [Table(ref[n].clone().cuda(), kind, "lazy") for n in ("P_gmf", "Q_gmf", "P_mlp", "Q_mlp")]
    dense = _pack(ref, layers, E).cuda()
    s1 = torch.full_like(dense, 0.1) if kind == "Adagrad" else (torch.zeros_like(dense) if kind == "Adam" else None)
    s2 = torch.zeros_like(dense) if kind == "Adam" else None
    hp = {"reg1": 1e-2, "reg2": 1e-3, "n_layers": len(layers), "loss_func": "cross_entropy"}
    rs = np.random.RandomState(1)
    sparse = {"P_gmf": ["u"], "Q_gmf": ["i"], "P_mlp": ["u"], "Q_mlp": ["i"]}
    for B in (128, 1, 77):
        u, i = rs.randint(0, U, B), rs.randint(0, I, B)
        y = (rs.rand(B) < 0.25).astype(np.float32)
        got = eng.train_step_neumf(tabs, dense, s1, s2, len(layers), opt, u, i, y, 1e-2, 1e-3, _lib.LOSS_CROSS_ENTROPY)
        b = {"u": torch.tensor(u), "i": torch.tensor(i), "y": torch.tensor(y)}
        want = T.train_step(T.neumf_loss, ref, b, hp, ropt, sparse_index=sparse)
        assert abs(got - want) <= 5e-5 * abs(want), (got, want)
    rtol, atol = (3e-4, 3e-5) if kind == "Adam" else (3e-5, 2e-6)
    for t_, name in zip(tabs, ("P_gmf", "Q_gmf", "P_mlp", "Q_mlp")):
        got, want = t_.w.cpu().numpy(), ref[name].numpy()
        bad = ~np.isclose(got, want, rtol=rtol, atol=atol)
        assert bad.sum() <= max(1, 3e-3 * bad.size), (name, int(bad.sum()), float(np.abs(got - want).max()))
    got, want = dense.cpu().numpy(), _pack(ref, layers, E).numpy()
    bad = ~np.isclose(got, want, rtol=rtol, atol=atol)
    assert bad.sum() <= max(1, 3e-3 * bad.size), ("dense", int(bad.sum()), float(np.abs(got - want).max()))
    # scoring: canonical logits agree with the restatement's logits to fp32 rounding
    u, i = rs.randint(0, U, 500), rs.randint(0, I, 500)
    sc = eng.score_pairs_neumf(tabs, dense, len(layers), u, i).cpu().numpy()
    cur = {n: t_.w.cpu() for t_, n in zip(tabs, ("P_gmf", "Q_gmf", "P_mlp", "Q_mlp"))}
    cur.update({k: v for k, v in ref.items() if k not in cur})
    lg, _ = T.neumf_logits({k: v.double() for k, v in cur.items()}, torch.tensor(u), torch.tensor(i), len(layers))
    np.testing.assert_allclose(sc, lg.numpy(), rtol=2e-4, atol=2e-5)


def test_mlp_model_step_matches_restated_tower(eng):
    """model/ranking/MLP.py (the tower alone, E = 0 in the fused kernels) against the restated NeuMF graph with the GMF branch
    removed: same loss and same tables after 3 steps."""
    import torch
    from oracle import tf1_restatement as T
    from cleverrec_b200.engine import Optimizer, Table
    U, I, Em, layers = 30, 50, 8, 2
    g = torch.Generator().manual_seed(4)
    p = {"P": torch.randn(U, Em, generator=g) * 0.3, "Q": torch.randn(I, Em, generator=g) * 0.3,
         "W_0": torch.randn(16, 8, generator=g) * 0.3, "b_0": torch.randn(8, generator=g) * 0.1,
         "W_1": torch.randn(8, 4, generator=g) * 0.3, "b_1": torch.randn(4, generator=g) * 0.1, "h_mlp": torch.randn(4, generator=g) * 0.3}

    def mlp_loss(q, b, hp):   # MLP.py:44-70
        x = torch.cat([q["P"][b["u"]], q["Q"][b["i"]]], 1)
        y_ = T.mlp_tower(x, q, 2)
        logits = y_ @ q["h_mlp"]
        return T.get_loss("cross_entropy", b["y"], logits=logits) + hp["reg"] * (T.l2_loss(q["P"][b["u"]]) + T.l2_loss(q["Q"][b["i"]]))
    for kind in ("SGD", "Adagrad"):
        ref = {k: v.clone() for k, v in p.items()}
        ropt = T.TF1Optimizer(kind, 0.05)
        P, Q = Table(p["P"].clone().cuda(), kind, "lazy"), Table(p["Q"].clone().cuda(), kind, "lazy")
        dense = torch.cat([p["W_0"].reshape(-1), p["b_0"], p["W_1"].reshape(-1), p["b_1"], p["h_mlp"]]).cuda()
        s1 = torch.full_like(dense, 0.1) if kind == "Adagrad" else None
        opt = Optimizer(kind, 0.05)
        rs = np.random.RandomState(0)
        for step in range(3):
            u, i = rs.randint(0, U, 64), rs.randint(0, I, 64)
            y = (rs.rand(64) < 0.3).astype(np.float32)
            got = eng.train_step_neumf([None, None, P, Q], dense, s1, None, layers, opt, u, i, y, 0.0, 0.01, 1)
            b = {"u": torch.tensor(u.astype(np.int64)), "i": torch.tensor(i.astype(np.int64)), "y": torch.tensor(y)}
            want = T.train_step(mlp_loss, ref, b, {"reg": 0.01}, ropt)
            assert abs(got - want) <= 1e-5 * abs(want)
        np.testing.assert_allclose(P.w.cpu().numpy(), ref["P"].numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(Q.w.cpu().numpy(), ref["Q"].numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(dense[:128].cpu().numpy().reshape(16, 8), ref["W_0"].numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(dense[-4:].cpu().numpy(), ref["h_mlp"].numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("layers,E", [([128, 64, 32], 32), ([16, 8], 8), ([64, 32, 16], 0)])
def test_neumf_warp_scorer_is_bit_identical_to_the_thread_scorer(eng, layers, E, monkeypatch):
    """neumf_score_warp_kernel (one warp per pair, layer outputs by lanes) and neumf_score_kernel (one thread per pair) run the same
    sequential chains: identical bits, including the MLP model (E = 0)."""
    from cleverrec_b200.engine import Table
    U, I = 70, 90
    g = torch.Generator().manual_seed(E + len(layers))
    rnd = lambda *s: torch.randn(*s, generator=g) * 0.3
    Em = layers[0] // 2
    tabs = [Table(rnd(U, E).cuda(), "SGD") if E else None, Table(rnd(I, E).cuda(), "SGD") if E else None, Table(rnd(U, Em).cuda(), "SGD"),
            Table(rnd(I, Em).cuda(), "SGD")]
    parts = []
    for n in layers:
        parts += [rnd(n, n // 2).reshape(-1), rnd(n // 2)]
    parts.append(rnd(E + layers[-1] // 2))
    dense = torch.cat(parts).cuda()
    rs = np.random.RandomState(0)
    u, i = rs.randint(0, U, 4001), rs.randint(0, I, 4001)
    monkeypatch.delenv("CRB_NEUMF_SCORE_SIMPLE", raising=False)
    fast = eng.score_pairs_neumf(tabs, dense, len(layers), u, i).cpu().numpy()
    monkeypatch.setenv("CRB_NEUMF_SCORE_SIMPLE", "1")
    slow = eng.score_pairs_neumf(tabs, dense, len(layers), u, i).cpu().numpy()
    assert np.array_equal(fast.view(np.uint32), slow.view(np.uint32))
    # ... and both are the C oracle's canonical logit, bit for bit (oracle/crb_oracle.c::oracle_score_pairs_neumf)
    from oracle import c_oracle as O
    want = O.score_pairs_neumf(tabs[0].w.cpu().numpy() if E else None, tabs[1].w.cpu().numpy() if E else None, tabs[2].w.cpu().numpy(),
                               tabs[3].w.cpu().numpy(), dense.cpu().numpy(), len(layers), u, i)
    assert np.array_equal(fast.view(np.uint32), want.view(np.uint32))
