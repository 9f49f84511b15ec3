"""Self-consistency pins of the TF-1 restatement (oracle/tf1_restatement.py): the reference has no tests and TF is
absent, so the restatement is checked by (i) fp64 finite differences of every loss graph, (ii) hand-computed
optimizer examples with TF-1's documented defaults."""
import math

import numpy as np
import torch

from oracle import tf1_restatement as T


def _fd_check(loss_fn, params, batch, hp, extra=(), eps=1e-6, n_probe=12):
    leaves = {k: v.clone().double().requires_grad_(True) for k, v in params.items()}
    loss = loss_fn(leaves, batch, hp, *extra)
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    rs = np.random.RandomState(0)
    for (name, leaf), g in zip(leaves.items(), grads):
        if g is None:
            continue
        flat = leaf.detach().reshape(-1)
        nz = torch.nonzero(g.reshape(-1)).reshape(-1).numpy()
        if nz.size == 0:
            continue
        for idx in rs.choice(nz, size=min(n_probe, nz.size), replace=False):
            def f(delta):
                p2 = {k: v.detach().clone() for k, v in leaves.items()}
                p2[name].reshape(-1)[idx] += delta
                return float(loss_fn(p2, batch, hp, *extra))
            num = (f(eps) - f(-eps)) / (2 * eps)
            assert abs(num - float(g.reshape(-1)[idx])) <= 1e-5 * max(1.0, abs(num)), (name, idx)


def _tables(U, I, d, seed, scale=0.3):
    g = torch.Generator().manual_seed(seed)
    return {"P": torch.randn(U, d, generator=g, dtype=torch.float64) * scale, "Q": torch.randn(I, d, generator=g, dtype=torch.float64) * scale}


def test_bpr_mf_gmf_gradients():
    p = _tables(7, 9, 8, 0)
    b = {"u": torch.tensor([0, 1, 1, 6]), "i": torch.tensor([2, 3, 3, 8]), "j": torch.tensor([4, 3, 0, 1])}
    _fd_check(T.bpr_loss, p, b, {"reg": 0.01})
    b2 = {"u": b["u"], "i": b["i"], "y": torch.tensor([1., 0., 1., 0.], dtype=torch.float64)}
    _fd_check(T.mf_loss, p, b2, {"reg": 0.01, "loss_func": "square"})
    _fd_check(T.mf_loss, p, b2, {"reg": 0.01, "loss_func": "cross_entropy"})
    p["h"] = torch.randn(8, dtype=torch.float64)
    _fd_check(T.gmf_loss, p, b2, {"reg": 0.01, "loss_func": "cross_entropy"})


def test_cml_fism_nais_transcf_neumf_gradients():
    p = _tables(6, 12, 4, 1)
    b = {"u": torch.tensor([0, 2, 5]), "i": torch.tensor([1, 3, 7]), "neg": torch.tensor([[2, 4, 9], [0, 5, 6], [8, 10, 11]])}
    _fd_check(T.cml_loss, p, b, {"reg": 10.0, "margin": 1.0, "item_nums": 12, "neg_ratio": 3})
    # FISM
    I = 12
    pf = {"P": torch.randn(I + 1, 4, dtype=torch.float64) * .3, "Q": torch.randn(I + 1, 4, dtype=torch.float64) * .3,
          "b": torch.randn(I + 1, dtype=torch.float64) * .1}
    rows = torch.tensor([0, 0, 0, 1, 1, 2]); cols = torch.tensor([1, 2, 3, 4, 5, 6]); vals = torch.tensor([1 / 3., 1 / 3., 1 / 3., .5, .5, 1.], dtype=torch.float64)
    bf = {"u": torch.tensor([0, 1, 2]), "i": torch.tensor([1, 4, 6]), "j": torch.tensor([7, 8, 9]), "nbr_num": torch.tensor([3, 2, 1])}
    _fd_check(T.fism_loss, pf, bf, {"reg": 1e-3, "reg_bias": 1e-3, "alpha": 0.4, "batch_size": 64, "user_nums": 3, "loss_func": "bpr"}, extra=((rows, cols, vals),))
    # NAIS
    pn = dict(pf); pn["bias"] = pn.pop("b")
    pn.update({"W": torch.randn(4, 3, dtype=torch.float64), "b_att": torch.randn(3, dtype=torch.float64) * .1, "h": torch.randn(3, dtype=torch.float64)})
    bn = {"hist": torch.tensor([1, 2, 3]), "i": torch.tensor([1, 7, 2, 9]), "y": torch.tensor([1., 0., 1., 0.], dtype=torch.float64)}
    _fd_check(T.nais_loss, pn, bn, {"reg": 1e-3, "beta": 0.5})
    # TransCF
    ui = (torch.tensor([0, 0, 2, 5]), torch.tensor([1, 3, 7, 7]), torch.tensor([.5, .5, 1., 1.], dtype=torch.float64))
    iu = (torch.tensor([1, 3, 7, 7]), torch.tensor([0, 0, 2, 5]), torch.tensor([1., 1., .5, .5], dtype=torch.float64))
    bt = {"u": torch.tensor([0, 2, 5]), "i": torch.tensor([1, 7, 7]), "j": torch.tensor([4, 5, 6])}
    _fd_check(T.transcf_loss, p, bt, {"reg1": 0.1, "reg2": 0.01, "margin": 0.5, "user_nums": 6, "item_nums": 12}, extra=(ui, iu))
    # NeuMF
    g = torch.Generator().manual_seed(3)
    pm = {"P_gmf": torch.randn(6, 4, generator=g, dtype=torch.float64), "Q_gmf": torch.randn(12, 4, generator=g, dtype=torch.float64),
          "P_mlp": torch.randn(6, 4, generator=g, dtype=torch.float64), "Q_mlp": torch.randn(12, 4, generator=g, dtype=torch.float64),
          "W_0": torch.randn(8, 4, generator=g, dtype=torch.float64), "b_0": torch.randn(4, generator=g, dtype=torch.float64),
          "W_1": torch.randn(4, 2, generator=g, dtype=torch.float64), "b_1": torch.randn(2, generator=g, dtype=torch.float64),
          "h_neumf": torch.randn(6, generator=g, dtype=torch.float64)}
    bm = {"u": torch.tensor([0, 2, 5]), "i": torch.tensor([1, 7, 7]), "y": torch.tensor([1., 0., 1.], dtype=torch.float64)}
    _fd_check(T.neumf_loss, pm, bm, {"reg1": 1e-2, "reg2": 1e-3, "n_layers": 2, "loss_func": "cross_entropy"})


def test_lrml_sbpr_gradients():
    p = _tables(6, 12, 8, 5)
    g = torch.Generator().manual_seed(6)
    p["K"], p["M"] = torch.randn(8, 5, generator=g, dtype=torch.float64), torch.randn(5, 8, generator=g, dtype=torch.float64) * 0.5
    b = {"u": torch.tensor([0, 2, 2, 5]), "i": torch.tensor([1, 3, 3, 7]), "j": torch.tensor([4, 3, 0, 11])}
    _fd_check(T.lrml_loss, p, b, {"reg": 1e-3, "margin": 0.2})
    # a margin large enough that every hinge is active, and zero: both branches of the hinge
    _fd_check(T.lrml_loss, p, b, {"reg": 1e-3, "margin": 50.0})
    ps = _tables(6, 12, 8, 7)
    ps["bias"] = torch.randn(13, generator=g, dtype=torch.float64) * 0.1
    bs = {"u": torch.tensor([0, 2, 2, 5]), "i": torch.tensor([1, 3, 3, 7]), "k": torch.tensor([2, 5, 6, 7]), "j": torch.tensor([4, 9, 0, 11]),
          "suk": torch.tensor([1., 2., 3., 1.], dtype=torch.float64)}
    _fd_check(T.sbpr_loss, ps, bs, {"reg": 0.05})


def test_tf1_optimizer_hand_examples():
    # one variable, two rows, row 0 touched with g = 2, row 1 never touched
    def run(kind, mode="tf1", steps=2):
        var = {"w": torch.tensor([[1.0], [1.0]])}
        opt = T.TF1Optimizer(kind, 0.1, adam_mode=mode)
        for _ in range(steps):
            opt.step(var, {"w": torch.tensor([[2.0], [0.0]])}, {"w": torch.tensor([0])})
        return var["w"].reshape(-1).tolist(), opt
    w, _ = run("SGD")
    assert np.allclose(w, [1 - 0.1 * 2 * 2, 1.0])
    w, _ = run("Adagrad", steps=1)
    assert np.allclose(w, [1 - 0.1 * 2 / math.sqrt(0.1 + 4), 1.0])  # accumulator starts at 0.1, no epsilon
    w, opt = run("Adam", steps=1)
    lr1 = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    m, v = 0.1 * 2, 0.001 * 4
    assert np.allclose(w, [1 - lr1 * m / (math.sqrt(v) + 1e-8), 1.0], atol=1e-7)  # epsilon outside the bias correction
    # tf1 mode: an untouched row whose moments are non-zero keeps moving; lazy mode: it does not
    for mode, moves in (("tf1", True), ("lazy", False)):
        var = {"w": torch.tensor([[1.0], [1.0]])}
        opt = T.TF1Optimizer("Adam", 0.1, adam_mode=mode)
        opt.step(var, {"w": torch.tensor([[2.0], [2.0]])}, {"w": torch.tensor([0, 1])})
        before = float(var["w"][1])
        opt.step(var, {"w": torch.tensor([[2.0], [0.0]])}, {"w": torch.tensor([0])})
        assert (float(var["w"][1]) != before) == moves


def test_c_oracle_neumf_logit_is_the_restated_graph():
    """oracle/crb_oracle.c::oracle_score_pairs_neumf (the canonical fp32 order the GPU scorers are pinned to bit for bit) against the
    restated TF graph (oracle/tf1_restatement.py::neumf_logits, NeuMF.py:63-85) in fp64: NeuMF and the MLP model (E = 0)."""
    from oracle import c_oracle as O
    rs = np.random.RandomState(2)
    U, I, E, layers = 9, 11, 8, [32, 16, 8]
    Em = layers[0] // 2
    p = {"P_gmf": rs.randn(U, E), "Q_gmf": rs.randn(I, E), "P_mlp": rs.randn(U, Em), "Q_mlp": rs.randn(I, Em)}
    for k, n in enumerate(layers):
        p["W_%d" % k], p["b_%d" % k] = rs.randn(n, n // 2) * 0.3, rs.randn(n // 2) * 0.3
    p["h_neumf"] = rs.randn(E + layers[-1] // 2)
    p = {k: v.astype(np.float32) for k, v in p.items()}
    u, i = rs.randint(0, U, 50), rs.randint(0, I, 50)
    dense = np.concatenate([np.concatenate([p["W_%d" % k].ravel(), p["b_%d" % k]]) for k in range(len(layers))] + [p["h_neumf"]])
    got = O.score_pairs_neumf(p["P_gmf"], p["Q_gmf"], p["P_mlp"], p["Q_mlp"], dense, len(layers), u, i)
    want, _ = T.neumf_logits({k: torch.tensor(v).double() for k, v in p.items()}, torch.tensor(u), torch.tensor(i), len(layers))
    np.testing.assert_allclose(got, want.numpy(), rtol=2e-5, atol=2e-5)
