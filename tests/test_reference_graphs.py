"""Build-container only: the GENUINE reference model classes (model/ranking/*.py, unmodified, imported from /root/reference) are
EXECUTED -- `build_model()` builds their graph, `sess.run([self.train, self.loss], feed_dict)` runs it -- on a minimal TF-1 API shim
(oracle/tf1_shim.py: lazy graph over torch fp64, autograd, TF-1's documented optimizer rules), and oracle/tf1_restatement.py (the
oracle every CUDA training kernel is tested against) must reproduce them: loss and every variable after each of three consecutive
optimizer steps to 1e-10, for SGD / Adagrad / Adam, and `pre_scores` of the evaluation branches; the C scoring oracle
(oracle/crb_oracle.c, against which the GPU scores are bit-exact) is tied to the same `pre_scores` at fp32 accuracy.

This replaces "the restatement was transcribed by reading the reference" by a mechanical check against the reference's own
graph-building code.  It does not pin TensorFlow's kernels (summation order, fp32 rounding): DESIGN.md section 4."""
import logging

import numpy as np
import pytest
import torch

from conftest import Data
from oracle import refimport as R
from oracle import tf1_restatement as T
from oracle import tf1_shim as tf

pytestmark = pytest.mark.skipif(not R.available(), reason="/root/reference not present")

NAMES = ["BPR", "GMF", "MLP", "NeuMF", "CML", "FISM", "NAIS_single", "TransCF", "LRML", "SBPR"]
U, I, D = 24, 40, 8
LOG = logging.getLogger("refgraph")


@pytest.fixture(scope="module")
def classes():
    cls = tf.load_reference_models(R.REFERENCE_ROOT, NAMES)
    # GMF.py:48 calls get_loss but GMF.py:5-7 never imports it (reference defect): give the module the genuine function
    # its sibling modules import from utils.tools -- the one edit to the reference's namespace in this file
    cls["GMF"].__init__.__globals__["get_loss"] = cls["MLP"].__init__.__globals__["get_loss"]
    return cls


@pytest.fixture(scope="module")
def data():
    rs = np.random.RandomState(2)
    ui = {u: rs.choice(I, rs.randint(3, 9), replace=False).tolist() for u in range(U)}
    ui[3] = ui[3] + [ui[3][0]]                 # an item twice in one history: get_ui_sp_mat / get_sp_mat keep duplicates
    d = Data(U, I, ui, {u: [0] for u in range(U)})
    d.user_friends = {u: rs.choice(U, 3, replace=False).tolist() for u in range(0, U, 2)}
    return d


def _configs(name, optimizer, **over):
    base = {"init_method": "normal", "stddev": 0.3, "embed_size": D, "optimizer": optimizer, "lr": 0.05, "data.split_way": "loo",
            "test.neg_samples": 5, "batch_size": 48}
    # the keys the reference CODE reads where its shipped conf files name them differently (GMF / MLP / NeuMF: SURVEY 2.3)
    extra = {"GMF": {"reg": 0.02}, "MLP": {"reg": 0.02, "layers": "[16,8]"}, "NeuMF": {"reg1": 0.02, "reg2": 0.03, "layers": "[16,8]"},
             "NAIS_single": {"atten_size": 6}, "LRML": {"mem_size": 5}, "CML": {"neg_ratio": 4, "reg": 2.0}}.get(name, {})
    base.update(extra)
    base.update(over)
    return R.default_configs(recommender=name, **base)


def _build(classes, name, data, optimizer, patch=None, **over):
    tf.reset_default_graph()
    tf.seed_initializers(11)
    sess = tf.Session()
    m = classes[name](sess, data, _configs(name, optimizer, **over), LOG)
    if patch:
        patch(m)
    m.build_model()
    return sess, m


def _params(var_map):
    return {k: torch.tensor(v.numpy()) for k, v in var_map.items()}


def _check_steps(sess, m, var_map, feeds_of, loss_fn, hp, optimizer, sparse_index=None, extra=(), lr=0.05, steps=3):
    params = _params(var_map)
    start = _params(var_map)
    opt = T.TF1Optimizer(optimizer, lr)
    rs = np.random.RandomState(5)
    for _ in range(steps):
        feed, batch = feeds_of(rs)
        _, loss = sess.run([m.train, m.loss], feed)
        want = T.train_step(loss_fn, params, batch, hp, opt, sparse_index=sparse_index, extra=extra)
        assert abs(loss - want) <= 1e-10 * max(1.0, abs(want)), (loss, want)
        for k, v in var_map.items():
            got, ref = v.numpy(), params[k].numpy()
            assert got.shape == ref.shape, k
            assert np.max(np.abs(got - ref)) <= 1e-10 * max(1.0, float(np.max(np.abs(ref)))), k
    for k in var_map:       # every variable of the model was trained (a comparison of two untouched copies would prove nothing)
        assert float((params[k] - start[k]).abs().max()) > 1e-6, k
    return params


def _t(a):
    return torch.tensor(np.asarray(a).astype(np.int64))


OPTS = ["SGD", "Adagrad", "Adam"]


@pytest.mark.parametrize("optimizer", OPTS)
def test_bpr_graph(classes, data, optimizer):
    sess, m = _build(classes, "BPR", data, optimizer)

    def feeds(rs):
        u, i, j = rs.randint(0, U, 48), rs.randint(0, I, 48), rs.randint(0, I, 48)
        return {m.u_idx: u, m.i_idx: i, m.j_idx: j}, {"u": _t(u), "i": _t(i), "j": _t(j)}
    p = _check_steps(sess, m, {"P": m.P, "Q": m.Q}, feeds, T.bpr_loss, {"reg": m.reg}, optimizer, {"P": ["u"], "Q": ["i", "j"]})
    # _predict, loo branch (BPR.py:49) and the all-item branch (:51) on a second graph over the same variables' values
    u, i = np.arange(10), np.arange(10) + 3
    got = sess.run(m.pre_scores, {m.u_idx: u, m.i_idx: i, m.j_idx: i})
    np.testing.assert_allclose(got, T.bpr_scores_pairs(p["P"], p["Q"], _t(u), _t(i)).numpy(), rtol=1e-12)
    sess2, m2 = _build(classes, "BPR", data, optimizer, **{"data.split_way": "rs", "test.neg_samples": 0})
    m2.P.assign_value(p["P"].numpy()); m2.Q.assign_value(p["Q"].numpy())
    full = sess2.run(m2.pre_scores, {m2.u_idx: u, m2.batch_size_t_: len(u)})
    np.testing.assert_allclose(full, T.bpr_scores_all(p["P"], p["Q"], _t(u)).numpy(), rtol=1e-12)
    # the C scoring oracle (canonical fp32 fma chain) computes the same quantity
    from oracle import c_oracle as O
    uu, ii = np.repeat(u, I), np.tile(np.arange(I), len(u))
    c = O.score_pairs(0, p["P"].numpy(), p["Q"].numpy(), uu, ii).reshape(len(u), I)
    np.testing.assert_allclose(c, full, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("optimizer", OPTS)
@pytest.mark.parametrize("loss", ["cross_entropy", "square"])
def test_gmf_graph(classes, data, optimizer, loss):
    sess, m = _build(classes, "GMF", data, optimizer, loss_func=loss)

    def feeds(rs):
        u, i, y = rs.randint(0, U, 48), rs.randint(0, I, 48), rs.randint(0, 2, 48).astype(np.float64)
        return {m.u_idx: u, m.i_idx: i, m.y: y}, {"u": _t(u), "i": _t(i), "y": torch.tensor(y)}
    p = _check_steps(sess, m, {"P": m.P, "Q": m.Q, "h": m.h_gmf}, feeds, T.gmf_loss, {"reg": m.reg, "loss_func": loss}, optimizer,
                     {"P": ["u"], "Q": ["i"]})
    u, i = np.arange(12), np.arange(12) + 5
    got = sess.run(m.pre_scores, {m.u_idx: u, m.i_idx: i, m.y: np.zeros(12)})          # sigmoid(logits), GMF.py:51-57
    logits = T.gmf_logits_pairs(p["P"], p["Q"], p["h"], _t(u), _t(i))
    np.testing.assert_allclose(got, torch.sigmoid(logits).numpy(), rtol=1e-12)
    from oracle import c_oracle as O
    c = O.score_pairs(1, p["P"].numpy(), p["Q"].numpy(), u, i, hvec=p["h"].numpy())  # the library ranks on the logit (monotone)
    np.testing.assert_allclose(c, logits.numpy(), rtol=2e-5, atol=2e-6)


def _tower_vars(m, prefix=""):
    out = {}
    for k in range(len(m.layers)):
        out["W_%d" % k], out["b_%d" % k] = m.mlp_params["W_%d" % k], m.mlp_params["b_%d" % k]
    return out


@pytest.mark.parametrize("optimizer", OPTS)
def test_mlp_graph(classes, data, optimizer):
    sess, m = _build(classes, "MLP", data, optimizer)
    var_map = dict({"P": m.P, "Q": m.Q, "h_mlp": m.h_mlp}, **_tower_vars(m))

    def feeds(rs):
        u, i, y = rs.randint(0, U, 48), rs.randint(0, I, 48), rs.randint(0, 2, 48).astype(np.float64)
        return {m.u_idx: u, m.i_idx: i, m.y: y}, {"u": _t(u), "i": _t(i), "y": torch.tensor(y)}
    p = _check_steps(sess, m, var_map, feeds, T.mlp_loss, {"reg": m.reg, "loss_func": "cross_entropy", "n_layers": 2}, optimizer,
                     {"P": ["u"], "Q": ["i"]})
    # the C oracle's MLP logit (what both GPU scorers reproduce bit for bit) against the reference's pre_scores = sigmoid(logits)
    from oracle import c_oracle as O
    u, i = np.arange(12), np.arange(12) + 7
    got = sess.run(m.pre_scores, {m.u_idx: u, m.i_idx: i, m.y: np.zeros(12)})
    dense = np.concatenate([p[k].numpy().reshape(-1) for k in ("W_0", "b_0", "W_1", "b_1", "h_mlp")])
    c = O.score_pairs_neumf(None, None, p["P"].numpy(), p["Q"].numpy(), dense, 2, u, i)
    np.testing.assert_allclose(1.0 / (1.0 + np.exp(-c.astype(np.float64))), got, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("optimizer", OPTS)
def test_neumf_graph(classes, data, optimizer):
    sess, m = _build(classes, "NeuMF", data, optimizer)
    var_map = dict({"P_gmf": m.P_gmf, "Q_gmf": m.Q_gmf, "P_mlp": m.P_mlp, "Q_mlp": m.Q_mlp, "h_neumf": m.h_neumf}, **_tower_vars(m))

    def feeds(rs):
        u, i, y = rs.randint(0, U, 48), rs.randint(0, I, 48), rs.randint(0, 2, 48).astype(np.float64)
        return {m.u_idx: u, m.i_idx: i, m.y: y}, {"u": _t(u), "i": _t(i), "y": torch.tensor(y)}
    hp = {"reg1": m.reg1, "reg2": m.reg2, "loss_func": "cross_entropy", "n_layers": 2}
    p = _check_steps(sess, m, var_map, feeds, T.neumf_loss, hp, optimizer, {"P_gmf": ["u"], "Q_gmf": ["i"], "P_mlp": ["u"], "Q_mlp": ["i"]})
    # h_gmf / h_mlp exist in the reference graph (NeuMF.py:40,49) but reach no loss unless pretrained: they must not have moved
    from oracle import c_oracle as O
    u, i = np.arange(12), np.arange(12) + 7
    got = sess.run(m.pre_scores, {m.u_idx: u, m.i_idx: i, m.y: np.zeros(12)})
    dense = np.concatenate([p[k].numpy().reshape(-1) for k in ("W_0", "b_0", "W_1", "b_1", "h_neumf")])
    c = O.score_pairs_neumf(p["P_gmf"].numpy(), p["Q_gmf"].numpy(), p["P_mlp"].numpy(), p["Q_mlp"].numpy(), dense, 2, u, i)
    np.testing.assert_allclose(1.0 / (1.0 + np.exp(-c.astype(np.float64))), got, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("optimizer", OPTS)
def test_cml_graph(classes, data, optimizer):
    sess, m = _build(classes, "CML", data, optimizer)
    R_ = m.neg_ratio

    def feeds(rs):
        u, i, neg = rs.randint(0, U, 48), rs.randint(0, I, 48), rs.randint(0, I, (48, R_))
        return {m.u_idx: u, m.i_idx: i, m.neg_items: neg}, {"u": _t(u), "i": _t(i), "neg": _t(neg)}
    hp = {"reg": m.reg, "margin": m.margin, "item_nums": I, "neg_ratio": R_}
    p = _check_steps(sess, m, {"P": m.P, "Q": m.Q}, feeds, T.cml_loss, hp, optimizer)     # covariance term: both tables dense
    u, i = np.arange(10), np.arange(10) + 2
    got = sess.run(m.pre_scores, {m.u_idx: u, m.i_idx: i, m.neg_items: np.zeros((10, R_), dtype=np.int64)})
    np.testing.assert_allclose(got, T.cml_dist_pairs(p["P"], p["Q"], _t(u), _t(i)).numpy(), rtol=1e-12)
    from oracle import c_oracle as O
    np.testing.assert_allclose(O.score_pairs(2, p["P"].numpy(), p["Q"].numpy(), u, i), got, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("optimizer", OPTS)
def test_fism_graph(classes, data, optimizer):
    sess, m = _build(classes, "FISM", data, optimizer)
    sp = m.ui_sp_mat                                    # the genuine get_ui_sp_mat(data) (utils/tools.py:90-97), as built on the shim
    hist = (sp.rows, sp.cols, sp.values)
    assert sp.dense_shape == (U + 1, I + 1) and sp.rows.numel() == sum(len(v) for v in data.ui_train.values())

    def feeds(rs):
        u, i, j = rs.randint(0, U, 48), rs.randint(0, I, 48), rs.randint(0, I, 48)
        nbr = np.asarray([len(data.ui_train[x]) for x in u])
        return ({m.u_idx: u, m.i_idx: i, m.j_idx: j, m.u_neighbors_num: nbr},
                {"u": _t(u), "i": _t(i), "j": _t(j), "nbr_num": _t(nbr)})
    hp = {"reg": m.reg, "reg_bias": m.reg_bias, "alpha": m.alpha, "batch_size": m.batch_size, "user_nums": U, "loss_func": "bpr"}
    p = _check_steps(sess, m, {"P": m.P, "Q": m.Q, "b": m.b}, feeds, T.fism_loss, hp, optimizer, extra=(hist,))
    # pre_scores = ui_scores (FISM.py:66-68): q_i . (n_u^-alpha * sum of the history's P rows) + b_i -- the library scores it as
    # SCORE_DOT_BIAS over the user vectors of crb_fism_user_vectors; the C oracle's kind 3 is that scorer's oracle
    u, i = np.arange(10), np.arange(10) + 4
    nbr = np.asarray([len(data.ui_train[x]) for x in u])
    got = sess.run(m.pre_scores, {m.u_idx: u, m.i_idx: i, m.j_idx: i, m.u_neighbors_num: nbr})
    s_u = T.fism_user_embed(p, {"u": _t(u), "nbr_num": _t(nbr)}, hp, hist)
    np.testing.assert_allclose(got, ((p["Q"][_t(i)] * s_u).sum(1) + p["b"][_t(i)]).numpy(), rtol=1e-12)
    from oracle import c_oracle as O
    c = O.score_pairs(3, s_u.numpy(), p["Q"].numpy(), np.arange(10), i, hvec=p["b"].numpy())
    np.testing.assert_allclose(c, got, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("optimizer", OPTS)
@pytest.mark.parametrize("atten", ["prod", "concat"])
def test_nais_graph(classes, data, optimizer, atten):
    def patch(m):
        # NAIS_single.py:87 CALLS self.loss_func, which Recommender.py:23 set to the config STRING: the evident intent (and the
        # restatement's parity decision) is tf.nn.sigmoid_cross_entropy_with_logits
        m.loss_func = tf.nn.sigmoid_cross_entropy_with_logits
    sess, m = _build(classes, "NAIS_single", data, optimizer, patch=patch, atten_type=atten)
    var_map = {"P": m.P, "Q": m.Q, "bias": m.bias, "W": m.W, "b_att": m.b, "h": m.h}

    def feeds(rs):
        hist = np.asarray(data.ui_train[int(rs.randint(0, U))])
        tg = rs.randint(0, I, 15)
        y = rs.randint(0, 2, 15).astype(np.float64)
        return ({m.u_idx: hist, m.u_nbrs_num: len(hist), m.i_idx: tg, m.i_nums: len(tg), m.y: y},
                {"hist": _t(hist), "i": _t(tg), "y": torch.tensor(y)})
    hp = {"reg": m.reg, "beta": m.beta, "atten_type": atten}
    p = _check_steps(sess, m, var_map, feeds, T.nais_loss, hp, optimizer, {"P": ["hist"], "Q": ["i"], "bias": ["i"]})
    hist, tg = np.asarray(data.ui_train[5]), np.arange(12) + 9                      # pre_scores = ui_scores (NAIS_single.py:92-94)
    got = sess.run(m.pre_scores, {m.u_idx: hist, m.u_nbrs_num: len(hist), m.i_idx: tg, m.i_nums: len(tg), m.y: np.zeros(12)})
    q = p["Q"][_t(tg)]
    want = (T.nais_user_embed(p, _t(hist), q, hp) * q).sum(1) + p["bias"][_t(tg)]
    np.testing.assert_allclose(got, want.numpy(), rtol=1e-12)


@pytest.mark.parametrize("optimizer", OPTS)
def test_transcf_graph(classes, data, optimizer):
    sess, m = _build(classes, "TransCF", data, optimizer)
    ui, iu = m.ui_sp_mat, m.iu_sp_mat                  # the genuine get_sp_mat(data) (utils/tools.py:100-113)
    assert ui.dense_shape == (U, I) and iu.dense_shape == (I, U)

    def feeds(rs):
        u, i, j = rs.randint(0, U, 48), rs.randint(0, I, 48), rs.randint(0, I, 48)
        return {m.u_idx: u, m.i_idx: i, m.j_idx: j}, {"u": _t(u), "i": _t(i), "j": _t(j)}
    hp = {"reg1": m.reg1, "reg2": m.reg2, "margin": m.margin, "user_nums": U, "item_nums": I}
    coo = ((ui.rows, ui.cols, ui.values), (iu.rows, iu.cols, iu.values))
    p = _check_steps(sess, m, {"P": m.P, "Q": m.Q}, feeds, T.transcf_loss, hp, optimizer, extra=coo)
    u, i = np.arange(10), np.arange(10) + 6                                          # pre_scores = ui_dist (TransCF.py:80-81)
    got = sess.run(m.pre_scores, {m.u_idx: u, m.i_idx: i, m.j_idx: i})
    all_u, all_i = T.transcf_parts(p, None, hp, *coo)
    want = ((p["P"][_t(u)] + all_u[_t(u)] * all_i[_t(i)] - p["Q"][_t(i)]) ** 2).sum(1)
    np.testing.assert_allclose(got, want.numpy(), rtol=1e-12)


@pytest.mark.parametrize("optimizer", OPTS)
def test_lrml_graph(classes, data, optimizer):
    sess, m = _build(classes, "LRML", data, optimizer)

    def feeds(rs):
        u, i, j = rs.randint(0, U, 48), rs.randint(0, I, 48), rs.randint(0, I, 48)
        return {m.u_idx: u, m.i_idx: i, m.j_idx: j}, {"u": _t(u), "i": _t(i), "j": _t(j)}
    p = _check_steps(sess, m, {"P": m.P, "Q": m.Q, "K": m.K, "M": m.M}, feeds, T.lrml_loss, {"reg": m.reg, "margin": m.margin}, optimizer,
                     {"P": ["u"], "Q": ["i", "j"]})
    u, i = np.arange(10), np.arange(10) + 6                                          # pre_scores = ui_dist (LRML.py:72-73)
    got = sess.run(m.pre_scores, {m.u_idx: u, m.i_idx: i, m.j_idx: i})
    np.testing.assert_allclose(got, T.lrml_dist(p, p["P"][_t(u)], p["Q"][_t(i)]).numpy(), rtol=1e-12)


@pytest.mark.parametrize("optimizer", OPTS)
def test_sbpr_graph(classes, data, optimizer):
    sess, m = _build(classes, "SBPR", data, optimizer)
    assert m.SPu                                          # the genuine get_SPu ran in the constructor

    def feeds(rs):
        u, i, k, j = rs.randint(0, U, 48), rs.randint(0, I, 48), rs.randint(0, I, 48), rs.randint(0, I, 48)
        suk = rs.randint(1, 4, 48).astype(np.float64)
        return ({m.u_idx: u, m.i_idx: i, m.i_s_idx: k, m.i_neg_idx: j, m.suk: suk},
                {"u": _t(u), "i": _t(i), "k": _t(k), "j": _t(j), "suk": torch.tensor(suk)})
    p = _check_steps(sess, m, {"P": m.P, "Q": m.Q, "bias": m.bias}, feeds, T.sbpr_loss, {"reg": m.reg}, optimizer,
                     {"P": ["u"], "Q": ["i", "k", "j"], "bias": ["i", "k", "j"]})
    assert p["bias"].shape[0] == I + 1                    # SBPR.py:36
    u, i = np.arange(10), np.arange(10) + 6                                          # pre_scores = ui_scores (SBPR.py:60-61)
    got = sess.run(m.pre_scores, {m.u_idx: u, m.i_idx: i, m.i_s_idx: i, m.i_neg_idx: i, m.suk: np.ones(10)})
    np.testing.assert_allclose(got, ((p["P"][_t(u)] * p["Q"][_t(i)]).sum(1) + p["bias"][_t(i)]).numpy(), rtol=1e-12)
    from oracle import c_oracle as O
    np.testing.assert_allclose(O.score_pairs(3, p["P"].numpy(), p["Q"].numpy(), u, i, hvec=p["bias"].numpy()), got, rtol=2e-5, atol=2e-6)


def test_a_whole_reference_epoch_runs_on_the_shim(classes, data):
    """RankingRecommender.train_model itself (the genuine sampler + batch loop + sess.run, :33-61) over the shim: the epoch loss is the
    mean of the per-step sums, and equals the restatement driven with the same sampled epoch."""
    from oracle import ref_host as H
    sess, m = _build(classes, "BPR", data, "Adam", neg_ratio=2, batch_size=32)
    p = _params({"P": m.P, "Q": m.Q})
    np.random.seed(4)
    got = m.train_model()
    np.random.seed(4)
    n_b, u, i, j = H.pairwise_ranking_sampler(data, 2, 32)[:4]
    opt, total = T.TF1Optimizer("Adam", 0.05), 0.0
    for k in range(n_b):
        b = {n: _t(a[k * 32:(k + 1) * 32]) for n, a in (("u", u), ("i", i), ("j", j))}
        total += T.train_step(T.bpr_loss, p, b, {"reg": m.reg}, opt, sparse_index={"P": ["u"], "Q": ["i", "j"]})
    assert abs(got - total / n_b) <= 1e-10 * abs(got)
    np.testing.assert_allclose(m.P.numpy(), p["P"].numpy(), rtol=0, atol=1e-10)


@pytest.mark.parametrize("name", ["GMF", "MLP", "NeuMF", "CML", "FISM", "TransCF", "LRML", "SBPR"])
def test_all_item_predict_branch_equals_the_pair_branch(classes, data, name):
    """`_predict` has two branches in every model: candidate pairs (loo / sampled negatives) and all items (random split).  The
    library serves both with ONE scorer per model, so the reference's two graphs must give the same number for the same (u, i) --
    except where they verifiably do not, and then the packaged class must follow (CML / TransCF: clipped user row; SBPR: no bias):
    built twice on the shim from the same seed (identical variables), all-item scores [n, I(+1)] against pair scores of every (u, i).
    (Real TF-1 would refuse MLP / NeuMF / LRML's rank-3 x rank-2 tf.matmul in the all-item branch; torch broadcasts it.)"""
    sess_a, a = _build(classes, name, data, "SGD")
    sess_b, b = _build(classes, name, data, "SGD", **{"data.split_way": "rs", "test.neg_samples": 0})
    users = np.asarray([0, 3, 7, 11])
    n_items = I + 1 if name == "FISM" else I
    if name == "SBPR":      # the bias starts at zero (SBPR.py:36): give it a trained-looking value in both graphs
        bias = np.random.RandomState(1).randn(I + 1) * 0.3
        a.bias.assign_value(bias); b.bias.assign_value(bias)
    uu, ii = np.repeat(users, n_items), np.tile(np.arange(n_items), len(users))
    zeros = np.zeros(len(uu))

    def feed(m, u, i=None):
        f = {m.u_idx: u}
        if i is not None:
            f[m.i_idx] = i
        for attr, val in (("j_idx", i), ("y", zeros), ("i_s_idx", i), ("i_neg_idx", i), ("suk", zeros + 1),
                          ("neg_items", None if i is None else np.zeros((len(u), m.neg_ratio), dtype=np.int64))):
            if hasattr(m, attr) and val is not None:
                f[getattr(m, attr)] = val
        if hasattr(m, "u_neighbors_num"):
            f[m.u_neighbors_num] = np.asarray([len(data.ui_train[x]) for x in u])
        f[m.batch_size_t_] = len(u)
        return f
    pairs = sess_a.run(a.pre_scores, feed(a, uu, ii)).reshape(len(users), n_items)
    full = sess_b.run(b.pre_scores, feed(b, users))
    assert full.shape == pairs.shape
    if name == "SBPR":
        # SBPR.py:59-63: the candidate branch is p.q + bias, the all-item branch the plain matmul WITHOUT the bias
        # (the packaged class: SCORE_DOT_BIAS vs SCORE_DOT in SBPR._score_spec)
        np.testing.assert_allclose(full, pairs - bias[None, :I], rtol=1e-11, atol=1e-13)
        assert not np.allclose(full, pairs)
        return
    if name in ("CML", "TransCF"):
        # CML.py:58-61,84 / TransCF.py:59-62,83-85: _unit_clipping rebinds self.u_embed to clip_by_norm(u_embed, 1) before _predict is
        # built, so the all-item branch scores the CLIPPED user row (Q, and TransCF's neighbourhood means, stay unclipped) while the
        # pair branch was built from the unclipped one.  The packaged classes do the same (CML._score_spec, TransCF.test_model_rs).
        P, Q = torch.tensor(a.P.numpy()), torch.tensor(a.Q.numpy())
        norms = P.norm(dim=1, keepdim=True)
        assert float(norms[users].max()) > 1.0 > float(norms[users].min())         # both cases are exercised
        Pc = P / torch.clamp(norms, min=1.0)
        uu_t, ii_t = _t(uu), _t(ii)
        if name == "CML":
            want = T.cml_dist_pairs(Pc, Q, uu_t, ii_t)
        else:
            coo = ((a.ui_sp_mat.rows, a.ui_sp_mat.cols, a.ui_sp_mat.values), (a.iu_sp_mat.rows, a.iu_sp_mat.cols, a.iu_sp_mat.values))
            all_u, all_i = T.transcf_parts({"P": P, "Q": Q}, None, {"user_nums": U, "item_nums": I}, *coo)   # from the UNclipped P
            want = ((Pc[uu_t] + all_u[uu_t] * all_i[ii_t] - Q[ii_t]) ** 2).sum(1)
        np.testing.assert_allclose(full, want.numpy().reshape(len(users), n_items), rtol=1e-11, atol=1e-13)
        assert not np.allclose(full, pairs)
        return
    np.testing.assert_allclose(full, pairs, rtol=1e-11, atol=1e-13)


def _epoch_against(params, loss_fn, hp, opt, batches, sparse_index=None, extra=()):
    total, n = 0.0, 0
    for b in batches:
        total += T.train_step(loss_fn, params, b, hp, opt, sparse_index=sparse_index, extra=extra)
        n += 1
    return total, n


def test_reference_epoch_loops_of_the_special_models_run_on_the_shim(classes, data):
    """The other genuine epoch loops -- train_model (pointwise branch, GMF), train_model_cml, train_model_nais, train_model_sbpr
    (RankingRecommender.py:48-117): genuine sampler code + genuine batch slicing + genuine graph, against the restated samplers
    (oracle/ref_host.py) feeding the restated graphs under the same NumPy seed."""
    from oracle import ref_host as H
    # GMF, pointwise
    sess, m = _build(classes, "GMF", data, "Adagrad", neg_ratio=2, batch_size=40)
    p = _params({"P": m.P, "Q": m.Q, "h": m.h_gmf})
    np.random.seed(8)
    got = m.train_model()
    np.random.seed(8)
    n_b, u, i, y = H.pointwise_ranking_sampler(data, 2, 40)
    batches = [{"u": _t(u[k * 40:(k + 1) * 40]), "i": _t(i[k * 40:(k + 1) * 40]), "y": torch.tensor(y[k * 40:(k + 1) * 40].astype(np.float64))}
               for k in range(n_b)]
    total, n = _epoch_against(p, T.gmf_loss, {"reg": m.reg, "loss_func": "cross_entropy"}, T.TF1Optimizer("Adagrad", 0.05), batches, {"P": ["u"], "Q": ["i"]})
    assert n == n_b and abs(got - total / n) <= 1e-10 * abs(got)
    np.testing.assert_allclose(m.h_gmf.numpy(), p["h"].numpy(), rtol=0, atol=1e-10)

    # CML
    sess, m = _build(classes, "CML", data, "Adam", batch_size=40)
    p = _params({"P": m.P, "Q": m.Q})
    np.random.seed(9)
    got = m.train_model()            # = train_model_cml (CML.py:15)
    np.random.seed(9)
    n_b, u, i, neg = H.ranking_sampler_cml(data, m.neg_ratio, 40)
    batches = [{"u": _t(u[k * 40:(k + 1) * 40]), "i": _t(i[k * 40:(k + 1) * 40]), "neg": _t(neg[k * 40:(k + 1) * 40])} for k in range(n_b)]
    hp = {"reg": m.reg, "margin": m.margin, "item_nums": I, "neg_ratio": m.neg_ratio}
    total, n = _epoch_against(p, T.cml_loss, hp, T.TF1Optimizer("Adam", 0.05), batches)
    assert abs(got - total / n) <= 1e-10 * abs(got)
    np.testing.assert_allclose(m.Q.numpy(), p["Q"].numpy(), rtol=0, atol=1e-10)

    # NAIS_single: one step per user
    def patch(mm):
        mm.loss_func = tf.nn.sigmoid_cross_entropy_with_logits
    sess, m = _build(classes, "NAIS_single", data, "Adagrad", patch=patch, neg_ratio=2)
    p = _params({"P": m.P, "Q": m.Q, "bias": m.bias, "W": m.W, "b_att": m.b, "h": m.h})
    np.random.seed(10)
    got = m.train_model()            # = train_model_nais
    np.random.seed(10)
    batches = [{"hist": _t(hist), "i": _t(tg), "y": torch.tensor(np.asarray(y, dtype=np.float64))} for _, hist, tg, y in H.nais_user_batches(data, 2)]
    total, n = _epoch_against(p, T.nais_loss, {"reg": m.reg, "beta": m.beta, "atten_type": m.atten_type}, T.TF1Optimizer("Adagrad", 0.05), batches,
                              {"P": ["hist"], "Q": ["i"], "bias": ["i"]})
    assert n == len(data.ui_train) and abs(got - total / n) <= 1e-10 * abs(got)
    np.testing.assert_allclose(m.W.numpy(), p["W"].numpy(), rtol=0, atol=1e-10)

    # SBPR
    sess, m = _build(classes, "SBPR", data, "SGD", neg_ratio=2, batch_size=40)
    p = _params({"P": m.P, "Q": m.Q, "bias": m.bias})
    np.random.seed(12)
    got = m.train_model()            # = train_model_sbpr
    np.random.seed(12)
    n_b, u, i, k_, j, suk = H.ranking_sampler_sbpr(data, H.get_SPu(data), 2, 40)
    sl = lambda a, q: a[q * 40:(q + 1) * 40]          # noqa: E731
    batches = [{"u": _t(sl(u, q)), "i": _t(sl(i, q)), "k": _t(sl(k_, q)), "j": _t(sl(j, q)), "suk": torch.tensor(sl(suk, q).astype(np.float64))}
               for q in range(n_b)]
    total, n = _epoch_against(p, T.sbpr_loss, {"reg": m.reg}, T.TF1Optimizer("SGD", 0.05), batches,
                              {"P": ["u"], "Q": ["i", "k", "j"], "bias": ["i", "k", "j"]})
    assert abs(got - total / n) <= 1e-10 * abs(got)
    np.testing.assert_allclose(m.bias.numpy(), p["bias"].numpy(), rtol=0, atol=1e-10)


def test_reference_fism_epoch_runs_on_the_shim(classes, data):
    """train_model's pairwise branch with fism_like (RankingRecommender.py:36-46): the sampler's fifth array feeds u_neighbors_num."""
    from oracle import ref_host as H
    sess, m = _build(classes, "FISM", data, "Adam", neg_ratio=2, batch_size=40)
    assert m.fism_like and m.is_pairwise == "True"
    sp = m.ui_sp_mat
    p = _params({"P": m.P, "Q": m.Q, "b": m.b})
    np.random.seed(13)
    got = m.train_model()
    np.random.seed(13)
    n_b, u, i, j, nbr = H.pairwise_ranking_sampler(data, 2, 40, fism_like=True)
    sl = lambda a, q: _t(a[q * 40:(q + 1) * 40])          # noqa: E731
    batches = [{"u": sl(u, q), "i": sl(i, q), "j": sl(j, q), "nbr_num": sl(nbr, q)} for q in range(n_b)]
    hp = {"reg": m.reg, "reg_bias": m.reg_bias, "alpha": m.alpha, "batch_size": m.batch_size, "user_nums": U, "loss_func": "bpr"}
    total, n = _epoch_against(p, T.fism_loss, hp, T.TF1Optimizer("Adam", 0.05), batches, extra=((sp.rows, sp.cols, sp.values),))
    assert abs(got - total / n) <= 1e-10 * abs(got)
    np.testing.assert_allclose(m.P.numpy(), p["P"].numpy(), rtol=0, atol=1e-10)


@pytest.mark.parametrize("name", NAMES)
def test_checkpoint_variable_names_are_the_reference_savers(classes, data, name):
    """The names under which the packaged classes save / restore (`_variables`, one .npz keyed by Saver name) are the keys of the
    var_list the GENUINE `_save_model` hands to tf.train.Saver (typos such as 'FISM_paras/P' and 'NAIS_paras/P' included, because
    NAIS restores FISM by that spelling, NAIS_single.py:35-38)."""
    import os
    import re
    from conftest import ROOT

    def patch(m):
        if name == "NAIS_single":
            m.loss_func = tf.nn.sigmoid_cross_entropy_with_logits
        m.saved_model_dir = os.path.join(os.environ.get("TMPDIR", "/tmp"), "crb_refgraph_saver")    # _save_model makes the directory
    sess, m = _build(classes, name, data, "SGD", patch=patch)
    if not hasattr(m, "saver"):
        m._save_model()              # TransCF / LRML: the call is commented out in build_model (TransCF.py:99, LRML.py:92)
    keys = set(m.saver.var_list.keys())
    assert len(keys) >= 2
    src = open(os.path.join(ROOT, "cleverrec_b200", "model", "ranking", name + ".py")).read()
    if name in ("GMF", "MLP", "NeuMF"):     # built from '<Model>_params/' + the variable's own name
        prefix = {"GMF": "%s_params/", "MLP": "MLP_params/", "NeuMF": "NeuMF_params/"}[name]
        assert prefix in src
        assert {k.split("/")[0] for k in keys} == {name + "_params"}
        tower = {"W_%d" % k for k in range(len(getattr(m, "layers", [])))} | {"b_%d" % k for k in range(len(getattr(m, "layers", [])))}
        for k in keys:
            leaf = k.split("/")[1]
            assert leaf in tower or re.search(r"['/]%s'" % re.escape(leaf), src) or ("'%s'" % leaf) in src, k
        return
    for k in keys:
        assert ("'%s'" % k) in src, k


@pytest.mark.parametrize("atten", ["prod", "concat"])
def test_nais_all_item_branch_equals_the_pair_branch(classes, data, atten):
    """NAIS_single.py:92-97: the random-split branch attends over the user's history once per row of Q (item_nums + 1 rows); it is
    the candidate branch fed with every item as target, which is how the packaged class scores it (crb_score_nais)."""
    def patch(m):
        m.loss_func = tf.nn.sigmoid_cross_entropy_with_logits
    sess_a, a = _build(classes, "NAIS_single", data, "SGD", patch=patch, atten_type=atten)
    sess_b, b = _build(classes, "NAIS_single", data, "SGD", patch=patch, atten_type=atten, **{"data.split_way": "rs", "test.neg_samples": 0})
    hist = np.asarray(data.ui_train[7])
    tg = np.arange(I + 1)
    pairs = sess_a.run(a.pre_scores, {a.u_idx: hist, a.u_nbrs_num: len(hist), a.i_idx: tg, a.i_nums: len(tg), a.y: np.zeros(len(tg))})
    full = sess_b.run(b.pre_scores, {b.u_idx: hist, b.u_nbrs_num: len(hist), b.i_idx: tg[:1], b.i_nums: 1, b.y: np.zeros(1)})
    assert full.shape == (I + 1,)
    np.testing.assert_allclose(full, pairs, rtol=1e-11, atol=1e-13)


def test_the_reference_runs_end_to_end_on_the_shim(classes, split_loo):
    """`run_model()` of the genuine BPR class, untouched: genuine sampler, genuine batch loop, genuine graph (on the shim), genuine
    evaluation loop and metrics, on the ml-100k leave-one-out split its own preprocessing produced (golden fixture).  The reference
    learns: HR@10 over 99 sampled negatives leaves the 0.10 of a random ranking behind."""
    users = list(split_loo.ui_train.keys())[:300]
    data = Data(split_loo.user_nums, split_loo.item_nums, {u: split_loo.ui_train[u] for u in users},
                {u: split_loo.ui_test[u] for u in users if u in split_loo.ui_test})
    tf.reset_default_graph()
    tf.seed_initializers(3)
    cfg = R.default_configs(recommender="BPR", **{"init_method": "normal", "stddev": 0.01, "embed_size": 16, "optimizer": "Adam", "lr": 0.01,
                                                  "epoches": 3, "batch_size": 2048, "neg_ratio": 2, "test.interval": 1, "topk": "[10,20]",
                                                  "data.split_way": "loo", "test.neg_samples": 99, "test.batch_size": 128})

    class Log(object):
        lines = []

        def info(self, msg):
            Log.lines.append(msg)
    m = classes["BPR"](tf.Session(), data, cfg, Log())
    np.random.seed(0)
    m.run_model()
    text = "\n".join(Log.lines)
    import re
    hr10 = [float(x) for x in re.findall(r"\(k=10\) HR=([0-9.]+)", text)]
    losses = [float(x) for x in re.findall(r"Training loss: ([0-9.]+)", text)]
    assert len(hr10) == 4 and len(losses) == 3          # three epochs + the final summary line
    assert losses[2] < losses[0]
    assert max(hr10[:3]) > 0.2 and hr10[2] > hr10[0] - 0.02
    assert "best_epoch: " in text
