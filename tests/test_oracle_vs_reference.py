"""Build-container only: the oracle restatement against the GENUINE reference code imported from /root/reference.
Skipped where the reference tree is absent (the GPU box)."""
import numpy as np
import pytest

from conftest import Data
from oracle import ref_host as H
from oracle import refimport as R

pytestmark = pytest.mark.skipif(not R.available(), reason="/root/reference not present")


def test_samplers_bit_equal(split_loo):
    ref = R.load()
    users = list(split_loo.ui_train.keys())[:60]
    sub = Data(split_loo.user_nums, split_loo.item_nums, {u: split_loo.ui_train[u] for u in users}, {})
    for fn_ref, fn_or, args in ((ref.pairwise_ranking_sampler, H.pairwise_ranking_sampler, (3, 512)),
                                (ref.pointwise_ranking_sampler, H.pointwise_ranking_sampler, (2, 512)),
                                (ref.ranking_sampler_cml, H.ranking_sampler_cml, (6, 512))):
        np.random.seed(123)
        a = fn_ref(sub, *args)
        np.random.seed(123)
        b = fn_or(sub, *args)
        assert a[0] == b[0]
        for x, y in zip(a[1:], b[1:]):
            assert np.array_equal(x, y)


def test_metrics_bit_equal():
    ref = R.load()
    rs = np.random.RandomState(0)
    for _ in range(300):
        K = int(rs.choice([1, 3, 10, 20]))
        rec = rs.permutation(40)[:K]
        real = rs.permutation(40)[:rs.randint(1, 6)].tolist()
        assert ref.cal_ranking_metrics(real, rec, K) == H.cal_ranking_metrics(real, rec, K)


def _write_log(path, rs, fmt, n_users, n_items, n_rows, id_kind):
    """A synthetic interaction log in the reference's file layout (first line is consumed as the header, RankingPreprocess.py:24-33)."""
    uid = {"dense": lambda k: k, "sparse": lambda k: 1000 + 37 * k, "offset": lambda k: k + 1}[id_kind]
    iid = {"dense": lambda k: k, "sparse": lambda k: 5 + 101 * k, "offset": lambda k: k + 1}[id_kind]
    # a long-tailed user activity so that user_min / the "<= 3 rows stay in training" rule both bite
    w = 1.0 / np.arange(1, n_users + 1)
    users = rs.choice(n_users, n_rows, p=w / w.sum())
    items = rs.randint(0, n_items, n_rows)
    rating = rs.randint(1, 6, n_rows)
    time = rs.randint(0, 50, n_rows)            # many ties: the stable sort's order matters
    lines = ["header"]
    for r in range(n_rows):
        cols = [uid(users[r]), iid(items[r])] + ([rating[r]] if fmt in ("UIR", "UIRT") else []) + ([time[r]] if fmt == "UIRT" else [])
        lines.append(",".join(str(c) for c in cols))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


_MATRIX = [
    # fmt, split_way, by_time, user_min, item_min, ratio, neg_samples, ids
    ("UIRT", "loo", "True", 0, 0, "[0.7,0.2,0.1]", 7, "dense"),
    ("UIRT", "loo", "False", 3, 0, "[0.7,0.2,0.1]", 5, "sparse"),
    ("UIR", "loo", "False", 0, 2, "[0.7,0.2,0.1]", 3, "offset"),
    ("UI", "loo", "False", 2, 2, "[0.7,0.2,0.1]", 4, "sparse"),
    ("UIRT", "rs", "True", 0, 0, "[0.7,0.2,0.1]", 0, "dense"),
    ("UIR", "rs", "False", 2, 0, "[0.8,0,0.2]", 0, "sparse"),
    ("UI", "rs", "False", 0, 3, "[0.6,0.1,0.3]", 6, "offset"),      # random split WITH sampled evaluation negatives
    ("UIRT", "rs", "True", 4, 2, "[0.8,0,0.2]", 5, "sparse"),
]


@pytest.mark.parametrize("case", range(len(_MATRIX)))
def test_packaged_preprocess_equals_the_reference_class_over_its_config_space(case, tmp_path):
    """cleverrec_b200.model.RankingPreprocess against the GENUINE reference class on synthetic logs: every data.format, both splits,
    split_by_time, both filters, both split_ratio shapes, evaluation negatives on and off, dense / sparse / 1-based raw ids -- same
    counts, same dicts (keys in the same order, lists in the same order), same NumPy stream position afterwards."""
    import logging
    from cleverrec_b200.model.RankingPreprocess import RankingPreprocess as Ours
    fmt, split_way, by_time, umin, imin, ratio, negs, ids = _MATRIX[case]
    ref = R.load()
    rs = np.random.RandomState(100 + case)
    os_dir = tmp_path / "toy"
    os_dir.mkdir()
    _write_log(str(os_dir / "log.csv"), rs, fmt, 80, 400, 900, ids)
    cfg = {"data.root_dir": str(tmp_path), "data.dataset": "toy", "data.file_name": "log.csv", "data.sep": ",", "data.format": fmt,
           "data.user_min": str(umin), "data.item_min": str(imin), "data.split_way": split_way, "data.split_by_time": by_time,
           "data.split_ratio": ratio, "test.neg_samples": str(negs), "recommender": "BPR"}
    log = logging.getLogger("t")
    np.random.seed(5 + case)
    want = ref.RankingPreprocess(dict(cfg), log)
    want_next = np.random.randint(1 << 30)
    for lazy in (False, True):
        np.random.seed(5 + case)
        got = Ours(dict(cfg), log, lazy_dicts=lazy)
        assert np.random.randint(1 << 30) == want_next           # consumed exactly the reference's random numbers
        assert (got.user_nums, got.item_nums) == (want.user_nums, want.item_nums)
        for a, b in ((got.ui_train, want.ui_train), (got.ui_test, want.ui_test)):
            assert list(a.keys()) == list(b.keys())
            for u in b:
                assert list(a[u]) == list(b[u]), (u, lazy)


# ---- the packaged evaluation loops' HOST code against the genuine reference loops, on the same scores -------------------------------
def _pos_scores(u, n):
    """Distinct scores for the n candidates of user u, by position (a candidate list may hold the same item twice)."""
    return np.random.RandomState(1000 + int(u)).permutation(n).astype(np.float32)


def _full_scores(u, n_items):
    return np.random.RandomState(5000 + int(u)).permutation(n_items).astype(np.float32)


class _RefSession(object):
    """What the reference loops call as self.sess.run(self.pre_scores, feed) (RankingRecommender.py:221,278)."""

    def __init__(self, item_nums):
        self.item_nums = item_nums

    def run(self, fetches, feed):
        users = list(feed["u_idx"])
        if "i_idx" not in feed:
            return np.stack([_full_scores(u, self.item_nums) for u in users])
        out, k = [], 0
        while k < len(users):            # a batch's users are distinct: one run of equal ids per user
            e = k
            while e < len(users) and users[e] == users[k]:
                e += 1
            out.append(_pos_scores(users[k], e - k))
            k = e
        return np.concatenate(out)


class _StubEngine(object):
    """Stands where the device library stands in the packaged class: top-K positions / ids from the same scores, ranked with the
    library's documented order (score, then index ascending).  The scores are distinct, so this is np.argsort's order too."""

    def __init__(self, data):
        import torch
        self.torch, self.device, self.data = torch, torch.device("cpu"), data

    def score_pairs_topk(self, kind, P, Q, seg_users, items, offsets, K, hvec=None, ascending=False):
        n = len(offsets) - 1
        out = np.full((n, K), -1, dtype=np.int32)
        for s in range(n):
            sc = _pos_scores(int(seg_users[s]), int(offsets[s + 1] - offsets[s]))
            order = np.argsort(sc if ascending else -sc, kind="stable")[:K]
            out[s, :order.shape[0]] = order
        return self.torch.from_numpy(out)

    def score_topk(self, kind, P, Q, rows, K, hvec=None, hist_users=None, exact=False, n_items=None, ascending=False):
        out = np.full((len(rows), K), -1, dtype=np.int32)
        for r, u in enumerate(np.asarray(rows).tolist()):
            sc = _full_scores(u, n_items)
            seen = set(self.data.ui_train.get(u, []))
            keep = [i for i in np.argsort(sc if ascending else -sc, kind="stable").tolist() if i not in seen][:K]
            out[r, :len(keep)] = keep
        return out


def _packaged_driver(data, cfg, cml_like):
    from cleverrec_b200.model.RankingRecommender import RankingRecommender

    class Model(RankingRecommender):
        def _score_spec(self):
            return (1 if cml_like else 0), None, None, None
    m = object.__new__(Model)
    m.data, m.configs, m.engine = data, cfg, _StubEngine(data)
    m.neg_samples, m.topk, m.cml_like, m.fism_like = int(cfg["test.neg_samples"]), list(map(int, cfg["topk"][1:-1].split(","))), cml_like, False
    m.batch_size_t, m.score_exact = int(cfg["test.batch_size"]), False
    m.test_users = list(data.ui_test.keys())
    m._test_cache = None
    return m


def _assert_same_metrics(got, want, n_topk):
    for a, b in zip(got, want):
        for k in range(n_topk):
            assert a[k] == b[k]          # lists of Python floats, one per test user, in test_users order: bit-equal


@pytest.mark.parametrize("cml_like", [False, True])
def test_packaged_loo_loop_host_code_equals_the_reference_loop(split_loo, cml_like):
    cfg = R.default_configs(**{"data.split_way": "loo", "test.neg_samples": 99, "topk": "[5,10,20]", "test.batch_size": 100})
    if cml_like:
        cfg["cml_like"] = "True"
    ref = R.make_driver(cfg, split_loo, _RefSession(split_loo.item_nums))
    want = ref.test_model_loo()
    got = _packaged_driver(split_loo, cfg, cml_like).test_model_loo()
    _assert_same_metrics(got, want, 3)
    assert len(got[0][0]) == len(split_loo.ui_test) and 0.0 < np.mean(got[0][2]) < 1.0


def test_packaged_rs_loop_host_code_equals_the_reference_loop(split_rs):
    cfg = R.default_configs(**{"data.split_way": "rs", "test.neg_samples": 0, "topk": "[10,20]", "test.batch_size": 128})
    users = list(split_rs.ui_test.keys())[:150]
    data = Data(split_rs.user_nums, split_rs.item_nums, split_rs.ui_train, {u: split_rs.ui_test[u] for u in users})
    ref = R.make_driver(cfg, data, _RefSession(data.item_nums))
    want = ref.test_model_rs()
    got = _packaged_driver(data, cfg, False).test_model_rs()
    _assert_same_metrics(got, want, 2)


def test_packaged_loo_loop_with_several_real_items_per_user(tmp_path):
    """data.split_way=rs with test.neg_samples > 0 goes through test_model_loo with a LIST of real items per user
    (RankingRecommender.py:283-285, :415) -- on a split produced by the genuine preprocessing."""
    import logging
    ref = R.load()
    d = tmp_path / "toy"
    d.mkdir()
    _write_log(str(d / "log.csv"), np.random.RandomState(77), "UI", 80, 400, 900, "dense")
    cfg = R.default_configs(**{"data.root_dir": str(tmp_path), "data.dataset": "toy", "data.file_name": "log.csv", "data.sep": ",", "data.format": "UI",
                               "data.user_min": 0, "data.item_min": 0, "data.split_way": "rs", "data.split_by_time": "False",
                               "data.split_ratio": "[0.6,0.1,0.3]", "test.neg_samples": 30, "topk": "[5,10]", "test.batch_size": 16})
    np.random.seed(9)
    data = ref.RankingPreprocess(dict(cfg), logging.getLogger("t"))
    assert max(len(v) for v in data.ui_test.values()) > 32
    want = R.make_driver(cfg, data, _RefSession(data.item_nums)).test_model_loo()
    got = _packaged_driver(data, cfg, False).test_model_loo()
    _assert_same_metrics(got, want, 2)


def test_packaged_loo_loop_with_fewer_candidates_than_k(split_loo):
    """test.neg_samples + 1 < topk[-1]: the reference ranks the few candidates there are (argsort[:K] is just shorter)."""
    users = list(split_loo.ui_test.keys())[:200]
    data = Data(split_loo.user_nums, split_loo.item_nums, split_loo.ui_train, {u: split_loo.ui_test[u][-4:] for u in users})
    cfg = R.default_configs(**{"data.split_way": "loo", "test.neg_samples": 3, "topk": "[2,5,10]", "test.batch_size": 64})
    want = R.make_driver(cfg, data, _RefSession(data.item_nums)).test_model_loo()
    got = _packaged_driver(data, cfg, False).test_model_loo()
    _assert_same_metrics(got, want, 3)


# ---- model constructors: the shipped conf/<Model>.properties parsed by the genuine classes and by the packaged ones --------------------
class _AnyLib(object):
    """Every C entry point 'succeeds' without doing anything: constructors only install the history (no compute on the CPU)."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def f(*a, **k):
            self.calls.append(name)
            return 0
        return f


def _stub_engine():
    import torch
    from cleverrec_b200.engine import Engine

    class Stub(Engine):
        stream = None

        def __init__(self):
            self.device, self.h, self.lib, self._hist = torch.device("cpu"), None, _AnyLib(), None
    return Stub()


_SHARED = ("embed_size", "reg", "reg1", "reg2", "reg_bias", "margin", "alpha", "beta", "atten_size", "atten_type", "mem_size", "neg_ratio",
           "neg_samples", "topk", "T", "epoches", "batch_size", "batch_size_t", "lr", "is_pairwise", "fism_like", "cml_like", "model",
           "test_users", "test_batches", "model_params", "loss_func")


@pytest.mark.parametrize("name", ["BPR", "MF", "GMF", "MLP", "NeuMF", "CML", "FISM", "NAIS_single", "TransCF", "LRML", "SBPR"])
def test_model_constructors_read_the_shipped_conf_like_the_reference(split_loo, name):
    """Every packaged model class is constructed from the reference's OWN CleverRec.properties + conf/<name>.properties (the flat
    str dict of main.py:18-25); where the genuine class can be constructed too (TensorFlow stubbed: no graph is built in __init__),
    every hyper-parameter both hold is equal, as is the parameter string run_model logs.  The reference's GMF / MLP / NeuMF cannot be
    constructed from their shipped conf files (KeyError 'reg' / 'reg1': SURVEY 2.3) and MF has no reference source (F6)."""
    import importlib
    import logging
    import sys
    from cleverrec_b200.main import load_configs
    R.load()
    data = Data(split_loo.user_nums, split_loo.item_nums, split_loo.ui_train, split_loo.ui_test)
    data.user_friends = {0: [1, 2], 1: [0]}
    cfg = load_configs(R.REFERENCE_ROOT, {"recommender": name})
    assert cfg == R.default_configs(recommender=name)
    eng = _stub_engine()
    ours = getattr(importlib.import_module("cleverrec_b200.model.ranking." + name), name)(eng, data, dict(cfg), logging.getLogger("t"))
    assert "crb_set_history" in eng.lib.calls and ours.engine is eng
    if name == "MF":
        return
    sys.path.insert(0, R.REFERENCE_ROOT)
    try:
        cls = getattr(importlib.import_module("model.ranking." + name), name)
    finally:
        sys.path.remove(R.REFERENCE_ROOT)
    if name in ("GMF", "MLP", "NeuMF"):
        with pytest.raises(KeyError):
            cls(None, data, dict(cfg), logging.getLogger("t"))
        return
    theirs = cls(None, data, dict(cfg), logging.getLogger("t"))
    compared = 0
    for k in _SHARED:
        if hasattr(theirs, k) and hasattr(ours, k):
            assert getattr(ours, k) == getattr(theirs, k), k
            compared += 1
    assert compared >= 18


# ---- run_model: epoch orchestration, test interval, best-by-NDCG@topk[0], log lines -----------------------------------------------------
class _ListHandler(object):
    def __init__(self):
        self.lines = []

    def info(self, msg):
        self.lines.append(msg)


def _scripted(cls_or_obj, losses, metrics):
    """train_model / test_model_* replaced by scripts, build_model by a no-op: what is left is run_model's own logic."""
    from collections import defaultdict
    state = {"epoch": 0, "tests": 0}

    def train_model():
        state["epoch"] += 1
        return losses[state["epoch"] - 1]

    def test_model(which):
        def f():
            HR, MRR, NDCG = defaultdict(list), defaultdict(list), defaultdict(list)
            vals = metrics[state["tests"]]
            state["tests"] += 1
            for kid, (h, m, n) in enumerate(vals):
                HR[kid].extend([h, h]); MRR[kid].extend([m, m]); NDCG[kid].extend([n, n])
            state.setdefault("called", []).append(which)
            return HR, MRR, NDCG
        return f
    cls_or_obj.build_model = lambda: None
    cls_or_obj.train_model = train_model
    cls_or_obj.test_model_loo, cls_or_obj.test_model_rs = test_model("loo"), test_model("rs")
    return state


@pytest.mark.parametrize("split_way,neg,interval", [("loo", 99, 1), ("rs", 0, 2), ("rs", 50, 3)])
def test_run_model_orchestration_and_log_lines_equal_the_reference(split_loo, split_way, neg, interval):
    from cleverrec_b200.model.RankingRecommender import RankingRecommender as Ours
    cfg = R.default_configs(**{"data.split_way": split_way, "test.neg_samples": neg, "test.interval": interval, "epoches": 7, "topk": "[10,20]"})
    losses = [9.5, 7.25, 6.125, 5.0, 4.75, 4.5, 4.25]
    # NDCG@10 rises, dips, rises again, ties its best (a tie must NOT move best_epoch), then falls
    ndcg = [0.10, 0.30, 0.20, 0.40, 0.40, 0.35, 0.05]
    metrics = [((0.5 + n, 0.2 + n, n), (0.6 + n, 0.25 + n, 0.1 + n)) for n in ndcg]

    class Sess(object):
        class graph(object):
            @staticmethod
            def finalize():
                pass

        def run(self, *a, **k):
            return None
    theirs = R.make_driver(cfg, split_loo, Sess())
    theirs.logger = _ListHandler()
    st_ref = _scripted(theirs, losses, metrics)
    theirs.run_model()

    ours = object.__new__(Ours)
    ours.data, ours.configs, ours.logger = split_loo, cfg, _ListHandler()
    ours.epoches, ours.T, ours.neg_samples, ours.topk, ours.model = 7, interval, neg, [10, 20], "BPR"
    st_ours = _scripted(ours, losses, metrics)
    best_epoch, best = ours.run_model()

    assert st_ours["called"] == st_ref["called"] and len(st_ours["called"]) == 7 // interval
    assert set(st_ours["called"]) == ({"loo"} if (split_way == "loo" or neg > 0) else {"rs"})
    assert ours.logger.lines == theirs.logger.lines           # timings are all 00:00:00 on both sides
    tested = ndcg[:7 // interval]
    assert best_epoch == interval * (1 + int(np.argmax(tested)))   # the FIRST best
    assert ("best_epoch: %d" % best_epoch) in ours.logger.lines and set(best) == {0, 1}


def test_product_metric_functions_equal_the_genuine_ones():
    """cleverrec_b200.utils.metrics (the functions the packaged classes call) against the genuine utils/metrics.py: single user and
    the vectorised batch form, real items repeated / absent, recommendation lists shorter than K (-1 padded in the batch form)."""
    from cleverrec_b200.utils.metrics import batch_ranking_metrics, cal_ranking_metrics, cal_rmse_mae
    ref = R.load()
    rs = np.random.RandomState(4)
    reals, recs, K = [], [], 10
    for _ in range(400):
        n_rec = int(rs.choice([K, K, K, 7, 3]))
        rec = rs.permutation(30)[:n_rec].astype(np.int64)
        real = rs.randint(0, 30, rs.randint(1, 7)).tolist()           # repeats allowed
        assert cal_ranking_metrics(real, rec, K) == ref.cal_ranking_metrics(real, rec, K)
        reals.append(real)
        recs.append(np.pad(rec, (0, K - n_rec), constant_values=-1))
    hr, mrr, ndcg = batch_ranking_metrics(reals, np.asarray(recs), K)
    for k in range(400):
        want = ref.cal_ranking_metrics(reals[k], recs[k][recs[k] >= 0], K)
        assert (hr[k], mrr[k], ndcg[k]) == want
    y, p = rs.rand(50).tolist(), rs.rand(50).tolist()
    assert cal_rmse_mae(y, p) == ref.cal_rmse_mae(y, p)
