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
