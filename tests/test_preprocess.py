"""cleverrec_b200/model/RankingPreprocess.py against the golden splits made by the GENUINE reference class
(oracle/make_golden.py: model/RankingPreprocess.py under np.random.seed(7) / seed(11) on dataset/ml-100k): same user / item
counts, same train lists, same test lists including the sampled evaluation negatives -- bit for bit.  The dataset file itself
lives in the reference mount (it is data, not source, and is not copied): skipped where the mount is absent."""
import logging
import os

import numpy as np
import pytest

from conftest import REFERENCE_ROOT, load_split

DATA = os.path.join(REFERENCE_ROOT, "dataset")
needs_data = pytest.mark.skipif(not os.path.exists(os.path.join(DATA, "ml-100k", "u.data")), reason="reference dataset mount absent")


def configs(**over):
    c = {"data.root_dir": DATA, "data.dataset": "ml-100k", "data.file_name": "u.data", "data.sep": "\t", "data.format": "UIRT",
         "data.user_min": "0", "data.item_min": "0", "data.split_way": "loo", "data.split_by_time": "True",
         "data.split_ratio": "[0.7,0.2,0.1]", "test.neg_samples": "99", "recommender": "BPR"}
    c.update({k: str(v) for k, v in over.items()})
    return c


def assert_same(d, got_train, got_test, want):
    assert d.user_nums == want.user_nums and d.item_nums == want.item_nums
    assert list(got_train.keys()) == list(want.ui_train.keys())
    for u in want.ui_train:
        assert got_train[u] == want.ui_train[u], u
    assert list(got_test.keys()) == list(want.ui_test.keys())
    for u in want.ui_test:
        assert got_test[u] == want.ui_test[u], u


@needs_data
@pytest.mark.parametrize("lazy", [False, True])
def test_loo_split_and_eval_negatives_bit_exact(lazy):
    from cleverrec_b200.model.RankingPreprocess import RankingPreprocess
    want = load_split("split_ml100k_loo.npz")
    ref_cfg = _reference_defaults()
    np.random.seed(7)
    d = RankingPreprocess(configs(**dict(ref_cfg, **{"data.split_way": "loo", "test.neg_samples": 99})), logging.getLogger("t"), lazy_dicts=lazy)
    assert_same(d, d.ui_train, d.ui_test, want)
    # the columns handed to Engine.build_history enumerate the dict
    tu, ti = d.train_rows
    assert tu.dtype == np.int32 and np.all(np.diff(tu) >= 0)
    assert ti.tolist() == [i for u in want.ui_train for i in want.ui_train[u]]


@needs_data
def test_random_split_bit_exact():
    from cleverrec_b200.model.RankingPreprocess import RankingPreprocess
    want = load_split("split_ml100k_rs.npz")
    np.random.seed(11)
    d = RankingPreprocess(configs(**dict(_reference_defaults(), **{"data.split_way": "rs", "test.neg_samples": 0})), logging.getLogger("t"))
    assert_same(d, d.ui_train, d.ui_test, want)


def _reference_defaults():
    """The [default] data.* keys of the reference's CleverRec.properties as oracle/make_golden.py used them."""
    import configparser as cp
    conf = cp.ConfigParser()
    conf.read(os.path.join(REFERENCE_ROOT, "CleverRec.properties"), encoding="utf-8")
    keep = {k: v for k, v in conf.items("default") if k.startswith("data.")}
    keep.update({"data.root_dir": DATA, "data.dataset": "ml-100k", "data.file_name": "u.data", "data.sep": "\t", "data.format": "UIRT",
                 "data.item_min": "0"})
    return keep


def test_filters_reindex_and_small_users(tmp_path):
    """Self-contained: user/item minimum filters, set-order re-indexing, users with <= 3 rows stay whole in training."""
    from cleverrec_b200.model.RankingPreprocess import RankingPreprocess
    rows = ["u,i,r"]                                     # consumed as the header (header=0)
    raw = {10: [5, 6, 7, 8, 9], 20: [5, 6], 30: [5, 6, 7, 9], 40: [100]}   # item 100 and user 40 fall to the filters
    for u, items in raw.items():
        rows += ["%d,%d,1" % (u, i) for i in items]
    os.makedirs(tmp_path / "toy")
    (tmp_path / "toy" / "r.csv").write_text("\n".join(rows) + "\n")
    c = configs(**{"data.root_dir": str(tmp_path), "data.dataset": "toy", "data.file_name": "r.csv", "data.sep": ",", "data.format": "UIR",
                   "data.user_min": 2, "data.item_min": 2, "data.split_by_time": "False", "test.neg_samples": 1})
    np.random.seed(0)
    d = RankingPreprocess(c, logging.getLogger("t"))
    assert d.user_nums == 3 and d.item_nums == 4          # items 5, 6, 7, 9 survive (8 and 100 occur once)
    lens = sorted(len(v) for v in d.ui_train.values())
    assert lens == [2, 3, 3]                              # user 20 (2 rows) whole; 10 -> 4 rows -> 3 + 1 test; 30 -> 4 rows -> 3 + 1
    assert len(d.ui_test) == 2
    for u, lst in d.ui_test.items():
        assert len(lst) == 2 and lst[0] not in d.ui_train[u]   # one sampled negative (unseen), then the held-out item
