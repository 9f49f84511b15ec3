"""-m gpu: the device preprocessing (csrc/preprocess.cu: filter, re-index, leave-one-out split, evaluation negatives) against
(i) the golden ml-100k leave-one-out split made by the GENUINE reference class, rebuilt here as an interaction log (the dataset file
is not on the GPU box), (ii) the packaged host mirror of model/RankingPreprocess.py (itself bit-exact against the golden splits in
the CPU tests) on synthetic logs with user_min / item_min filters, sparse and negative raw ids, time ties, and (iii) the
integer-exact twin of the evaluation-negative sampler."""
import logging
import os

import numpy as np
import pytest
import torch

from conftest import load_split
from oracle import philox as X

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cleverrec_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def test_loo_split_of_the_golden_ml100k_log(eng):
    """A log that the reference's preprocessing turns into the golden split: per user the training items in order, then the test
    positive, users interleaved, raw ids = a sparse increasing function of the golden ids.  The device pipeline must give back the
    golden training lists and test positives."""
    want = load_split("split_ml100k_loo.npz")
    users = sorted(set(want.ui_train) | set(want.ui_test))
    per_user = {u: list(want.ui_train.get(u, [])) + ([want.ui_test[u][-1]] if u in want.ui_test else []) for u in users}
    rows = []
    depth = max(len(v) for v in per_user.values())
    for k in range(depth):                       # interleave users; time = position inside the user
        for u in users[::-1]:
            if k < len(per_user[u]):
                rows.append((7 * u + 3, 5 * per_user[u][k] + 11, k))
    raw_u, raw_i, t = (np.asarray(c, dtype=np.int64) for c in zip(*rows))
    res = eng.prep_filter_reindex(raw_u, raw_i)
    assert res["n_users"] == len(users) and res["n_items"] == len({i for v in per_user.values() for i in v})
    assert np.array_equal(res["user_ids"].cpu().numpy(), 7 * np.asarray(users) + 3)
    time = torch.from_numpy(t).cuda()[res["row"]]
    perm, is_test = eng.prep_split_loo(res["u"], res["n_users"], time)
    u, i = res["u"][perm].cpu().numpy(), res["i"][perm].cpu().numpy()
    is_test = is_test.cpu().numpy()
    # item ids: the golden split numbers items itself; map through the raw ids
    item_of = {k: (int(r) - 11) // 5 for k, r in enumerate(res["item_ids"].cpu().numpy().tolist())}
    user_of = {k: (int(r) - 3) // 7 for k, r in enumerate(res["user_ids"].cpu().numpy().tolist())}
    got_train, got_test = {}, {}
    for uu, ii, tt in zip(u.tolist(), i.tolist(), is_test.tolist()):
        (got_test if tt else got_train).setdefault(user_of[uu], []).append(item_of[ii])
    assert got_train == {k: list(v) for k, v in want.ui_train.items()}
    assert got_test == {k: [v[-1]] for k, v in want.ui_test.items()}


@pytest.mark.parametrize("user_min,item_min,by_time", [(0, 0, False), (5, 0, True), (3, 4, True), (0, 6, False)])
def test_device_pipeline_equals_host_mirror(eng, tmp_path, user_min, item_min, by_time):
    """data.preprocess=device == the packaged host class on a synthetic UIRT file: counts, training lists, test positives; filters
    on; negative and sparse raw ids; repeated (user, time) pairs (ties keep file order)."""
    from cleverrec_b200.model.RankingPreprocess import DeviceRankingPreprocess, RankingPreprocess
    rs = np.random.RandomState(user_min * 10 + item_min)
    n = 6000
    # a core of 200 users x 150 items plus long tails of rare users / items for the filters to remove; sparse raw ids, some negative
    raw_u = np.where(rs.rand(n) < 0.85, rs.randint(0, 200, n), 200 + rs.randint(0, 3000, n)) * 13 - 50
    raw_i = np.where(rs.rand(n) < 0.85, rs.randint(0, 150, n), 150 + rs.randint(0, 5000, n)) * 7 + 1
    t = rs.randint(0, 50, n)
    d = tmp_path / "syn"
    d.mkdir()
    with open(d / "log.tsv", "w") as f:
        f.write("u\ti\tr\tt\n")
        for a, b, c in zip(raw_u, raw_i, t):
            f.write("%d\t%d\t1\t%d\n" % (a, b, c))
    cfg = {"data.root_dir": str(tmp_path), "data.dataset": "syn", "data.file_name": "log.tsv", "data.sep": "\t", "data.format": "UIRT",
           "data.user_min": str(user_min), "data.item_min": str(item_min), "data.split_way": "loo", "data.split_by_time": str(by_time),
           "data.split_ratio": "[0.8,0.0,0.2]", "test.neg_samples": "20", "recommender": "BPR", "seed": "3"}
    np.random.seed(1)
    host = RankingPreprocess(dict(cfg), logging.getLogger("h"))
    dev = DeviceRankingPreprocess(dict(cfg), logging.getLogger("d"), engine=eng)
    assert (dev.user_nums, dev.item_nums) == (host.user_nums, host.item_nums)
    # the host class numbers ids in set-iteration order, the device class in ascending raw order: compare through the raw ids
    host_raw = _raw_maps(cfg, host)
    dev_res = eng.prep_filter_reindex(*_columns(cfg), user_min, item_min)
    dev_raw = ({k: int(v) for k, v in enumerate(dev_res["user_ids"].cpu().numpy().tolist())}, {k: int(v) for k, v in enumerate(dev_res["item_ids"].cpu().numpy().tolist())})
    def to_raw(d_, maps, strip):
        return {maps[0][u]: [maps[1][i] for i in (v[strip:] if strip else v)] for u, v in d_.items()}
    assert to_raw(dict(dev.ui_train.items()), dev_raw, 0) == to_raw(host.ui_train, host_raw, 0)
    assert to_raw(dev.ui_test, dev_raw, 20) == to_raw(host.ui_test, host_raw, 20)
    # evaluation negatives: the twin's, bit for bit; distinct, unseen, in range
    _, _, rp, sc = X.build_history({u: dev.ui_train[u] for u in dev.ui_train}, dev.user_nums)
    users = np.fromiter(dev.ui_test.keys(), dtype=np.int32)
    want = X.sample_eval_negatives(3, users, 20, dev.item_nums, rp, sc)
    got = np.asarray([dev.ui_test[int(u)][:20] for u in users], dtype=np.int32)
    assert np.array_equal(got, want)
    for k, u in enumerate(users.tolist()):
        assert len(set(got[k].tolist())) == 20 and not set(got[k].tolist()) & set(dev.ui_train[u] if u in dev.ui_train else [])


def _columns(cfg):
    import pandas as pd
    f = pd.read_csv(os.path.join(cfg["data.root_dir"], cfg["data.dataset"], cfg["data.file_name"]), sep="\t", header=0, names=["u", "i", "r", "t"])
    return f["u"].to_numpy(), f["i"].to_numpy()


def _raw_maps(cfg, host):
    """new id -> raw id of the host class (it does not keep the maps: rebuild them the way it numbers, utils/tools.py:9-15)."""
    import pandas as pd
    f = pd.read_csv(os.path.join(cfg["data.root_dir"], cfg["data.dataset"], cfg["data.file_name"]), sep="\t", header=0, names=["u_id", "i_id", "r", "t"])
    um, im = int(cfg["data.user_min"]), int(cfg["data.item_min"])
    if um > 0:
        f = f[f["u_id"].map(f["u_id"].value_counts()) >= um].reset_index(drop=True)
    if im > 0:
        f = f[f["i_id"].map(f["i_id"].value_counts()) >= im].reset_index(drop=True)
    return ({k: int(v) for k, v in enumerate(set(f["u_id"].unique()))}, {k: int(v) for k, v in enumerate(set(f["i_id"].unique()))})


def test_eval_negatives_large_request(eng):
    """1000 negatives per user out of a 3000-item catalogue with long histories: distinct, unseen, equal to the twin."""
    rs = np.random.RandomState(0)
    U, I = 40, 3000
    ui = {u: rs.choice(I, size=int(rs.randint(1, 1500)), replace=False).tolist() for u in range(U) if u % 7 != 3}
    eng.set_history(ui, U, I)
    _, _, rp, sc = X.build_history(ui, U)
    users = np.arange(U, dtype=np.int32)
    got = eng.prep_eval_negatives(11, users, 1000).cpu().numpy()
    want = X.sample_eval_negatives(11, users, 1000, I, rp, sc)
    assert np.array_equal(got, want)
    for u in range(U):
        assert len(set(got[u].tolist())) == 1000 and not set(got[u].tolist()) & set(ui.get(u, []))
