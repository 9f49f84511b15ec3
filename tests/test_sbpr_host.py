"""SBPR host side, no GPU: get_SPu and ranking_sampler_sbpr restated in oracle/ref_host.py and the packaged get_SPu
(cleverrec_b200/utils/tools.py) against golden vectors made by the GENUINE reference functions on dataset/Ciao
(oracle/make_golden.py section F); the packaged RankingPreprocess's `social_file` branch against the genuine class
(build container only: needs the dataset mount); the flat social arrays against the sampler's own loops."""
import logging
import os

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE_ROOT, Data, unflat
from oracle import philox as X
from oracle import ref_host as H

DATA = os.path.join(REFERENCE_ROOT, "dataset")
needs_data = pytest.mark.skipif(not os.path.exists(os.path.join(DATA, "Ciao", "trusts.csv")), reason="reference dataset mount absent")


def _fixture():
    z = np.load(os.path.join(GOLDEN, "sbpr_ciao.npz"))
    d = Data(int(z["user_nums"]), int(z["item_nums"]), unflat(z["train_keys"], z["train_lens"], z["train_items"]), {})
    d.user_friends = unflat(z["friends_keys"], z["friends_lens"], z["friends_items"])
    return z, d, unflat(z["spu_keys"], z["spu_lens"], z["spu_items"])


def test_get_spu_matches_the_reference_including_list_order():
    from cleverrec_b200.utils.tools import get_SPu
    z, d, want = _fixture()
    for fn in (H.get_SPu, get_SPu):
        got = fn(d)
        assert list(got.keys()) == list(want.keys())
        assert all(got[u] == want[u] for u in want)      # same order: the sampler indexes the list by position


def test_restated_sbpr_sampler_follows_the_reference_stream():
    z, d, SPu = _fixture()
    np.random.seed(3)
    out = H.ranking_sampler_sbpr(d, SPu, 2, 4096)
    assert out[0] == int(z["sb_batches"])
    for got, key in zip(out[1:], ("sb_u", "sb_i", "sb_k", "sb_j", "sb_suk")):
        assert np.array_equal(got, z[key]), key
    # invariants the device sampler is tested against as well
    u, i, k, j = out[1:5]
    for t in range(0, len(u), 97):
        assert k[t] in SPu[u[t]] and j[t] not in d.ui_train[u[t]] and j[t] not in SPu[u[t]] and i[t] in d.ui_train[u[t]]


def test_social_arrays_restate_the_sampler_loops():
    from cleverrec_b200.engine import social_arrays
    z, d, SPu = _fixture()
    pu, pi, start, items, suk, rp, cols = social_arrays(d.ui_train, d.user_friends, SPu, d.user_nums)
    opu, opi, ospu, osuk, oexcl = X.social_history(d.ui_train, d.user_friends, SPu, d.user_nums)
    assert np.array_equal(pu, opu) and np.array_equal(pi, opi)
    assert len(pu) * 2 == len(z["sb_u"])                                   # epoch rows = social positives x neg_ratio
    for u in range(d.user_nums):
        a, b = int(start[u]), int(start[u + 1])
        if u in SPu:
            assert items[a:b].tolist() == ospu[u] and suk[a:b].tolist() == [float(x) for x in osuk[u]]
            assert cols[rp[u]:rp[u + 1]].tolist() == oexcl[u]
            assert min(osuk[u]) >= 1                                        # every social item was consumed by at least one friend
        else:
            assert a == b and rp[u] == rp[u + 1]
    # the golden sampler output's suk values are the table's values for the (u, k) it drew
    lut = {(u, k): s for u in ospu for k, s in zip(ospu[u], osuk[u])}
    assert all(lut[(int(u), int(k))] == int(s) for u, k, s in zip(z["sb_u"][::53], z["sb_k"][::53], z["sb_suk"][::53]))


def test_device_sampler_twin_has_the_reference_law():
    """oracle/philox.py::sample_sbpr (the bit-exact twin of the CUDA sampler) draws k uniformly from SPu[u] and j uniformly from the
    admissible items, like utils/sampler.py:113-121: chi-square of both marginals for one user over many epochs."""
    z, d, SPu = _fixture()
    u0 = min(SPu, key=lambda u: len(SPu[u]) if len(SPu[u]) >= 8 else 10 ** 9)
    one = Data(d.user_nums, 64, {u0: [i % 64 for i in d.ui_train[u0]][:5]}, {})
    one.user_friends = {u0: [u0 + 1]}
    one.ui_train[u0 + 1] = [40, 41, 42, 43, 44, 45, 46, 47]
    sp = H.get_SPu(one)
    social = X.social_history(one.ui_train, one.user_friends, sp, one.user_nums)
    n = social[0].shape[0] * 4
    ks, js = [], []
    for epoch in range(150):
        u, i, k, j, suk = X.sample_sbpr(9, epoch, 0, n, 4, 64, social)
        keep = u == u0
        ks.append(k[keep]); js.append(j[keep])
    ks, js = np.concatenate(ks), np.concatenate(js)
    assert set(ks.tolist()) == set(sp[u0])
    ck = np.bincount(ks, minlength=64)[sp[u0]]
    assert ((ck - ck.mean()) ** 2 / ck.mean()).sum() < 3 * len(ck)
    allowed = sorted(set(range(64)) - set(one.ui_train[u0]) - set(sp[u0]))
    cj = np.bincount(js, minlength=64)
    assert cj.sum() == cj[allowed].sum()
    assert ((cj[allowed] - cj[allowed].mean()) ** 2 / cj[allowed].mean()).sum() < 2 * len(allowed)


@needs_data
def test_preprocess_social_file_matches_the_reference_class():
    from cleverrec_b200.model.RankingPreprocess import RankingPreprocess
    from oracle import refimport as R
    cfg = R.default_configs(recommender="SBPR", **{"data.dataset": "Ciao", "data.file_name": "ratings.csv", "data.sep": ",", "data.format": "UI",
                                                   "data.split_way": "rs", "test.neg_samples": 0})
    np.random.seed(5)
    want = R.preprocess(cfg)
    np.random.seed(5)
    got = RankingPreprocess(dict(cfg, **{"data.root_dir": DATA}), logging.getLogger("t"))
    assert got.user_nums == want.user_nums and got.item_nums == want.item_nums
    assert list(got.user_friends.keys()) == list(want.user_friends.keys())
    assert all(got.user_friends[u] == want.user_friends[u] for u in want.user_friends)
    assert list(got.ui_train.keys()) == list(want.ui_train.keys()) and all(got.ui_train[u] == want.ui_train[u] for u in want.ui_train)
    z = np.load(os.path.join(GOLDEN, "sbpr_ciao.npz"))
    assert len(got.user_friends) == int(z["n_friend_users"]) and sum(len(v) for v in got.user_friends.values()) == int(z["n_friend_pairs"])
