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
