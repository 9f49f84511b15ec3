"""-m gpu: the CUDA sampler (csrc/sampler.cu, through the C ABI) against its integer-exact CPU twin."""
import numpy as np
import pytest

from conftest import synthetic_data
from oracle import philox as X

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cleverrec_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _hist(eng, d):
    h = X.build_history(d.ui_train, d.user_nums)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    return h


@pytest.mark.parametrize("shape", [(80, 300, 20, 4), (7, 33, 3, 1), (300, 5000, 40, 8), (50, 64, 30, 3)])
def test_pairwise_bit_exact(eng, shape):
    U, I, L, R = shape
    d = synthetic_data(U, I, L, seed=U)
    pu, pi, rp, sc = _hist(eng, d)
    n = pu.shape[0] * R
    assert eng.epoch_rows(R, "pairwise") == n
    for seed, epoch in ((0, 0), (0xFEEDFACECAFE, 3)):
        u, i, j, nbr = eng.sample_pairwise(seed, epoch, 0, n, R, with_nbr=True)
        ru, ri, rj, rn = X.sample_pairwise(seed, epoch, 0, n, R, d.item_nums, pu, pi, rp, sc)
        assert np.array_equal(u.cpu().numpy(), ru) and np.array_equal(i.cpu().numpy(), ri)
        assert np.array_equal(j.cpu().numpy(), rj) and np.array_equal(nbr.cpu().numpy(), rn)
    # a window in the middle of the epoch (ragged last batch included)
    first = n // 3
    cnt = min(777, n - first)
    u, i, j = eng.sample_pairwise(5, 1, first, cnt, R)
    ru, ri, rj, _ = X.sample_pairwise(5, 1, first, cnt, R, d.item_nums, pu, pi, rp, sc)
    assert np.array_equal(j.cpu().numpy(), rj) and np.array_equal(u.cpu().numpy(), ru)


def test_pointwise_and_cml_bit_exact(eng):
    d = synthetic_data(90, 400, 15, seed=9)
    pu, pi, rp, sc = _hist(eng, d)
    R = 3
    n = eng.epoch_rows(R, "pointwise")
    assert n == pu.shape[0] * (R + 1)
    u, i, y = eng.sample_pointwise(11, 2, 0, n, R)
    ru, ri, ry, _ = X.sample_pointwise(11, 2, 0, n, R, d.item_nums, pu, pi, rp, sc)
    assert np.array_equal(u.cpu().numpy(), ru) and np.array_equal(i.cpu().numpy(), ri) and np.array_equal(y.cpu().numpy(), ry)
    R = 20  # conf/CML.properties neg_ratio
    n = eng.epoch_rows(R, "cml")
    u, i, neg = eng.sample_cml(4, 0, 0, n, R)
    ru, ri, rneg = X.sample_cml(4, 0, 0, n, R, d.item_nums, pu, pi, rp, sc)
    assert np.array_equal(u.cpu().numpy(), ru) and np.array_equal(neg.cpu().numpy(), rneg)


def test_duplicate_interactions_and_empty_users(eng):
    d = synthetic_data(30, 100, 6, seed=2)
    d.ui_train[0] = d.ui_train[0] + d.ui_train[0]  # duplicated rows stay separate positives (utils/sampler.py:52)
    pu, pi, rp, sc = _hist(eng, d)
    n = pu.shape[0] * 2
    u, i, j = eng.sample_pairwise(1, 0, 0, n, 2)
    ru, ri, rj, _ = X.sample_pairwise(1, 0, 0, n, 2, d.item_nums, pu, pi, rp, sc)
    assert np.array_equal(j.cpu().numpy(), rj)
    assert eng.sample_pairwise(1, 0, 0, 0, 2)[0].numel() == 0  # empty request


def test_catalogue_exhausted_is_an_error(eng):
    from cleverrec_b200._lib import CrbError
    ui = {0: list(range(10)), 1: [0]}  # user 0 has seen every item: the reference would loop forever
    eng.set_history(ui, 2, 10)
    with pytest.raises(CrbError) as e:
        eng.sample_pairwise(0, 0, 0, 11, 1)
    assert e.value.code == -4


@pytest.mark.parametrize("shape", [(60, 40, 18, 3), (200, 3000, 2, 2), (50, 20000, 600, 4)])
def test_bloom_filter_never_changes_the_output(eng, shape, monkeypatch):
    """The per-user seen-item Bloom filters (csrc/api.cu::crb_bloom_build) only short-cut the rejection test of
    utils/sampler.py:58-59: with the filters disabled (CRB_NO_BLOOM, read when the history is installed) the sampler takes the exact
    binary search for every candidate and must return the same triplets -- including saturated filters (histories much longer than
    the 2048-bit cap) and nearly empty ones (32 bits per user)."""
    U, I, L, R = shape
    d = synthetic_data(U, I, L, seed=U + 1)
    pu, pi, rp, sc = X.build_history(d.ui_train, d.user_nums)
    n = pu.shape[0] * R
    outs = []
    # three ways to test a candidate: exact seen-item bitmap (small catalogues, the default here), hashed filter + search, search only
    for env in ({}, {"CRB_NO_EXACT_BITMAP": "1"}, {"CRB_NO_BLOOM": "1"}):
        monkeypatch.delenv("CRB_NO_BLOOM", raising=False)
        monkeypatch.delenv("CRB_NO_EXACT_BITMAP", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng.set_history(d.ui_train, d.user_nums, d.item_nums)
        outs.append([t.cpu().numpy() for t in eng.sample_pairwise(0xABCDEF, 2, 0, n, R)])
        outs[-1] += [t.cpu().numpy() for t in eng.sample_cml(5, 1, 0, pu.shape[0], R)]
    monkeypatch.delenv("CRB_NO_BLOOM", raising=False)
    monkeypatch.delenv("CRB_NO_EXACT_BITMAP", raising=False)
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)
    ru, ri, rj, _ = X.sample_pairwise(0xABCDEF, 2, 0, n, R, d.item_nums, pu, pi, rp, sc)
    assert np.array_equal(outs[0][2], rj) and np.array_equal(outs[0][0], ru)
