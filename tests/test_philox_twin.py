"""The integer-exact CPU twin of the device sampler (oracle/philox.py): known answers, bijection, the sampler
invariants of utils/sampler.py:58-61, and its distribution against the reference sampler's."""
import numpy as np

from conftest import synthetic_data
from oracle import philox as X
from oracle import ref_host as H


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        got = X.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == out


def test_feistel_is_a_bijection():
    for n in (1, 2, 3, 17, 256, 257, 5000, 65537):
        keys = X.perm_keys(0xDEADBEEF12345, 9)
        p = X.feistel_perm(np.arange(n), n, keys)
        assert np.array_equal(np.sort(p), np.arange(n, dtype=np.uint64))
    a = X.feistel_perm(np.arange(1000), 1000, X.perm_keys(1, 0))
    b = X.feistel_perm(np.arange(1000), 1000, X.perm_keys(1, 1))
    assert not np.array_equal(a, b)  # a new shuffle every epoch


def test_sampler_invariants_and_coverage():
    d = synthetic_data(80, 300, 20, seed=1)
    pu, pi, rp, sc = X.build_history(d.ui_train, d.user_nums)
    R = 4
    n = pu.shape[0] * R
    u, i, j, nbr = X.sample_pairwise(77, 2, 0, n, R, d.item_nums, pu, pi, rp, sc)
    groups = {}
    for a, b, c, nb in zip(u.tolist(), i.tolist(), j.tolist(), nbr.tolist()):
        assert c not in d.ui_train[a] and 0 <= c < d.item_nums
        assert nb == len(set(d.ui_train[a]))
        groups.setdefault((a, b), []).append(c)
    # every positive appears exactly neg_ratio times with distinct negatives (utils/sampler.py:52-61)
    assert len(groups) == pu.shape[0]
    assert all(len(v) == R and len(set(v)) == R for v in groups.values())
    # chunked sampling == one-shot sampling (rows are a pure function of their epoch position)
    u2, i2, j2, _ = X.sample_pairwise(77, 2, 1000, 500, R, d.item_nums, pu, pi, rp, sc)
    assert np.array_equal(u2, u[1000:1500]) and np.array_equal(j2, j[1000:1500])


def test_pointwise_and_cml_twins():
    d = synthetic_data(40, 200, 10, seed=2)
    pu, pi, rp, sc = X.build_history(d.ui_train, d.user_nums)
    u, i, y, _ = X.sample_pointwise(5, 0, 0, pu.shape[0] * 4, 3, d.item_nums, pu, pi, rp, sc)
    assert int(y.sum()) == pu.shape[0]
    for a, b, lab in zip(u.tolist(), i.tolist(), y.tolist()):
        assert (b in d.ui_train[a]) == (lab == 1.0)
    u, i, neg = X.sample_cml(5, 0, 0, pu.shape[0], 6, d.item_nums, pu, pi, rp, sc)
    for a, b, row in zip(u.tolist(), i.tolist(), neg.tolist()):
        assert b in d.ui_train[a] and len(set(row)) == 6 and not (set(row) & set(d.ui_train[a]))


def test_negative_distribution_matches_reference_sampler():
    # same marginal law: uniform over the user's unseen items.  Compare per-item negative counts of one user.
    d = synthetic_data(3, 40, 8, seed=3)
    pu, pi, rp, sc = X.build_history(d.ui_train, d.user_nums)
    R = 2
    cnt_twin = np.zeros(d.item_nums)
    cnt_ref = np.zeros(d.item_nums)
    np.random.seed(0)
    for e in range(300):
        u, i, j, _ = X.sample_pairwise(9, e, 0, pu.shape[0] * R, R, d.item_nums, pu, pi, rp, sc)
        np.add.at(cnt_twin, j[u == 0], 1)
        out = H.pairwise_ranking_sampler(d, R, 64)
        np.add.at(cnt_ref, out[3][out[1] == 0], 1)
    unseen = np.setdiff1d(np.arange(d.item_nums), d.ui_train[0])
    assert cnt_twin[d.ui_train[0]].sum() == 0 and cnt_ref[d.ui_train[0]].sum() == 0
    pt, pr = cnt_twin[unseen] / cnt_twin.sum(), cnt_ref[unseen] / cnt_ref.sum()
    # both within 5 sigma of uniform
    n = cnt_twin.sum()
    sigma = np.sqrt((1 / len(unseen)) * (1 - 1 / len(unseen)) / n)
    assert np.abs(pt - 1 / len(unseen)).max() < 5 * sigma
    assert np.abs(pr - 1 / len(unseen)).max() < 5 * sigma


def test_eval_negative_twin_draws_distinct_unseen_items_uniformly():
    """oracle/philox.py::sample_eval_negatives (the twin of crb_prep_eval_negatives) has the law of the reference's
    np.random.choice(list(item_set - seen), size, replace=False) (RankingPreprocess.py:120-129): distinct, unseen, in range, a pure
    function of (seed, user, history), and every unseen item equally likely (chi-square over many seeds)."""
    rs = np.random.RandomState(0)
    I, U = 61, 5
    ui = {u: rs.choice(I, size=int(rs.randint(3, 30)), replace=False).tolist() for u in range(U)}
    _, _, rp, sc = X.build_history(ui, U)
    users = np.arange(U, dtype=np.int32)
    a = X.sample_eval_negatives(7, users, 10, I, rp, sc)
    assert np.array_equal(a, X.sample_eval_negatives(7, users, 10, I, rp, sc))
    assert not np.array_equal(a, X.sample_eval_negatives(8, users, 10, I, rp, sc))
    counts = np.zeros((U, I))
    n_seeds = 400
    for seed in range(n_seeds):
        out = X.sample_eval_negatives(seed, users, 10, I, rp, sc)
        for u in range(U):
            row = out[u].tolist()
            assert len(set(row)) == 10 and not set(row) & set(ui[u]) and min(row) >= 0 and max(row) < I
            counts[u, row] += 1
    for u in range(U):
        unseen = np.setdiff1d(np.arange(I), ui[u])
        expected = n_seeds * 10 / unseen.size
        chi2 = float(((counts[u, unseen] - expected) ** 2 / expected).sum())
        assert chi2 < 2.2 * unseen.size, (u, chi2, unseen.size)     # far above the 99.9 % quantile would mean a biased sampler
        assert counts[u, ui[u]].sum() == 0
