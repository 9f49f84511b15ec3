"""-m gpu: the `numpy_stream` sampler mode (csrc/sampler_np.cu) reproduces the GENUINE reference samplers bit for bit: same
arrays, same consumption of NumPy's global MT19937 stream (golden vectors made by the reference itself, oracle/make_golden.py)."""
import numpy as np
import pytest

from conftest import Data, synthetic_data
from oracle import ref_host as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cleverrec_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def test_golden_reference_stream(eng, split_loo, golden):
    g = golden["sampler_seed3"]
    users = list(split_loo.ui_train.keys())[:int(g["n_sub_users"])]
    sub = {u: split_loo.ui_train[u] for u in users}
    eng.set_history(sub, split_loo.user_nums, split_loo.item_nums)
    eng.np_seed(3)  # np.random.seed(3); the reference then ran the three samplers back to back on the same stream
    u, i, j, nbr = eng.sample_epoch_numpy("pairwise", 4, with_nbr=True)
    for got, name in ((u, "pw_u"), (i, "pw_i"), (j, "pw_j"), (nbr, "pw_nbr")):
        assert np.array_equal(got.cpu().numpy(), g[name]), name
    u, i, y = eng.sample_epoch_numpy("pointwise", 2)
    assert np.array_equal(u.cpu().numpy(), g["pt_u"]) and np.array_equal(i.cpu().numpy(), g["pt_i"])
    assert np.array_equal(y.cpu().numpy(), g["pt_y"])
    u, i, neg = eng.sample_epoch_numpy("cml", 5)
    assert np.array_equal(u.cpu().numpy(), g["cm_u"]) and np.array_equal(i.cpu().numpy(), g["cm_i"])
    assert np.array_equal(neg.cpu().numpy(), g["cm_neg"])


@pytest.mark.parametrize("shape", [(40, 64, 30, 3), (300, 2000, 25, 4), (5, 60, 2, 2), (12, 40, 1, 1)])
def test_matches_restated_reference_and_hands_the_stream_back(eng, shape):
    U, I, L, R = shape
    d = synthetic_data(U, I, L, seed=U + 1)
    if U == 40:
        d.ui_train[0] = d.ui_train[0] + d.ui_train[0][:4]   # duplicated interactions stay separate positives
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    np.random.seed(1234)
    eng.np_set_state()                      # adopt NumPy's global state ...
    want = H.pairwise_ranking_sampler(d, R, 512, fism_like=True)
    after = np.random.get_state()
    got = eng.sample_epoch_numpy("pairwise", R, with_nbr=True)
    for a, b in zip(got, want[1:]):
        assert np.array_equal(a.cpu().numpy(), np.asarray(b)), shape
    st = eng.np_get_state()                 # ... and hand it back exactly where the reference would have left it
    assert st[2] == after[2] and np.array_equal(st[1], after[1])
    # a second epoch continues the same stream
    want2 = H.ranking_sampler_cml(d, R, 512)
    got2 = eng.sample_epoch_numpy("cml", R)
    assert np.array_equal(got2[2].cpu().numpy(), want2[3])
    # negatives in draw order (train_model_nais)
    np.random.seed(7)
    eng.np_seed(7)
    negs = eng.sample_epoch_numpy("negatives", R).cpu().numpy()
    k = 0
    for u, hist, i_idx, y in H.nais_user_batches(d, R):
        for p in range(len(hist)):
            assert i_idx[p * (R + 1) + 1:(p + 1) * (R + 1)] == negs[k].tolist()
            k += 1


def test_unseeded_is_an_error(eng):
    from cleverrec_b200._lib import CrbError
    from cleverrec_b200.engine import Engine
    e2 = Engine(0)
    e2.set_history({0: [1, 2]}, 1, 10)
    with pytest.raises(CrbError) as err:
        e2.sample_epoch_numpy("pairwise", 2)
    assert err.value.code == -3
    e2.close()


def test_drop_in_sampler_functions_return_the_reference_arrays(split_loo, golden):
    """cleverrec_b200.utils.sampler in numpy_stream mode: same call, same seed, same tuple as the reference's utils/sampler.py."""
    from cleverrec_b200.utils import sampler as S
    g = golden["sampler_seed3"]
    users = list(split_loo.ui_train.keys())[:int(g["n_sub_users"])]
    sub = Data(split_loo.user_nums, split_loo.item_nums, {u: split_loo.ui_train[u] for u in users}, {})
    S.set_mode("numpy_stream")
    try:
        np.random.seed(3)
        pw = S.pairwise_ranking_sampler(sub, 4, 6144, fism_like=True)
        pt = S.pointwise_ranking_sampler(sub, 2, 1000)
        cm = S.ranking_sampler_cml(sub, 5, 512)
    finally:
        S.set_mode("philox")
    assert pw[0] == int(g["pw_batches"]) and pt[0] == int(g["pt_batches"]) and cm[0] == int(g["cm_batches"])
    assert pw[1].dtype == np.int64 and np.array_equal(pw[3], g["pw_j"]) and np.array_equal(pw[4], g["pw_nbr"])
    assert np.array_equal(pt[2], g["pt_i"]) and pt[3].dtype == np.float64 and np.array_equal(pt[3], g["pt_y"])
    assert np.array_equal(cm[3], g["cm_neg"])
