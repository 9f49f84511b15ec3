"""-m gpu: crb_build_history (csrc/history.cu) against the dict / list / set forms the reference builds
(model/RankingPreprocess.py:117 `groupby('u_id').i_id.apply(list).to_dict()`, utils/sampler.py:53 `set(ui_train[u])`)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cleverrec_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def reference_forms(users, items, n_users):
    """What the reference holds: dict user -> list of items in row order, users in ascending id order (pandas groupby sorts)."""
    ui = {}
    for u in np.unique(users):
        ui[int(u)] = items[users == u].tolist()
    return ui


@pytest.mark.parametrize("n,n_users,n_items,device_in", [(5000, 300, 70, False), (5000, 300, 70, True), (1, 5, 5, False), (200000, 1000, 50000, True)])
def test_matches_dict_forms(eng, n, n_users, n_items, device_in):
    from cleverrec_b200.engine import history_from_dict
    rs = np.random.RandomState(n % 97)
    users = rs.randint(0, n_users, n).astype(np.int32)
    users[users % 7 == 3] = 0                      # some users never occur, user 0 is a hub
    items = rs.randint(0, n_items, n).astype(np.int32)   # duplicates (same user, same item) occur
    ui = reference_forms(users, items, n_users)
    want_pu, want_pi, want_rp, want_sc = history_from_dict(ui, n_users)
    a, b = (torch.from_numpy(users).cuda(), torch.from_numpy(items).cuda()) if device_in else (users, items)
    pu, pi, rp, sc = eng.build_history(a, b, n_users, n_items)
    assert np.array_equal(pu.cpu().numpy(), want_pu) and np.array_equal(pi.cpu().numpy(), want_pi)
    assert np.array_equal(rp.cpu().numpy(), want_rp)
    assert np.array_equal(sc.cpu().numpy()[: want_sc.shape[0]], want_sc) and int(rp[-1]) == want_sc.shape[0]
    start, ln = (t.cpu().numpy() for t in eng._lists)
    for u in range(n_users):
        assert ln[u] == len(ui.get(u, []))
        if ln[u]:
            assert want_pi[start[u]:start[u] + ln[u]].tolist() == ui[u]


def test_sampler_sees_the_same_history(eng):
    """Same triplets whether the history came from the dict path or the native builder."""
    rs = np.random.RandomState(3)
    n, U, I = 4000, 120, 400
    users, items = rs.randint(0, U, n).astype(np.int32), rs.randint(0, I, n).astype(np.int32)
    ui = reference_forms(users, items, U)
    eng.set_history(ui, U, I)
    rows = eng.epoch_rows(3)
    a = [t.cpu().numpy() for t in eng.sample_pairwise(5, 1, 0, rows, 3)]
    eng.build_history(users, items, U, I)
    assert eng.epoch_rows(3) == rows
    b = [t.cpu().numpy() for t in eng.sample_pairwise(5, 1, 0, rows, 3)]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_empty_and_bad_ids(eng):
    from cleverrec_b200._lib import CrbError
    z = np.zeros(0, dtype=np.int32)
    pu, pi, rp, sc = eng.build_history(z, z, 4, 4)
    assert pu.numel() == 0 and rp.cpu().tolist() == [0, 0, 0, 0, 0]
    with pytest.raises(CrbError):
        eng.build_history(np.array([0, 9], dtype=np.int32), np.array([1, 1], dtype=np.int32), 4, 4)
    with pytest.raises(CrbError):
        eng.build_history(np.array([0, 1], dtype=np.int32), np.array([1, -1], dtype=np.int32), 4, 4)
