"""-m gpu: evaluation kernels against the C oracle (oracle/crb_oracle.c): scores, ranks and top-K ids bit-exact."""
import numpy as np
import pytest
import torch

from conftest import synthetic_data
from oracle import c_oracle as O
from oracle import philox as X

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cleverrec_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("d", [7, 16, 64, 128])
def test_score_pairs_bit_exact(eng, kind, d):
    rs = np.random.RandomState(d + kind)
    U, I, n = 50, 70, 5000
    P, Q = rs.randn(U, d).astype(np.float32), rs.randn(I, d).astype(np.float32)
    hvec = rs.randn(d if kind == 1 else I).astype(np.float32) if kind in (1, 3) else None
    u, i = rs.randint(0, U, n), rs.randint(0, I, n)
    want = O.score_pairs(kind, P, Q, u, i, hvec)
    Pd, Qd = torch.tensor(P).cuda(), torch.tensor(Q).cuda()
    hd = torch.tensor(hvec).cuda() if hvec is not None else None
    got_host = eng.score_pairs(kind, Pd, Qd, u, i, hvec=hd)  # host feed -> host scores (sess.run style)
    got_dev = eng.score_pairs(kind, Pd, Qd, torch.tensor(u, dtype=torch.int32).cuda(), torch.tensor(i, dtype=torch.int32).cuda(), hvec=hd)
    assert np.array_equal(got_host.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got_dev.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("ascending", [False, True])
def test_topk_segments_with_ties_and_ragged(eng, ascending):
    rs = np.random.RandomState(1)
    lens = np.array([100, 1, 0, 1001, 19, 20, 21, 333])
    offsets = np.concatenate([[0], np.cumsum(lens)])
    scores = np.round(rs.randn(offsets[-1]) * 4).astype(np.float32) / 4  # heavy ties
    scores[5] = -0.0
    want = O.topk_segments(scores, offsets, 20, ascending)
    got = eng.topk_segments(scores, offsets, 20, ascending)
    assert np.array_equal(got, want)
    got_d = eng.topk_segments(torch.tensor(scores).cuda(), offsets, 20, ascending).cpu().numpy()
    assert np.array_equal(got_d, want)


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("exact", [True, False])
def test_fullrank_topk_bit_exact(eng, kind, exact):
    d = synthetic_data(150, 3000, 60, seed=kind)
    pu, pi, rp, sc = X.build_history(d.ui_train, d.user_nums)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    rs = np.random.RandomState(kind)
    dim = 64
    P, Q = (rs.randn(d.user_nums, dim) * 0.1).astype(np.float32), (rs.randn(d.item_nums, dim) * 0.1).astype(np.float32)
    Q[100] = Q[200]  # exact score ties between two items
    hvec = (rs.randn(dim if kind == 1 else d.item_nums) * 0.1).astype(np.float32) if kind in (1, 3) else None
    users = np.arange(d.user_nums, dtype=np.int32)  # includes users with no history (u % 11 == 7)
    want_i, want_s = O.fullrank_topk(kind, P, Q, users, rp, sc, 20, hvec)
    Pd, Qd = torch.tensor(P).cuda(), torch.tensor(Q).cuda()
    hd = torch.tensor(hvec).cuda() if hvec is not None else None
    got_i, got_s = eng.score_topk(kind, Pd, Qd, users, 20, hvec=hd, exact=exact, return_scores=True)
    assert np.array_equal(got_i, want_i)
    assert np.array_equal(got_s.view(np.uint32), want_s.view(np.uint32))


def test_fullrank_small_catalogue_pads(eng):
    ui = {0: [0, 1, 2], 1: [3]}
    eng.set_history(ui, 3, 8)
    rs = np.random.RandomState(0)
    P, Q = rs.randn(3, 16).astype(np.float32), rs.randn(8, 16).astype(np.float32)
    pu, pi, rp, sc = X.build_history(ui, 3)
    want, _ = O.fullrank_topk(0, P, Q, np.arange(3), rp, sc, 10)
    got = eng.score_topk(0, torch.tensor(P).cuda(), torch.tensor(Q).cuda(), np.arange(3, dtype=np.int32), 10, exact=True)
    assert np.array_equal(got, want) and (got[0, 5:] == -1).all()


@pytest.mark.parametrize("kind", [0, 2])
def test_fullrank_exact_few_users_chunked_catalogue(eng, kind):
    """Few users against a catalogue large enough to be cut into ranges (fullrank_exact_kernel n_chunks > 1 + merge): the path the
    tensor-core kernel's uncertified users take.  Same ids and score bits as the oracle's single sweep; heavy ties included."""
    d = synthetic_data(7, 40000, 50, seed=10 + kind)
    pu, pi, rp, sc = X.build_history(d.ui_train, d.user_nums)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    rs = np.random.RandomState(kind)
    dim = 32
    P, Q = (rs.randn(d.user_nums, dim) * 0.1).astype(np.float32), (rs.randn(d.item_nums, dim) * 0.1).astype(np.float32)
    Q[5000:5040] = Q[100]     # 41-way exact tie that straddles range boundaries
    Q[39990:] = Q[100]
    users = np.array([6, 0, 3], dtype=np.int32)
    want_i, want_s = O.fullrank_topk(kind, P, Q, users, rp, sc, 50, None)
    got_i, got_s = eng.score_topk(kind, torch.tensor(P).cuda(), torch.tensor(Q).cuda(), users, 50, exact=True, return_scores=True)
    assert np.array_equal(got_i, want_i)
    assert np.array_equal(got_s.view(np.uint32), want_s.view(np.uint32))


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("d", [7, 64, 100, 128, 200])
def test_score_pairs_large_calls_bit_exact(eng, kind, d):
    """Calls of >= 8192 pairs take score_pairs_tiled_kernel (rows staged in shared memory by coalesced reads; each thread still runs the
    canonical sequential chain): bit-identical to the C oracle for user-uniform warps (test_model_loo's layout: all candidates of a user
    are contiguous), mixed warps, a ragged tail and dimensions that are not multiples of the 64-column chunk."""
    rs = np.random.RandomState(d * 7 + kind)
    U, I = 300, 500
    P, Q = rs.randn(U, d).astype(np.float32), rs.randn(I, d).astype(np.float32)
    hvec = rs.randn(d if kind == 1 else I).astype(np.float32) if kind in (1, 3) else None
    u = np.concatenate([np.repeat(rs.randint(0, U, 150), 101), rs.randint(0, U, 5003)])      # loo layout, then fully mixed users
    i = rs.randint(0, I, u.shape[0])
    assert u.shape[0] >= 8192 and u.shape[0] % 32 != 0
    want = O.score_pairs(kind, P, Q, u, i, hvec)
    hd = torch.tensor(hvec).cuda() if hvec is not None else None
    got = eng.score_pairs(kind, torch.tensor(P).cuda(), torch.tensor(Q).cuda(), torch.tensor(u, dtype=torch.int32).cuda(),
                          torch.tensor(i, dtype=torch.int32).cuda(), hvec=hd)
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("d", [8, 64, 100, 128, 256])
def test_score_pairs_topk_fused_bit_exact(eng, kind, d):
    """loo_topk_kernel (test_model_loo's predict + argsort fused; scores never reach HBM) == oracle scores + oracle top-K per segment:
    positions bit-exact, ragged / empty / shorter-than-K segments, exact score ties (duplicate candidates), K in {1, 20, 32}."""
    rs = np.random.RandomState(100 * kind + d)
    U, I = 40, 300
    P, Q = rs.randn(U, d).astype(np.float32), rs.randn(I, d).astype(np.float32)
    hvec = rs.randn(d if kind == 1 else I).astype(np.float32) if kind in (1, 3) else None
    lens = np.array([1001, 0, 1, 19, 20, 21, 33, 64, 100, 333, 7, 32], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)])
    seg_users = rs.randint(0, U, lens.shape[0]).astype(np.int32)
    items = rs.randint(0, I, offsets[-1]).astype(np.int32)        # 1001 draws over 300 items: many exact ties inside a segment
    u_flat = np.repeat(seg_users, lens)
    scores = O.score_pairs(kind, P, Q, u_flat, items, hvec)
    Pd, Qd = torch.tensor(P).cuda(), torch.tensor(Q).cuda()
    hd = torch.tensor(hvec).cuda() if hvec is not None else None
    for K in (1, 20, 32):
        for asc in (False, True):
            want = O.topk_segments(scores, offsets, K, asc)
            got_host = eng.score_pairs_topk(kind, Pd, Qd, seg_users, items, offsets, K, hvec=hd, ascending=asc)
            got_dev = eng.score_pairs_topk(kind, Pd, Qd, torch.tensor(seg_users).cuda(), torch.tensor(items).cuda(), torch.tensor(offsets).cuda(), K,
                                           hvec=hd, ascending=asc)
            assert np.array_equal(got_host, want), (K, asc)
            assert np.array_equal(got_dev.cpu().numpy(), want), (K, asc)
    want = O.topk_segments(scores, offsets, 50, False)            # K > 32: the two-kernel route behind the same call
    assert np.array_equal(eng.score_pairs_topk(kind, Pd, Qd, seg_users, items, offsets, 50, hvec=hd), want)


def test_fullrank_item_table_cache_tracks_writes(eng):
    """The tensor-core path keeps its bf16 copy of Q between calls (the reference's test.batch_size loop ranks users in many small
    calls): a second call with the same table skips the conversion, a training step through the library drops the copy, so does a torch
    in-place write (the engine watches the tensor's version counter; invalidate_eval_cache() is for writes torch cannot see).  Every result equals the oracle on the CURRENT table."""
    from cleverrec_b200.engine import Optimizer, Table
    d = synthetic_data(120, 2500, 40, seed=21)
    pu, pi, rp, sc = X.build_history(d.ui_train, d.user_nums)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    rs = np.random.RandomState(3)
    dim = 64
    P = Table(torch.tensor((rs.randn(d.user_nums, dim) * 0.1).astype(np.float32)).cuda(), "SGD")
    Q = Table(torch.tensor((rs.randn(d.item_nums, dim) * 0.1).astype(np.float32)).cuda(), "SGD")
    users = np.arange(d.user_nums, dtype=np.int32)

    def check_now():
        want, _ = O.fullrank_topk(0, P.w.cpu().numpy(), Q.w.cpu().numpy(), users, rp, sc, 20)
        l0 = eng.launches
        got = eng.score_topk(0, P.w, Q.w, users, 20)
        assert np.array_equal(got, want)
        return eng.launches - l0
    first = check_now()
    assert check_now() == first - 1                       # prep_kernel<Q> skipped
    opt = Optimizer("SGD", 0.5)
    u, i, j = rs.randint(0, d.user_nums, 4096), rs.randint(0, d.item_nums, 4096), rs.randint(0, d.item_nums, 4096)
    eng.train_step_bpr(P, Q, opt, u, i, j, 0.01)
    assert check_now() == first                           # the step dropped the cached copy
    Q.w.mul_(-1.0)
    eng.invalidate_eval_cache()
    assert check_now() == first
    Q.w[:100].mul_(0.5)                                   # no explicit call: the engine sees torch's version counter move (views share it)
    assert check_now() == first
    assert check_now() == first - 1


def test_fullrank_topk_beyond_32(eng):
    """topk=[10,20,50] with data.split_way=rs: K > 32 is beyond the tensor-core path's candidate lists and runs the exact kernel behind
    the same call (the reference's argsort takes any K)."""
    d = synthetic_data(60, 900, 30, seed=33)
    pu, pi, rp, sc = X.build_history(d.ui_train, d.user_nums)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    rs = np.random.RandomState(5)
    P, Q = (rs.randn(d.user_nums, 32) * 0.1).astype(np.float32), (rs.randn(d.item_nums, 32) * 0.1).astype(np.float32)
    users = np.arange(d.user_nums, dtype=np.int32)
    want, _ = O.fullrank_topk(0, P, Q, users, rp, sc, 50)
    got = eng.score_topk(0, torch.tensor(P).cuda(), torch.tensor(Q).cuda(), users, 50, exact=False)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(150, 3000, 20), (600, 40000, 32), (40, 700, 5)])
def test_fullrank_rescore_fast_path_equals_list_path(eng, kind, shape, monkeypatch):
    """rescore_kernel ranks a user from shared memory / registers (rescore_user_fast: bisection on a SUBSET of the list entries when the
    lists are long, survivors compacted) or, beyond 256 survivors, list by list in global memory (rescore_user_lists, the only path until
    round 2): same ids and score bits from both and from the exact CUDA-core kernel, at split counts from a few to ~50 (two lists per split)."""
    n_users, n_items, K = shape
    d = synthetic_data(n_users, n_items, 30, seed=kind + 10)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    rs = np.random.RandomState(kind + 3)
    dim = 64
    P = torch.tensor((rs.randn(d.user_nums, dim) * 0.1).astype(np.float32)).cuda()
    Q = torch.tensor((rs.randn(d.item_nums, dim) * 0.1).astype(np.float32)).cuda()
    hvec = torch.tensor((rs.randn(dim if kind == 1 else d.item_nums) * 0.1).astype(np.float32)).cuda() if kind in (1, 3) else None
    users = np.arange(d.user_nums, dtype=np.int32)
    monkeypatch.delenv("CRB_RESCORE_LISTS", raising=False)
    fast_i, fast_s = eng.score_topk(kind, P, Q, users, K, hvec=hvec, return_scores=True)
    st = eng.score_topk_stats()
    assert st["certified"] + st["exact_rerun"] == d.user_nums and st["certified"] >= 0.9 * d.user_nums
    monkeypatch.setenv("CRB_RESCORE_LISTS", "1")
    list_i, list_s = eng.score_topk(kind, P, Q, users, K, hvec=hvec, return_scores=True)
    monkeypatch.delenv("CRB_RESCORE_LISTS", raising=False)
    exact_i, exact_s = eng.score_topk(kind, P, Q, users, K, hvec=hvec, exact=True, return_scores=True)
    assert np.array_equal(fast_i, list_i) and np.array_equal(fast_s.view(np.uint32), list_s.view(np.uint32))
    assert np.array_equal(fast_i, exact_i) and np.array_equal(fast_s.view(np.uint32), exact_s.view(np.uint32))


def test_fullrank_rescore_many_survivors_take_the_list_path(eng):
    """Every item row identical: every candidate of a user scores the same, nothing can be filtered, a user has thousands of survivors
    (> 256 = what rescore_user_fast keeps in registers) and is ranked by rescore_user_lists or the exact re-run: the top K are the K
    smallest unseen ids (tie rule: ascending index), as from the exact kernel."""
    d = synthetic_data(64, 5000, 20, seed=4)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    rs = np.random.RandomState(0)
    P = torch.tensor((rs.randn(d.user_nums, 32) * 0.1).astype(np.float32)).cuda()
    Q = torch.tensor(np.tile((rs.randn(1, 32) * 0.1).astype(np.float32), (d.item_nums, 1))).cuda()
    users = np.arange(d.user_nums, dtype=np.int32)
    got = eng.score_topk(0, P, Q, users, 20)
    want = eng.score_topk(0, P, Q, users, 20, exact=True)
    assert np.array_equal(got, want)
    pu, pi, rp, sc = X.build_history(d.ui_train, d.user_nums)
    for u in (0, 5, 63):
        seen = set(sc[rp[u]:rp[u + 1]].tolist())
        assert got[u].tolist() == [i for i in range(d.item_nums) if i not in seen][:20]


def test_fullrank_popularity_sorted_catalogue_stays_exact(eng):
    """Item ids sorted by 'popularity' (the first 1500 rows have 4x the norm, so every user's best items sit in the first item split):
    the per-list compaction margin assumes an even spread, so certificates may fail -- the results must still be the exact kernel's (failed
    users are re-run exactly), and a second call on the same table, which runs with the two-list margin if the first one re-ran more than
    1/64 of its users, returns the same."""
    d = synthetic_data(600, 40000, 30, seed=21)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    rs = np.random.RandomState(21)
    P = torch.tensor((rs.randn(d.user_nums, 64) * 0.1).astype(np.float32)).cuda()
    q = (rs.randn(d.item_nums, 64) * 0.1).astype(np.float32)
    q[:1500] *= 4.0
    Q = torch.tensor(q).cuda()
    users = np.arange(d.user_nums, dtype=np.int32)
    want_i, want_s = eng.score_topk(0, P, Q, users, 20, exact=True, return_scores=True)
    assert int(want_i.max()) < 1500                                   # the skew is what the test says it is
    for call in range(2):
        got_i, got_s = eng.score_topk(0, P, Q, users, 20, return_scores=True)
        st = eng.score_topk_stats()
        assert st["certified"] + st["exact_rerun"] == d.user_nums
        assert np.array_equal(got_i, want_i) and np.array_equal(got_s.view(np.uint32), want_s.view(np.uint32)), call
