"""-m gpu: the hot path at BASELINE.json's FULL size (configs[4]: 10M users x 2M items, d = 128, ~1e9 interactions, 2^20 triplets per
step) checked through properties that do not need the CPU oracle to finish: every sampled triplet is admissible against the history,
windows of the epoch agree, one SGD step equals an independent plain-torch fp32 formulation of BPR.py:31-44 on the same triplets
(loss, touched rows, untouched rows bit-unchanged), two Adam steps equal TF-1's dense-moment Adam written with dense torch ops, and the
tensor-core full-rank top-20 equals the exact path and a torch matmul.  The torch formulations are the 'plain PyTorch fp32 reference'
of the same op; the small-size bit-exact parity against the oracle lives in the other test_gpu_* files."""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

USERS, ITEMS, DIM, MEAN_HIST, B, R = 10_000_000, 2_000_000, 128, 100, 1 << 20, 4


@pytest.fixture(scope="module")
def world():
    free, _ = torch.cuda.mem_get_info(0)
    if free < 100 * (1 << 30):
        pytest.skip("needs ~100 GB of free HBM (full-size tables, history and the torch reference copies)")
    from bench import build_history_device
    from cleverrec_b200.engine import Engine
    dev = torch.device("cuda", 0)
    eng = Engine(0)
    pu, pi, rowptr = build_history_device(torch, dev, USERS, ITEMS, MEAN_HIST, seed=1234)
    eng.set_history_arrays(USERS, ITEMS, pu, pi, rowptr, pi)
    keys = pu.to(torch.int64) * ITEMS + pi.to(torch.int64)          # ascending: users ascending, items sorted inside a user
    yield eng, dev, keys, pu, pi, rowptr
    eng.close()


def _member(keys, u, i):
    q = u.to(torch.int64) * ITEMS + i.to(torch.int64)
    pos = torch.searchsorted(keys, q).clamp_(max=keys.numel() - 1)
    return keys[pos] == q


def test_sampler_is_admissible_and_windows_agree(world):
    eng, dev, keys, pu, pi, rowptr = world
    assert keys.numel() > 9e8 and bool((keys[1:] > keys[:-1]).all())
    rows = eng.epoch_rows(R, "pairwise")
    assert rows == R * keys.numel()
    u, i, j = eng.sample_pairwise(7, 0, 0, B, R)
    assert bool(_member(keys, u, i).all())                            # (u, i) is a training positive   (utils/sampler.py:50-52)
    assert not bool(_member(keys, u, j).any())                        # j is not in u's history          (:58-59)
    assert int(j.min()) >= 0 and int(j.max()) < ITEMS
    # negatives are uniform over the (almost whole) catalogue: mean within 6 sigma of (I-1)/2
    assert abs(float(j.double().mean()) - (ITEMS - 1) / 2) < 6 * (ITEMS / math.sqrt(12)) / math.sqrt(B)
    # windows of the same epoch are slices of each other; epochs and seeds differ
    first = rows - 5000                                                # the ragged end of the epoch
    u2, i2, j2 = eng.sample_pairwise(7, 0, first, 5000, R)
    u3, i3, j3 = eng.sample_pairwise(7, 0, first + 1000, 3000, R)
    assert torch.equal(u2[1000:4000], u3) and torch.equal(i2[1000:4000], i3) and torch.equal(j2[1000:4000], j3)
    ua, ia, ja = eng.sample_pairwise(7, 0, 0, B, R)
    assert torch.equal(ua, u) and torch.equal(ja, j)
    ub, _, jb = eng.sample_pairwise(7, 1, 0, B, R)
    assert not torch.equal(ub, u) and not torch.equal(jb, j)
    # the shuffle spreads a batch over the user range (global permutation, utils/sampler.py:68)
    assert int(u.min()) < USERS // 50 and int(u.max()) > USERS - USERS // 50


def _bpr_torch(P0, Q0, u, i, j, reg):
    """BPR.py:31-44 in plain torch fp32 on the device: -> (loss fp64, dense grad P, dense grad Q)."""
    ul, il, jl = u.long(), i.long(), j.long()
    p, qi, qj = P0[ul], Q0[il], Q0[jl]
    x = (p * qi).sum(1) - (p * qj).sum(1)
    loss = torch.nn.functional.softplus(-x).double().sum() + reg * 0.5 * ((p * p).double().sum() + (qi * qi).double().sum() + (qj * qj).double().sum())
    g = -torch.sigmoid(-x)[:, None]
    GP = torch.zeros_like(P0).index_add_(0, ul, g * (qi - qj) + reg * p)
    GQ = torch.zeros_like(Q0).index_add_(0, il, g * p + reg * qi).index_add_(0, jl, -g * p + reg * qj)
    return float(loss), GP, GQ


def test_sgd_step_equals_plain_torch_and_leaves_other_rows_alone(world):
    from cleverrec_b200.engine import Optimizer, Table
    eng, dev, keys, pu, pi, rowptr = world
    g = torch.Generator(device=dev).manual_seed(5)
    P0 = torch.randn(USERS, DIM, device=dev, generator=g) * 0.1
    Q0 = torch.randn(ITEMS, DIM, device=dev, generator=g) * 0.1
    P, Q = Table(P0.clone(), "SGD"), Table(Q0.clone(), "SGD")
    u, i, j = eng.sample_pairwise(11, 0, 12345, B, R)
    lr, reg = 0.05, 0.01
    loss = eng.train_step_bpr(P, Q, Optimizer("SGD", lr), u, i, j, reg)
    want, GP, GQ = _bpr_torch(P0, Q0, u, i, j, reg)
    assert abs(loss - want) <= 2e-6 * abs(want), (loss, want)
    for got, w0, G, idx, name in ((P.w, P0, GP, u, "P"), (Q.w, Q0, GQ, torch.cat([i, j]), "Q")):
        ref = w0 - lr * G
        err = (got - ref).abs().max().item()
        assert err <= 2e-7, (name, err)                                 # |w| ~ 0.1..0.5: a few fp32 ulps (summation order of duplicates)
        touched = torch.zeros(w0.shape[0], dtype=torch.bool, device=dev)
        touched[idx.long()] = True
        changed = (got != w0).any(1)
        assert not bool((changed & ~touched).any())                     # rows outside the batch are bit-unchanged
        assert int(changed.sum()) >= 0.999 * int(touched.sum())
    # lr = 0: a step changes nothing (idempotence of the gather / scatter plumbing)
    P2, Q2 = Table(P0.clone(), "SGD"), Table(Q0.clone(), "SGD")
    eng.train_step_bpr(P2, Q2, Optimizer("SGD", 0.0), u, i, j, reg)
    assert torch.equal(P2.w, P0) and torch.equal(Q2.w, Q0)


def test_two_adam_steps_equal_dense_tf1_adam(world):
    """tf.train.AdamOptimizer's sparse apply decays the moments of EVERY row and moves every row; the product replays
    missed decay steps lazily and flushes before a read.  Reference: the dense recurrence with dense torch ops on the full tables."""
    from cleverrec_b200.engine import Optimizer, Table
    eng, dev, keys, pu, pi, rowptr = world
    g = torch.Generator(device=dev).manual_seed(6)
    P0 = torch.randn(USERS, DIM, device=dev, generator=g) * 0.1
    Q0 = torch.randn(ITEMS, DIM, device=dev, generator=g) * 0.1
    P, Q = Table(P0.clone(), "Adam", "tf1"), Table(Q0.clone(), "Adam", "tf1")
    opt = Optimizer("Adam", 1e-3, adam_mode="tf1")
    lr, b1, b2, eps, reg = 1e-3, 0.9, 0.999, 1e-8, 0.01
    Pr, Qr = P0.clone(), Q0.clone()
    mP, vP, mQ, vQ = torch.zeros_like(P0), torch.zeros_like(P0), torch.zeros_like(Q0), torch.zeros_like(Q0)
    for step in (1, 2, 3):
        u, i, j = eng.sample_pairwise(3, 0, (step - 1) * B, B, R)
        loss = eng.train_step_bpr(P, Q, opt, u, i, j, reg)
        want, GP, GQ = _bpr_torch(Pr, Qr, u, i, j, reg)
        assert abs(loss - want) <= 5e-6 * abs(want), (step, loss, want)
        lr_t = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        for w, m, v, G in ((Pr, mP, vP, GP), (Qr, mQ, vQ, GQ)):
            m.mul_(b1).add_(G, alpha=1 - b1)
            v.mul_(b2).addcmul_(G, G, value=1 - b2)
            w.sub_(lr_t * m / (v.sqrt() + eps))
        del GP, GQ
    eng.adam_flush(P, opt)
    eng.adam_flush(Q, opt)
    for got, ref, w0, name in ((P.w, Pr, P0, "P"), (Q.w, Qr, Q0, "Q")):
        moved = (ref - w0).abs().max().item()
        err = (got - ref).abs()
        assert moved > 1e-3                                             # Adam's first steps move touched entries by ~lr each
        # Adam divides by sqrt(v) + eps: where a gradient entry cancels down to ~eps = 1e-8 the update lr * g / (|g| + eps) is
        # ill-conditioned (d/dg = lr / (4 eps)), so fp32 summation-order differences of 1e-9 in g move w by up to ~1e-4 there.  The bar
        # is therefore the one of DESIGN.md section 4: >= 99.9 % of the entries within 2e-6 (1e-4 relative at |w| ~ 0.02) and no entry
        # off by more than a tenth of what three steps can move it.
        assert (err > 2e-6).float().mean().item() < 1e-3, (name, (err > 2e-6).float().mean().item())
        assert err.max().item() <= 3e-4, (name, err.max().item())
    # the flushed moments are the dense recurrences' (absolute bars: entries that cancel to ~0 have no meaningful relative error)
    # (an entry of w that sits ~1e-4 off after step 1 shifts the later gradients by ~1e-5 and the moments by ~1e-6: same statistical bar)
    em, ev = (P.s1 - mP).abs(), (Q.s2 - vQ).abs()
    assert (em > 1e-7).float().mean().item() < 1e-3 and em.max().item() <= 1e-4 * mP.abs().max().item(), (em.max().item(), mP.abs().max().item())
    assert (ev > 1e-4 * vQ.abs().max().item()).float().mean().item() < 1e-3, (ev.max().item(), vQ.abs().max().item())


def test_fullrank_top20_tensor_core_equals_exact_and_torch(world):
    from cleverrec_b200 import _lib
    eng, dev, keys, pu, pi, rowptr = world
    g = torch.Generator(device=dev).manual_seed(8)
    P = torch.randn(USERS, DIM, device=dev, generator=g) * 0.1
    Q = torch.randn(ITEMS, DIM, device=dev, generator=g) * 0.1
    users = torch.randint(0, USERS, (2048,), device=dev, generator=g, dtype=torch.int32)
    K = 20
    ids, sc = eng.score_topk(_lib.SCORE_DOT, P, Q, users, K, return_scores=True)
    st = eng.score_topk_stats()
    assert st["certified"] + st["exact_rerun"] == users.numel()
    ids_exact = eng.score_topk(_lib.SCORE_DOT, P, Q, users[:96], K, exact=True)
    assert torch.equal(ids[:96], ids_exact)                            # tcgen05 + certificate == fp32 CUDA-core path, id for id
    # returned items are unseen, distinct, in range; scores descend; ties broken by ascending id
    rep = users.repeat_interleave(K)
    assert not bool(_member(keys, rep, ids.reshape(-1)).any())
    assert int(ids.min()) >= 0 and int(ids.max()) < ITEMS
    srt = torch.sort(ids.long(), dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    assert bool(((sc[:, :-1] > sc[:, 1:]) | ((sc[:, :-1] == sc[:, 1:]) & (ids[:, :-1] < ids[:, 1:]))).all())
    # the scores are the fp32 dot products of the returned pairs (fp64 torch on the same rows)
    ref = (P[rep.long()].double() * Q[ids.reshape(-1).long()].double()).sum(1).reshape(-1, K)
    assert (sc.double() - ref).abs().max().item() <= 2e-6
    # nothing unseen beats the K-th returned score: torch matmul over the whole catalogue for 64 users
    sub = users[:64].long()
    full = (P[sub] @ Q.t()).double()
    for k in range(64):
        lo, hi = int(rowptr[sub[k]]), int(rowptr[sub[k] + 1])
        full[k, pi[lo:hi].long()] = -float("inf")
    top = torch.topk(full, K, dim=1)
    assert (top.values[:, -1] - sc[:64, -1].double()).abs().max().item() <= 1e-5
    same = (torch.sort(top.indices, dim=1).values == srt[:64]).float().mean().item()
    assert same >= 0.99                                                 # the tf32/fp32 matmul of torch may swap near-ties at rank 20


def test_sampled_candidate_scores_at_full_size(world, monkeypatch):
    """test_model_loo's scoring at the full table sizes: the tiled pair scorer (rows staged by cp.async) returns the bits of the
    thread-per-pair scorer, both are the fp32 dot products of the pairs (fp64 torch on the same rows), and the per-user top-20 of
    crb_topk_segments is torch.topk's on the same scores."""
    from cleverrec_b200 import _lib
    eng, dev, keys, pu, pi, rowptr = world
    g = torch.Generator(device=dev).manual_seed(9)
    P = torch.randn(USERS, DIM, device=dev, generator=g) * 0.1
    Q = torch.randn(ITEMS, DIM, device=dev, generator=g) * 0.1
    n_users, n_cand = 20000, 101
    u = torch.randint(0, USERS, (n_users,), device=dev, generator=g, dtype=torch.int32).repeat_interleave(n_cand)
    i = torch.randint(0, ITEMS, (n_users * n_cand,), device=dev, generator=g, dtype=torch.int32)
    monkeypatch.delenv("CRB_SCORE_PAIRS_SIMPLE", raising=False)
    tiled = eng.score_pairs(_lib.SCORE_DOT, P, Q, u, i)
    monkeypatch.setenv("CRB_SCORE_PAIRS_SIMPLE", "1")
    simple = eng.score_pairs(_lib.SCORE_DOT, P, Q, u, i)
    monkeypatch.delenv("CRB_SCORE_PAIRS_SIMPLE", raising=False)
    assert torch.equal(tiled.view(torch.int32), simple.view(torch.int32))
    ref = (P[u.long()].double() * Q[i.long()].double()).sum(1)
    assert (tiled.double() - ref).abs().max().item() <= 2e-6
    seg = torch.arange(n_users + 1, device=dev, dtype=torch.int64) * n_cand
    top = eng.topk_segments(tiled, seg, 20).long()
    want = torch.topk(tiled.reshape(n_users, n_cand), 20, dim=1)
    got_scores = torch.gather(tiled.reshape(n_users, n_cand), 1, top)
    assert torch.equal(got_scores, want.values)                          # same score sequence; ids may differ only inside exact ties
    assert bool(((got_scores[:, :-1] > got_scores[:, 1:]) | (top[:, :-1] < top[:, 1:])).all())   # ties: ascending position
