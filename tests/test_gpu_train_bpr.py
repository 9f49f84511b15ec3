"""-m gpu: the fused BPR step (csrc/train.cu) against the torch restatement of model/ranking/BPR.py:31-44 with
TF-1 optimizer semantics.  Tolerance: 2e-5 relative (north_star allows 1e-4) on every table entry, 1e-5 on the loss.  (The device
result is bit-identical from run to run; the torch-CPU restatement's own fp32 round-off moves with the host's vector width and
thread count -- scripts/smoke_stress.py measured the worst entry at 0.77 of a 1e-5 tolerance on one box.)"""
import numpy as np
import pytest
import torch

from conftest import synthetic_data
from oracle import philox as X
from oracle import tf1_restatement as T

pytestmark = pytest.mark.gpu
RTOL, ATOL = 2e-5, 4e-7


@pytest.fixture(scope="module")
def eng():
    from cleverrec_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _make(eng, U, I, d, kind, mode, seed=0, scale=0.1):
    from cleverrec_b200.engine import Optimizer, Table
    g = torch.Generator().manual_seed(seed)
    P0, Q0 = torch.randn(U, d, generator=g) * scale, torch.randn(I, d, generator=g) * scale
    opt = Optimizer(kind, 0.05 if kind != "Adam" else 0.01, adam_mode=mode)
    P, Q = Table(P0.cuda(), kind, mode), Table(Q0.cuda(), kind, mode)
    ref = {"P": P0.clone(), "Q": Q0.clone()}
    ropt = T.TF1Optimizer(kind, opt.lr, adam_mode=mode)
    return P, Q, opt, ref, ropt


def _compare(eng, P, Q, opt, ref, ref64=None):
    eng.adam_flush(P, opt)
    eng.adam_flush(Q, opt)
    torch.cuda.synchronize()
    for T_, name in ((P, "P"), (Q, "Q")):
        got, want = T_.w.cpu().numpy(), ref[name].numpy()
        if opt.kind != "Adam":
            np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
            continue
        # Adam's update lr*m/(sqrt(v)+eps) is sign-like where |g| ~ eps: an element whose summed gradient nearly cancels moves
        # by a different fraction of one step depending on the fp32 summation order, in ANY fp32 implementation.  So: (i) all
        # but a vanishing fraction of elements agree to 1e-4 relative; (ii) measured against the fp64 run of the same graph,
        # the CUDA path is as accurate as the fp32 restatement itself.
        bad = ~np.isclose(got, want, rtol=1e-4, atol=1e-5)
        assert bad.mean() <= 1e-3, (name, bad.sum())
        assert np.abs(got - want).max() <= 0.02 * opt.lr  # never more than 2% of one step
        if ref64 is not None:
            truth = ref64[name].numpy()
            err_cuda, err_ref = np.abs(got - truth).max(), np.abs(want.astype(np.float64) - truth).max()
            assert err_cuda <= max(4 * err_ref, 1e-6), (name, err_cuda, err_ref)


OPTS = [("SGD", "tf1"), ("Adagrad", "tf1"), ("Adam", "tf1"), ("Adam", "lazy")]


@pytest.mark.parametrize("kind,mode", OPTS)
@pytest.mark.parametrize("d", [8, 32, 64, 100, 128, 256, 512])
def test_steps_match_restatement(eng, kind, mode, d):
    U, I = 40, 60
    P, Q, opt, ref, ropt = _make(eng, U, I, d, kind, mode, seed=d)
    ref64, ropt64 = {k: v.double() for k, v in ref.items()}, T.TF1Optimizer(kind, opt.lr, adam_mode=mode)
    rs = np.random.RandomState(d)
    for B in (64, 1, 257, 64):  # duplicates guaranteed (B > rows), B=1, odd size
        u, i, j = rs.randint(0, U, B), rs.randint(0, I, B), rs.randint(0, I, B)
        loss = eng.train_step_bpr(P, Q, opt, u, i, j, reg=0.01)  # host feed, host loss (the e2e path)
        b = {"u": torch.tensor(u), "i": torch.tensor(i), "j": torch.tensor(j)}
        rloss = T.train_step(T.bpr_loss, ref, b, {"reg": 0.01}, ropt, sparse_index={"P": ["u"], "Q": ["i", "j"]})
        T.train_step(T.bpr_loss, ref64, b, {"reg": 0.01}, ropt64, sparse_index={"P": ["u"], "Q": ["i", "j"]})
        assert abs(loss - rloss) <= RTOL * abs(rloss)
    _compare(eng, P, Q, opt, ref, ref64)


@pytest.mark.parametrize("kind,mode", OPTS)
def test_hub_rows_multi_chunk_reduction(eng, kind, mode):
    # every triplet hits the same user and the same positive: 700 occurrences -> 3 chunks of the duplicate reduction
    U, I, d, B = 5, 900, 64, 700
    P, Q, opt, ref, ropt = _make(eng, U, I, d, kind, mode, seed=1)
    u, i, j = np.full(B, 3), np.full(B, 7), np.arange(100, 100 + B)
    for _ in range(2):
        loss = eng.train_step_bpr(P, Q, opt, u, i, j, reg=0.001)
        b = {"u": torch.tensor(u), "i": torch.tensor(i), "j": torch.tensor(j)}
        rloss = T.train_step(T.bpr_loss, ref, b, {"reg": 0.001}, ropt, sparse_index={"P": ["u"], "Q": ["i", "j"]})
        assert abs(loss - rloss) <= 1e-4 * abs(rloss)
    eng.adam_flush(P, opt); eng.adam_flush(Q, opt)
    np.testing.assert_allclose(P.w.cpu().numpy(), ref["P"].numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(Q.w.cpu().numpy(), ref["Q"].numpy(), rtol=1e-4, atol=1e-6)


def test_adam_tf1_untouched_rows_keep_moving(eng):
    # tf.train.AdamOptimizer's sparse apply moves every row every step; rows touched only at step 1 must
    # match the dense restatement after 12 further steps that never touch them (replay at flush).
    U, I, d = 30, 30, 32
    P, Q, opt, ref, ropt = _make(eng, U, I, d, "Adam", "tf1", seed=4)
    rs = np.random.RandomState(0)
    steps = [(np.arange(30), np.arange(30), (np.arange(30) + 1) % 30)] + [(rs.randint(0, 5, 16), rs.randint(0, 5, 16), rs.randint(5, 10, 16)) for _ in range(12)]
    for u, i, j in steps:
        eng.train_step_bpr(P, Q, opt, u, i, j, reg=0.01)
        b = {"u": torch.tensor(u), "i": torch.tensor(i), "j": torch.tensor(j)}
        T.train_step(T.bpr_loss, ref, b, {"reg": 0.01}, ropt, sparse_index={"P": ["u"], "Q": ["i", "j"]})
    _compare(eng, P, Q, opt, ref)
    assert not np.allclose(ref["P"][20:].numpy(), T.to_torch({"x": np.zeros(1)})["x"].numpy())


def test_device_feed_equals_host_feed_and_is_deterministic(eng):
    U, I, d = 2000, 3000, 128  # every row occurs <= 32 times: slot sums are taken in triplet order -> bit-identical runs
    rs = np.random.RandomState(3)
    u, i, j = rs.randint(0, U, 4096), rs.randint(0, I, 4096), rs.randint(0, I, 4096)
    outs = []
    for feed in ("host", "device", "device"):
        P, Q, opt, _, _ = _make(eng, U, I, d, "Adagrad", "tf1", seed=8)
        if feed == "host":
            loss = eng.train_step_bpr(P, Q, opt, u, i, j, reg=0.01)
        else:
            lo = torch.zeros(1, dtype=torch.float64, device="cuda")
            eng.train_step_bpr(P, Q, opt, torch.tensor(u, dtype=torch.int32).cuda(), torch.tensor(i, dtype=torch.int32).cuda(),
                               torch.tensor(j, dtype=torch.int32).cuda(), reg=0.01, loss_out=lo)
            loss = float(lo.item())
        outs.append((loss, P.w.cpu().numpy().copy(), Q.w.cpu().numpy().copy()))
    for k in (1, 2):  # bit-identical: slot sums of rows with <= 32 occurrences are taken in triplet order
        assert outs[k][0] == outs[0][0]
        assert np.array_equal(outs[k][1], outs[0][1]) and np.array_equal(outs[k][2], outs[0][2])


@pytest.mark.parametrize("kind,mode", [("Adam", "tf1"), ("SGD", "tf1")])
def test_fused_epoch_equals_step_by_step_on_twin_triplets(eng, kind, mode):
    """crb_train_epoch_bpr (sampler fused on the device) == feeding the CPU twin's triplets step by step."""
    d = synthetic_data(120, 400, 25, seed=6)
    pu, pi, rp, sc = X.build_history(d.ui_train, d.user_nums)
    eng.set_history(d.ui_train, d.user_nums, d.item_nums)
    R, B, dim = 4, 1000, 64
    n = pu.shape[0] * R
    n_steps = (n + B - 1) // B
    P, Q, opt, ref, ropt = _make(eng, d.user_nums, d.item_nums, dim, kind, mode, seed=2)
    losses = torch.zeros(n_steps, dtype=torch.float64, device="cuda")
    eng.train_epoch_bpr(P, Q, opt, 42, 0, 0, B, n_steps, R, 0.01, losses)
    losses = losses.cpu().numpy()
    u, i, j, _ = X.sample_pairwise(42, 0, 0, n, R, d.item_nums, pu, pi, rp, sc)
    for k in range(n_steps):
        sl = slice(k * B, min((k + 1) * B, n))
        b = {"u": torch.tensor(u[sl].astype(np.int64)), "i": torch.tensor(i[sl].astype(np.int64)), "j": torch.tensor(j[sl].astype(np.int64))}
        rloss = T.train_step(T.bpr_loss, ref, b, {"reg": 0.01}, ropt, sparse_index={"P": ["u"], "Q": ["i", "j"]})
        assert abs(losses[k] - rloss) <= RTOL * abs(rloss), k
    _compare(eng, P, Q, opt, ref)


def test_argument_errors(eng):
    from cleverrec_b200._lib import CrbError
    from cleverrec_b200.engine import Optimizer, Table
    P = Table(torch.zeros(4, 6).cuda(), "SGD")  # dim % 4 != 0
    Q = Table(torch.zeros(4, 6).cuda(), "SGD")
    with pytest.raises(CrbError) as e:
        eng.train_step_bpr(P, Q, Optimizer("SGD", 0.1), [0], [1], [2], reg=0.0)
    assert e.value.code == -1
    with pytest.raises(ValueError):
        Optimizer("RMSProp", 0.1)


@pytest.mark.parametrize("kind", ["SGD", "Adam"])
def test_epoch_over_host_feeds_equals_step_by_step(kind):
    """crb_train_epoch_bpr_feeds (the reference's epoch loop, RankingRecommender.py:39-46, over caller-sampled host arrays with the
    feeds staged one step ahead) == one blocking crb_train_step_bpr per slice: same losses, same tables, bit for bit; ragged tail;
    pinned / pageable / device feeds and loss buffers."""
    import torch
    from cleverrec_b200.engine import Engine, Optimizer, Table
    eng = Engine(0)
    U, I, d, B = 300, 500, 64, 1000
    g = torch.Generator().manual_seed(2)
    P0, Q0 = torch.randn(U, d, generator=g) * 0.1, torch.randn(I, d, generator=g) * 0.1
    rs = np.random.RandomState(4)
    n = 4 * B + 137
    u, i, j = rs.randint(0, U, n).astype(np.int32), rs.randint(0, I, n).astype(np.int32), rs.randint(0, I, n).astype(np.int32)
    Pa, Qa, oa = Table(P0.clone().cuda(), kind), Table(Q0.clone().cuda(), kind), Optimizer(kind, 0.01)
    want = [eng.train_step_bpr(Pa, Qa, oa, u[k:k + B], i[k:k + B], j[k:k + B], 0.01) for k in range(0, n, B)]
    eng.adam_flush(Pa, oa); eng.adam_flush(Qa, oa)
    for mode in ("pageable", "pinned", "device"):
        Pb, Qb, ob = Table(P0.clone().cuda(), kind), Table(Q0.clone().cuda(), kind), Optimizer(kind, 0.01)
        if mode == "pageable":
            feeds, losses = (u, i, j), np.zeros(5)
        elif mode == "pinned":
            feeds, losses = tuple(torch.from_numpy(x).pin_memory() for x in (u, i, j)), torch.zeros(5, dtype=torch.float64).pin_memory()
        else:
            feeds, losses = tuple(torch.from_numpy(x).cuda() for x in (u, i, j)), torch.zeros(5, dtype=torch.float64, device="cuda")
        eng.train_epoch_bpr_feeds(Pb, Qb, ob, feeds[0], feeds[1], feeds[2], B, 0.01, losses)
        got = losses.tolist() if mode == "pageable" else losses.cpu().tolist()
        assert got == want, mode
        assert ob.t == oa.t == 5
        eng.adam_flush(Pb, ob); eng.adam_flush(Qb, ob)
        assert torch.equal(Pb.w, Pa.w) and torch.equal(Qb.w, Qa.w), mode
    eng.close()


@pytest.mark.parametrize("kind,mode", OPTS)
@pytest.mark.parametrize("d", [32, 64, 100, 128, 256, 512])
def test_ring_kernel_is_bit_identical_to_register_kernel(eng, kind, mode, d, monkeypatch):
    """bpr_ring_kernel (row gathers as cp.async.bulk copies into a shared-memory ring, mbarrier hand-off, CRB_BPR_RING=1) runs the
    same device functions on the same values as bpr_step_kernel: tables bit-identical after several device-sampled steps with
    duplicated rows, hub rows, ragged last batch and (Adam tf1) rows that skip steps and are replayed."""
    from cleverrec_b200.engine import Optimizer, Table
    U, I, B, R = 700, 1500, 1000, 3     # row multiplicities stay <= 32: summation order (and so every bit) is defined
    data = synthetic_data(U, I, 14, seed=d)
    eng.set_history(data.ui_train, U, I)
    n_rows = eng.epoch_rows(R)
    n_steps = min(5, -(-n_rows // B))
    g = torch.Generator().manual_seed(d)
    P0, Q0 = torch.randn(U, d, generator=g) * 0.1, torch.randn(I, d, generator=g) * 0.1
    out = []
    for ring in ("0", "1"):
        monkeypatch.setenv("CRB_BPR_RING", ring)
        P, Q = Table(P0.clone().cuda(), kind, mode), Table(Q0.clone().cuda(), kind, mode)
        opt = Optimizer(kind, 0.05 if kind != "Adam" else 0.01, adam_mode=mode)
        losses = torch.zeros(n_steps, dtype=torch.float64, device="cuda")
        eng.train_epoch_bpr(P, Q, opt, 3, 0, n_rows - n_steps * B if n_rows % B == 0 else n_rows - (n_steps - 1) * B - n_rows % B, B, n_steps, R, 0.01, losses)
        # a host-fed step with a 30-fold row and a 7-row step
        rs = np.random.RandomState(1)
        u, i, j = rs.randint(0, U, 777), rs.randint(0, I, 777), rs.randint(0, I, 777)
        i[:750:25] = 5
        assert np.bincount(np.concatenate([i, j])).max() <= 32 and np.bincount(u).max() <= 32
        l1 = eng.train_step_bpr(P, Q, opt, u, i, j, 0.01)
        l2 = eng.train_step_bpr(P, Q, opt, u[:7], i[:7], j[:7], 0.01)
        eng.adam_flush(P, opt); eng.adam_flush(Q, opt)
        torch.cuda.synchronize()
        out.append((losses.cpu().numpy(), l1, l2, P.w.clone(), Q.w.clone(), None if P.s1 is None else P.s1.clone(), None if Q.s1 is None else Q.s1.clone()))
    a, b = out
    np.testing.assert_allclose(a[0], b[0], rtol=1e-12)
    assert abs(a[1] - b[1]) <= 1e-12 * abs(a[1]) and abs(a[2] - b[2]) <= 1e-12 * abs(a[2])
    for k in (3, 4, 5, 6):
        assert (a[k] is None and b[k] is None) or torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("kind,mode", OPTS)
def test_epoch_graph_is_bit_identical(eng, kind, mode, monkeypatch):
    """CRB_EPOCH_GRAPH=1 replays a captured whole-epoch CUDA graph (one launch per epoch; sampler keys / epoch word / optimizer step
    read from device memory).  Three epochs through the graph == the same epochs step by step (the default): losses and
    tables bit for bit, including the epoch-dependent sampling and Adam's step-dependent lr_t and replay."""
    from cleverrec_b200.engine import Optimizer, Table
    U, I, d, B, R = 400, 900, 64, 512, 2
    data = synthetic_data(U, I, 12, seed=7)
    eng.set_history(data.ui_train, U, I)
    n_rows = eng.epoch_rows(R)
    n_steps = -(-n_rows // B)
    assert n_steps >= 8
    g = torch.Generator().manual_seed(2)
    P0, Q0 = torch.randn(U, d, generator=g) * 0.1, torch.randn(I, d, generator=g) * 0.1
    out = []
    for no_graph in (False, True):
        if no_graph:
            monkeypatch.delenv("CRB_EPOCH_GRAPH", raising=False)
        else:
            monkeypatch.setenv("CRB_EPOCH_GRAPH", "1")
        P, Q = Table(P0.clone().cuda(), kind, mode), Table(Q0.clone().cuda(), kind, mode)
        opt = Optimizer(kind, 0.05 if kind != "Adam" else 0.01, adam_mode=mode)
        all_losses = []
        l0 = eng.launches
        for epoch in range(3):
            losses = torch.zeros(n_steps, dtype=torch.float64, device="cuda")
            eng.train_epoch_bpr(P, Q, opt, 5, epoch, 0, B, n_steps, R, 0.01, losses)
            all_losses.append(losses.cpu().numpy())
        assert eng.launches - l0 == 3 * n_steps * 4   # the same four kernels per step either way (sample, assign, step, dup_tail)
        eng.adam_flush(P, opt); eng.adam_flush(Q, opt)
        torch.cuda.synchronize()
        out.append((np.concatenate(all_losses), P.w.clone(), Q.w.clone()))
    assert np.array_equal(out[0][0], out[1][0])
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    assert out[0][0][-1] < out[0][0][0]


def test_bad_feed_id_is_an_error(eng):
    """A host feed with an id outside the table raises (TF's gather would) instead of reading out of bounds."""
    from cleverrec_b200._lib import CrbError
    P, Q, opt, ref, ropt = _make(eng, 50, 60, 32, "SGD", "tf1")
    u, i, j = np.array([1, 2, 3]), np.array([4, 60, 6]), np.array([7, 8, 9])
    with pytest.raises(CrbError, match="i_idx"):
        eng.train_step_bpr(P, Q, opt, u, i, j, 0.01)
    with pytest.raises(CrbError, match="u_idx"):
        eng.train_step_bpr(P, Q, opt, np.array([1, -1, 3]), np.array([4, 5, 6]), j, 0.01)


@pytest.mark.parametrize("kind,mode", [("SGD", "tf1"), ("Adam", "tf1")])
def test_determinism_across_runs_and_hub_rows(eng, kind, mode):
    """Two runs of the same steps: every row with <= 32 occurrences per step is BIT-identical (its duplicate gradients are summed in
    triplet order); a hub row with hundreds of occurrences is summed in slot order, which follows the arrival order of the counting
    atomics -- its two results agree to fp32 round-off (documented deviation: DESIGN 4), never more."""
    from cleverrec_b200.engine import Optimizer, Table
    U, I, d, B = 3000, 2000, 64, 4096
    rs = np.random.RandomState(4)
    feeds = []
    for step in range(3):
        u, i, j = rs.randint(0, U, B), rs.randint(0, I, B), rs.randint(0, I, B)
        i[::7] = 11          # hub item: ~585 occurrences per step (3 chunks of 256 slots)
        u[::16] = 5          # hub user: 256 occurrences
        feeds.append((u, i, j))
    g = torch.Generator().manual_seed(0)
    P0, Q0 = torch.randn(U, d, generator=g) * 0.1, torch.randn(I, d, generator=g) * 0.1
    runs = []
    for rep in range(2):
        P, Q = Table(P0.clone().cuda(), kind, mode), Table(Q0.clone().cuda(), kind, mode)
        opt = Optimizer(kind, 0.05 if kind != "Adam" else 0.01, adam_mode=mode)
        snaps = []
        for f in feeds:
            eng.train_step_bpr(P, Q, opt, f[0], f[1], f[2], 0.01)
            torch.cuda.synchronize()
            snaps.append((P.w.cpu().numpy().copy(), Q.w.cpu().numpy().copy()))
        runs.append(snaps)
    hub_i = np.bincount(np.concatenate([feeds[0][1], feeds[0][2]]), minlength=I) > 32
    hub_u = np.bincount(feeds[0][0], minlength=U) > 32
    assert hub_i.sum() == 1 and hub_u.sum() == 1
    # after ONE step (later steps read the hub rows, so their round-off spreads): everything but the two hub rows is bit-identical
    assert np.array_equal(runs[0][0][0][~hub_u], runs[1][0][0][~hub_u])
    assert np.array_equal(runs[0][0][1][~hub_i], runs[1][0][1][~hub_i])
    tol = 1e-6 if kind == "SGD" else 2e-4   # Adam: one sign-like step of lr where a summed gradient nearly cancels
    for step in range(3):
        assert np.abs(runs[0][step][0] - runs[1][step][0]).max() <= tol * (step + 1)
        assert np.abs(runs[0][step][1] - runs[1][step][1]).max() <= tol * (step + 1)


@pytest.mark.parametrize("kind,mode", OPTS)
def test_fused_tail_is_bit_identical_to_separate_launches(eng, kind, mode, monkeypatch):
    """Batches up to 2^16 rows end a step with ONE launch (dup_tail_kernel: duplicate reduction, the multi-chunk rows and the loss sum by
    the block that finishes last, counters reset for the next step) instead of dup_reduce + dup_final + loss_final + a memset.
    CRB_DUP_TAIL=0 keeps the separate launches: losses and tables bit for bit over three epochs; then two steps on a hub row with 700
    occurrences (three chunks -> the last block's dup_final share; > 32 occurrences are summed in slot order, so 1e-6, not bits)."""
    from cleverrec_b200.engine import Optimizer, Table
    U, I, d, B, R = 400, 900, 64, 512, 2
    data = synthetic_data(U, I, 12, seed=9)
    eng.set_history(data.ui_train, U, I)
    n_steps = -(-eng.epoch_rows(R) // B)
    g = torch.Generator().manual_seed(4)
    P0, Q0 = torch.randn(U, d, generator=g) * 0.1, torch.randn(I, d, generator=g) * 0.1
    hub = (np.full(700, 3), np.full(700, 7), np.arange(100, 800))
    out = []
    for tail in ("0", "1"):
        monkeypatch.setenv("CRB_DUP_TAIL", tail)
        P, Q = Table(P0.clone().cuda(), kind, mode), Table(Q0.clone().cuda(), kind, mode)
        opt = Optimizer(kind, 0.05 if kind != "Adam" else 0.01, adam_mode=mode)
        got = []
        l0 = eng.launches
        for epoch in range(3):
            losses = torch.zeros(n_steps, dtype=torch.float64, device="cuda")
            eng.train_epoch_bpr(P, Q, opt, 5, epoch, 0, B, n_steps, R, 0.01, losses)
            got.append(losses.cpu().numpy())
        assert eng.launches - l0 == 3 * n_steps * (4 if tail == "1" else 6)
        torch.cuda.synchronize()
        snap = (np.concatenate(got), P.w.clone(), Q.w.clone())
        hub_losses = [eng.train_step_bpr(P, Q, opt, *hub, reg=0.001) for _ in range(2)]
        eng.adam_flush(P, opt); eng.adam_flush(Q, opt)
        torch.cuda.synchronize()
        out.append(snap + (np.array(hub_losses), P.w.clone(), Q.w.clone()))
    monkeypatch.delenv("CRB_DUP_TAIL", raising=False)
    assert np.array_equal(out[0][0], out[1][0])
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    np.testing.assert_allclose(out[0][3], out[1][3], rtol=1e-5)
    assert torch.allclose(out[0][4], out[1][4], rtol=1e-5, atol=1e-6) and torch.allclose(out[0][5], out[1][5], rtol=1e-5, atol=1e-6)
