"""Worker for tests/test_gpu2_sharded.py and tests/test_gpu_sharded_one_device.py (launched with torch.distributed.run): ShardedBPR
over peer memory must reproduce the single-GPU step on the union batch.  One process per GPU over NCCL, or -- CRB_SHARED_DEVICE=1 --
every process on device 0 with a gloo control plane: CUDA IPC works between processes on one GPU, so the same kernels
(item_fetch / shard_step / dup_reduce<SHARD> / inbox_apply) run with real cross-process peer pointers on a single-GPU box."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleverrec_b200.dist import ShardedBPR, all_reduce_dev, broadcast_dev, user_range  # noqa: E402
from cleverrec_b200.engine import Engine, Optimizer, Table  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    shared = os.environ.get("CRB_SHARED_DEVICE", "0") == "1"
    if shared:
        local = 0
    torch.cuda.set_device(local)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    U, I, d, B = 64, 101, 64, 256
    g = torch.Generator().manual_seed(0)
    P0, Q0 = torch.randn(U, d, generator=g) * 0.1, torch.randn(I, d, generator=g) * 0.1
    ok = True
    for kind, mode in (("SGD", "tf1"), ("Adagrad", "tf1"), ("Adam", "tf1"), ("Adam", "lazy")):
        lo, hi = user_range(U, rank, world)
        m = ShardedBPR(eng, U, I, d, kind, 0.05 if kind != "Adam" else 0.01, mode, B, init_P=P0[lo:hi], init_Q=Q0)
        ref = None
        if rank == 0:
            ref = (Table(P0.clone().cuda(), kind, mode), Table(Q0.clone().cuda(), kind, mode), Optimizer(kind, m.opt.lr, adam_mode=mode))
        rs = np.random.RandomState(5)
        my_feeds, my_losses = [], []
        for step in range(4):
            feeds = []
            for r in range(world):  # every rank draws all feeds so rank 0 can build the union batch
                l, h_ = user_range(U, r, world)
                n = B if step != 2 else 37
                feeds.append((rs.randint(l, h_, n), rs.randint(0, I, n), rs.randint(0, I, n)))
            u, i, j = feeds[rank]
            loss = m.step(0.01, feed=(u - lo, i, j))
            my_feeds.append((u - lo, i, j)); my_losses.append(loss)
            t = torch.tensor([loss], device="cuda", dtype=torch.float64)
            all_reduce_dev(t)
            if rank == 0:
                uu, ii, jj = (np.concatenate([f[k] for f in feeds]) for k in range(3))
                want = eng.train_step_bpr(ref[0], ref[1], ref[2], uu, ii, jj, 0.01)
                if abs(float(t.item()) - want) > 1e-5 * abs(want):
                    print("LOSS MISMATCH", kind, mode, step, float(t.item()), want)
                    ok = False
        m.check()
        m.flush()
        Qfull = m.gather_Q()
        Pparts = [torch.zeros(user_range(U, r, world)[1] - user_range(U, r, world)[0], d, device="cuda") for r in range(world)]
        for r in range(world):
            if r == rank:
                Pparts[r].copy_(m.P.w)
            broadcast_dev(Pparts[r], r)
        if rank == 0:
            eng.adam_flush(ref[0], ref[2]); eng.adam_flush(ref[1], ref[2])
            rtol, atol = (1e-4, 1e-5) if kind == "Adam" else (1e-5, 2e-7)
            for name, got, want in (("P", torch.cat(Pparts), ref[0].w), ("Q", Qfull, ref[1].w)):
                bad = ~torch.isclose(got, want, rtol=rtol, atol=atol)
                if bad.float().mean() > 1e-3:
                    print("TABLE MISMATCH", kind, mode, name, int(bad.sum()), float((got - want).abs().max()))
                    ok = False
        # the same feeds through run_steps(feeds=...) -- staged one step ahead on the copy stream, losses to pinned host memory --
        # must give this rank's step-by-step losses and tables bit for bit
        m2 = ShardedBPR(eng, U, I, d, kind, 0.05 if kind != "Adam" else 0.01, mode, B, init_P=P0[lo:hi], init_Q=Q0)
        hl = torch.zeros(4, dtype=torch.float64).pin_memory()
        pinned = [tuple(torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).pin_memory() for x in f) for f in my_feeds]
        m2.run_steps(4, 0.01, neg_ratio=1, seed=0, epoch=0, feeds=pinned, host_losses=hl)
        torch.cuda.synchronize()
        m2.flush()
        if hl.tolist() != my_losses or not torch.equal(m2.P.w, m.P.w) or not torch.equal(m2.q["w"].tensor, m.q["w"].tensor):
            print("FEED PIPELINE MISMATCH", kind, mode, hl.tolist(), my_losses)
            ok = False
        m2.close()
        m.close()
    # ---- the pointwise family (MF: dot, GMF: dot with the replicated vector h) over the same partition == the single-GPU pointwise
    # step on the union batch; h stays bit-identical on every rank
    from cleverrec_b200 import _lib as L
    from cleverrec_b200.dist import ShardedPointwise
    h0 = torch.randn(d, generator=g) * 0.3
    for score_kind, loss_kind in ((L.SCORE_DOT, L.LOSS_SQUARE), (L.SCORE_GMF, L.LOSS_CROSS_ENTROPY)):
        for kind, mode in (("SGD", "tf1"), ("Adagrad", "tf1"), ("Adam", "tf1")):
            lo, hi = user_range(U, rank, world)
            lr = 0.05 if kind != "Adam" else 0.01
            m = ShardedPointwise(eng, U, I, d, kind, lr, mode, B, kind=score_kind, loss_kind=loss_kind, init_P=P0[lo:hi], init_Q=Q0,
                                 init_h=h0 if score_kind == L.SCORE_GMF else None)
            ref = None
            if rank == 0:
                ref = (Table(P0.clone().cuda(), kind, mode), Table(Q0.clone().cuda(), kind, mode), Optimizer(kind, lr, adam_mode=mode))
                rh = h0.clone().cuda() if score_kind == L.SCORE_GMF else None
                rs1 = (torch.full_like(rh, 0.1) if kind == "Adagrad" else torch.zeros_like(rh)) if rh is not None and kind != "SGD" else None
                rs2 = torch.zeros_like(rh) if rh is not None and kind == "Adam" else None
            rs = np.random.RandomState(9)
            for step in range(4):
                feeds = []
                for r in range(world):
                    l, h_ = user_range(U, r, world)
                    n = B if step != 2 else 41
                    feeds.append((rs.randint(l, h_, n), rs.randint(0, I, n), (rs.rand(n) < 0.3).astype(np.float32)))
                u, i, y = feeds[rank]
                loss = m.step(0.01, feed=(u - lo, i, y))
                t = torch.tensor([loss], device="cuda", dtype=torch.float64)
                all_reduce_dev(t)
                if rank == 0:
                    uu, ii, yy = (np.concatenate([f[k] for f in feeds]) for k in range(3))
                    want = eng.train_step_pointwise(score_kind, ref[0], ref[1], ref[2], uu, ii, yy, 0.01, loss_kind, rh, rs1, rs2)
                    if abs(float(t.item()) - want) > 1e-5 * abs(want):
                        print("PW LOSS MISMATCH", score_kind, kind, step, float(t.item()), want)
                        ok = False
            m.check()
            m.flush()
            Qfull = m.gather_Q()
            Pfull = m.gather_P()
            hs = None
            if score_kind == L.SCORE_GMF:
                hs = [torch.zeros_like(m.h) for _ in range(world)]
                for r in range(world):
                    if r == rank:
                        hs[r].copy_(m.h)
                    broadcast_dev(hs[r], r)
            if rank == 0:
                eng.adam_flush(ref[0], ref[2]); eng.adam_flush(ref[1], ref[2])
                rtol, atol = (1e-4, 1e-5) if kind == "Adam" else (1e-5, 2e-7)
                checks = [("P", Pfull, ref[0].w), ("Q", Qfull, ref[1].w)]
                if hs is not None:
                    checks.append(("h", hs[0], rh))
                    if not all(torch.equal(hs[0], x) for x in hs[1:]):
                        print("PW h NOT REPLICATED", kind)
                        ok = False
                for name, got, want in checks:
                    bad = ~torch.isclose(got, want, rtol=rtol, atol=atol)
                    if bad.float().mean() > 1e-3:
                        print("PW TABLE MISMATCH", score_kind, kind, name, int(bad.sum()), float((got - want).abs().max()))
                        ok = False
            m.close()
    # ---- evaluation across the item shards: per-shard exact top-K merged at the owner == single-GPU top-K on the gathered tables
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from conftest import synthetic_data
    from cleverrec_b200.dist import ShardedEval, shard_history
    from cleverrec_b200.engine import history_from_dict
    data = synthetic_data(U, I, 9, seed=4)
    lo, hi = user_range(U, rank, world)
    m = ShardedBPR(eng, U, I, d, "Adagrad", 0.05, "tf1", B, init_P=P0[lo:hi], init_Q=Q0)
    mine, n_local = shard_history(data.ui_train, U, rank, world)
    m.set_history(mine, n_local)
    for step in range(3):   # device-sampled steps
        m.step(0.01, neg_ratio=2, seed=9, epoch=0, first=step * 64, batch=64)
    _, _, rp, sc = history_from_dict(mine, n_local)
    for exact, mode in ((True, "shard"), (False, "shard"), (True, "replicate"), (False, "replicate")):
        ev = ShardedEval(m, torch.from_numpy(rp), torch.from_numpy(sc), mode=mode)
        got = ev.topk(10, batch_users=20, exact=exact)
        Qfull = m.gather_Q()
        Pparts = [torch.zeros(user_range(U, r, world)[1] - user_range(U, r, world)[0], d, device="cuda") for r in range(world)]
        for r in range(world):
            if r == rank:
                Pparts[r].copy_(m.P.w)
            broadcast_dev(Pparts[r], r)
        ref_eng = Engine(local)
        ref_eng.set_history(data.ui_train, U, I)
        want = ref_eng.score_topk(0, torch.cat(Pparts), Qfull, torch.arange(lo, hi, dtype=torch.int32, device="cuda"), 10, exact=True)
        if not torch.equal(got, want):
            print("EVAL MISMATCH rank", rank, "exact", exact, mode, int((got != want).sum()))
            ok = False
        ref_eng.close()
    m.close()
    # ---- run_steps (next step's sampling / counting prepared on the auxiliary stream) == the same steps one by one
    outs = []
    for pipelined in (False, True):
        m = ShardedBPR(eng, U, I, d, "Adam", 0.01, "tf1", 64, init_P=P0[lo:hi], init_Q=Q0)
        m.set_history(mine, n_local)
        losses = torch.zeros(5, dtype=torch.float64, device="cuda")
        if pipelined:
            m.run_steps(5, 0.01, neg_ratio=2, seed=9, epoch=0, first=0, batch=64, loss_out=losses)
        else:
            for step in range(5):
                m.step(0.01, neg_ratio=2, seed=9, epoch=0, first=step * 64, batch=64, loss_out=losses[step:step + 1])
        m.flush()
        outs.append((losses.cpu().numpy().copy(), m.P.w.cpu().numpy().copy(), m.gather_Q().cpu().numpy().copy()))
        m.close()
    for k, name in enumerate(("loss", "P", "Q")):
        if not np.allclose(outs[0][k], outs[1][k], rtol=1e-6, atol=1e-8):
            print("RUN_STEPS MISMATCH rank", rank, name, float(np.abs(outs[0][k] - outs[1][k]).max()))
            ok = False
    # device-sampled path runs and stays finite
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    all_reduce_dev(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_OK" if flag.item() == 1.0 else "SHARDED_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
