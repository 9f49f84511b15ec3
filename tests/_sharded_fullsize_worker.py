"""Worker for tests/test_gpu_sharded_one_device.py::test_sharded_step_at_full_size (torch.distributed.run, world 2, CRB_SHARED_DEVICE=1):
the multi-GPU step at BASELINE.json's full table sizes (10M users partitioned, 2M items row-sharded, d = 128, 2^20 triplets per rank
and step) checked through properties, the way tests/test_gpu_fullsize.py checks the single-GPU step:
  * one SGD step over peer memory == a plain-torch fp32 formulation of BPR.py:31-44 on the union of the ranks' triplets (loss, every
    touched row of P and of the gathered Q; rows outside the batch bit-unchanged);
  * lr = 0 leaves every table bit-unchanged (the fetch / send / inbox plumbing moves nothing by itself);
  * the device-sampled step (run_steps) == the step fed with the same window's triplets from sample_pairwise (same indices);
  * two TF-1 Adam steps: the sum of the ranks' losses equals the torch loss on the union batch at the torch-updated tables.
The torch formulation is the 'plain PyTorch fp32 reference' of the op; bit-exact small-size parity lives in _sharded_worker.py."""
import math
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import build_history_device  # noqa: E402
from cleverrec_b200.dist import ShardedBPR, all_reduce_dev, broadcast_dev, user_range  # noqa: E402
from cleverrec_b200.engine import Engine  # noqa: E402
from test_gpu_fullsize import _bpr_torch  # noqa: E402

USERS, ITEMS, DIM, MEAN_HIST, B, R = 10_000_000, 2_000_000, 128, 40, 1 << 20, 4


def gather_parts(x, rank, world, shapes):
    """Every rank's tensor on every rank (broadcast_dev works over gloo with ranks sharing a device)."""
    parts = []
    for r in range(world):
        t = x.clone() if r == rank else torch.empty(shapes[r], dtype=x.dtype, device=x.device)
        broadcast_dev(t, r)
        parts.append(t)
    return parts


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    shared = os.environ.get("CRB_SHARED_DEVICE", "0") == "1"
    local = 0 if shared else int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    free, _ = torch.cuda.mem_get_info(local)
    if free < (60 if shared else 40) * (1 << 30):
        if rank == 0:
            print("SHARDED_FULLSIZE_SKIP not enough free HBM")
        dist.destroy_process_group()
        return
    eng = Engine(local)
    lo, hi = user_range(USERS, rank, world)
    n_local = hi - lo
    pu, pi, rowptr = build_history_device(torch, dev, n_local, ITEMS, MEAN_HIST, seed=77 + rank)
    eng.set_history_arrays(n_local, ITEMS, pu, pi, rowptr, pi)
    del pu, pi, rowptr
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    P0 = torch.randn(n_local, DIM, device=dev, generator=g) * 0.1
    gq = torch.Generator(device=dev).manual_seed(99)                  # the same full item table on every rank; each keeps its rows
    Q0 = torch.randn(ITEMS, DIM, device=dev, generator=gq) * 0.1
    ok = True

    def fail(*a):
        nonlocal ok
        ok = False
        print("FULLSIZE MISMATCH rank", rank, *a, flush=True)

    lr, reg, seed, first = 0.05, 0.01, 21, 54321
    u, i, j = eng.sample_pairwise(seed, 0, first, B, R)                # u: local rows of this rank's partition
    assert int(u.min()) >= 0 and int(u.max()) < n_local and int(j.max()) < ITEMS

    # ---- lr = 0: nothing moves ----
    m = ShardedBPR(eng, USERS, ITEMS, DIM, "SGD", 0.0, "tf1", B, init_P=P0.clone(), init_Q=Q0)
    m.step(reg, feed=(u, i, j))
    m.check()
    if not torch.equal(m.P.w, P0) or not torch.equal(m.q["w"].tensor, Q0[rank::world]):
        fail("lr=0 moved a table")
    m.close()

    # ---- one SGD step, fed ----
    m = ShardedBPR(eng, USERS, ITEMS, DIM, "SGD", lr, "tf1", B, init_P=P0.clone(), init_Q=Q0)
    loss = m.step(reg, feed=(u, i, j))
    m.check()
    # ---- the same window sampled on the device inside run_steps ----
    m2 = ShardedBPR(eng, USERS, ITEMS, DIM, "SGD", lr, "tf1", B, init_P=P0.clone(), init_Q=Q0)
    l2 = torch.zeros(1, dtype=torch.float64, device=dev)
    m2.run_steps(1, reg, neg_ratio=R, seed=seed, epoch=0, first=first, batch=B, loss_out=l2)
    m2.check()
    if abs(float(l2.item()) - loss) > 1e-9 * abs(loss):
        fail("device-sampled loss", float(l2.item()), loss)
    # (rows whose duplicates arrive in another order may differ in the last bits: tolerance, not bits)
    if (m2.P.w - m.P.w).abs().max().item() > 1e-7 or (m2.q["w"].tensor - m.q["w"].tensor).abs().max().item() > 1e-7:
        fail("device-sampled tables")
    m2.close()

    t = torch.tensor([loss], device=dev, dtype=torch.float64)
    all_reduce_dev(t)
    Qgot = m.gather_Q()
    # P rows: this rank's partition against torch on its own triplets (P's gradient only involves the rank's own batch)
    want_local, GP, _ = _bpr_torch(P0, Q0, u, i, j, reg)
    if abs(loss - want_local) > 2e-6 * abs(want_local):
        fail("rank loss", loss, want_local)
    refP = P0 - lr * GP
    del GP
    err = (m.P.w - refP).abs().max().item()
    if err > 2e-7:
        fail("P rows", err)
    touched = torch.zeros(n_local, dtype=torch.bool, device=dev)
    touched[u.long()] = True
    changed = (m.P.w != P0).any(1)
    if bool((changed & ~touched).any()) or int(changed.sum()) < 0.999 * int(touched.sum()):
        fail("P untouched rows", int((changed & ~touched).sum()), int(changed.sum()), int(touched.sum()))
    del refP, touched, changed
    # Q rows: the owners summed both ranks' gradients -- torch on the union batch, evaluated by rank 0
    shapes = [(B,)] * world
    iu, ju = torch.cat(gather_parts(i, rank, world, shapes)), torch.cat(gather_parts(j, rank, world, shapes))
    if rank == 0:
        GQ = torch.zeros_like(Q0)
    for r in range(world):                                              # the P rows each rank's triplets read, one rank at a time
        l_, h_ = user_range(USERS, r, world)
        pr = P0[u.long()].clone() if r == rank else torch.empty(B, DIM, device=dev)
        broadcast_dev(pr, r)
        if rank == 0:
            ir, jr = iu[r * B:(r + 1) * B].long(), ju[r * B:(r + 1) * B].long()
            qi, qj = Q0[ir], Q0[jr]
            x = (pr * qi).sum(1) - (pr * qj).sum(1)
            gg = -torch.sigmoid(-x)[:, None]
            GQ.index_add_(0, ir, gg * pr + reg * qi).index_add_(0, jr, -gg * pr + reg * qj)
        del pr
    if rank == 0:
        refQ = Q0 - lr * GQ
        err = (Qgot - refQ).abs().max().item()
        if err > 2e-7:
            fail("Q rows", err)
        touched = torch.zeros(ITEMS, dtype=torch.bool, device=dev)
        touched[iu.long()] = True
        touched[ju.long()] = True
        changed = (Qgot != Q0).any(1)
        if bool((changed & ~touched).any()) or int(changed.sum()) < 0.999 * int(touched.sum()):
            fail("Q untouched rows", int((changed & ~touched).sum()), int(changed.sum()), int(touched.sum()))
        del GQ, refQ, touched, changed
    m.close()
    del Qgot

    # ---- two TF-1 Adam steps: the second step's loss is evaluated at tables the first step moved (every row of Q decays / moves
    # under TF-1 semantics); torch: dense Adam on rank-local P and on the full Q with the union gradient ----
    m = ShardedBPR(eng, USERS, ITEMS, DIM, "Adam", 1e-3, "tf1", B, init_P=P0.clone(), init_Q=Q0)
    b1, b2, eps, alr = 0.9, 0.999, 1e-8, 1e-3
    Pr, Qr = P0.clone(), Q0.clone()
    mP, vP, mQ, vQ = torch.zeros_like(P0), torch.zeros_like(P0), torch.zeros_like(Q0), torch.zeros_like(Q0)
    for step in (1, 2):
        u, i, j = eng.sample_pairwise(seed + 1, 0, (step - 1) * B, B, R)
        loss = m.step(reg, feed=(u, i, j))
        want, GP, GQ = _bpr_torch(Pr, Qr, u, i, j, reg)
        if abs(loss - want) > 5e-6 * abs(want):
            fail("adam loss", step, loss, want)
        all_reduce_dev(GQ)                                               # both ranks' item gradients (the owners' sums)
        lr_t = alr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        for w, mm, vv, G in ((Pr, mP, vP, GP), (Qr, mQ, vQ, GQ)):
            mm.mul_(b1).add_(G, alpha=1 - b1)
            vv.mul_(b2).addcmul_(G, G, value=1 - b2)
            w.sub_(lr_t * mm / (vv.sqrt() + eps))
        del GP, GQ
    m.check()
    m.flush()
    Qgot = m.gather_Q()
    for got, ref, name in ((m.P.w, Pr, "P"), (Qgot, Qr, "Q")):
        err = (got - ref).abs()
        # the bar of tests/test_gpu_fullsize.py::test_two_adam_steps_equal_dense_tf1_adam (Adam's division is ill-conditioned where a
        # gradient entry cancels to ~eps): >= 99.9 % of the entries within 2e-6, none off by more than 3e-4
        if (err > 2e-6).float().mean().item() >= 1e-3 or err.max().item() > 3e-4:
            fail("adam table", name, (err > 2e-6).float().mean().item(), err.max().item())
    m.close()

    flag = torch.tensor([0 if ok else 1], device=dev, dtype=torch.float64)
    all_reduce_dev(flag)
    if rank == 0 and float(flag.item()) == 0:
        print("SHARDED_FULLSIZE_OK")
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
