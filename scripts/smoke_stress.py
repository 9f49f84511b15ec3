"""Runs the training part of __graft_entry__.smoke() repeatedly and prints the deviation from the TF-1 restatement each time
(is the smoke tolerance met with margin, and is the device result the same from run to run?)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleverrec_b200.engine import Engine, Optimizer, Table
from oracle import philox as X
from oracle import tf1_restatement as T

rng = np.random.default_rng(0)
U, I, d, R, B = 64, 500, 64, 4, 512
ui = {u: rng.choice(I, size=int(rng.integers(3, 20)), replace=False).tolist() for u in range(U)}
eng = Engine(0)
eng.set_history(ui, U, I)
pu, pi, rp, sc = X.build_history(ui, U)
g = torch.Generator().manual_seed(0)
P0, Q0 = torch.randn(U, d, generator=g) * 0.1, torch.randn(I, d, generator=g) * 0.1
ref, ropt = {"P": P0.clone(), "Q": Q0.clone()}, T.TF1Optimizer("Adam", 1e-2)
ru, ri, rj, _ = X.sample_pairwise(1, 0, 0, 2 * B, R, I, pu, pi, rp, sc)
for k in range(2):
    b = {n: torch.tensor(a[k * B:(k + 1) * B].astype(np.int64)) for n, a in (("u", ru), ("i", ri), ("j", rj))}
    T.train_step(T.bpr_loss, ref, b, {"reg": 0.01}, ropt, sparse_index={"P": ["u"], "Q": ["i", "j"]})
first = None
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    P, Q, opt = Table(P0.cuda(), "Adam"), Table(Q0.cuda(), "Adam"), Optimizer("Adam", 1e-2)
    losses = torch.zeros(2, dtype=torch.float64, device="cuda")
    eng.train_epoch_bpr(P, Q, opt, 1, 0, 0, B, 2, R, 0.01, losses)
    eng.adam_flush(P, opt); eng.adam_flush(Q, opt)
    torch.cuda.synchronize()
    gp, gq = P.w.cpu().numpy(), Q.w.cpu().numpy()
    if first is None:
        first = (gp.copy(), gq.copy())
    out = []
    for name, got, want in (("P", gp, ref["P"].numpy()), ("Q", gq, ref["Q"].numpy())):
        err = np.abs(got - want)
        tol = 2e-7 + 1e-5 * np.abs(want)
        out.append("%s max|d| %.2e worst err/tol %.2f n_bad %d" % (name, err.max(), (err / tol).max(), int((err > tol).sum())))
    print(rep, " | ".join(out), "| same as run 0:", np.array_equal(gp, first[0]) and np.array_equal(gq, first[1]), flush=True)
