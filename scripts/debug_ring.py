"""Experiment: the ring K3 kernel at growing batch sizes (does the failure start when the shared-memory ring wraps?)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleverrec_b200.engine import Engine, Optimizer, Table

eng = Engine(0)
U, I, d = 200000, 50000, int(os.environ.get("DIM", "128"))
kind = os.environ.get("OPT", "Adam")
for B in (3000, 6000, 7000, 20000, 100000, 262144):
    out = []
    for ring in ("0", "1"):
        os.environ["CRB_BPR_RING"] = ring
        g = torch.Generator().manual_seed(0)
        P = Table((torch.randn(U, d, generator=g) * 0.1).cuda(), kind, "tf1")
        Q = Table((torch.randn(I, d, generator=g) * 0.1).cuda(), kind, "tf1")
        opt = Optimizer(kind, 0.01)
        rs = np.random.RandomState(B)
        for step in range(3):
            u, i, j = rs.randint(0, U, B), rs.randint(0, I, B), rs.randint(0, I, B)
            loss = eng.train_step_bpr(P, Q, opt, u, i, j, 0.01)
        eng.adam_flush(P, opt); eng.adam_flush(Q, opt)
        torch.cuda.synchronize()
        out.append((loss, P.w.clone(), Q.w.clone()))
    print("B", B, "loss", out[0][0], out[1][0], "P equal", torch.equal(out[0][1], out[1][1]), "Q equal", torch.equal(out[0][2], out[1][2]),
          "max dQ", float((out[0][2] - out[1][2]).abs().max()), flush=True)
