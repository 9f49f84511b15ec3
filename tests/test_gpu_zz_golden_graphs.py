"""-m gpu: golden vectors of the reference itself.  tests/golden/refgraph_steps.npz was produced by EXECUTING the genuine,
unmodified reference classes model/ranking/BPR.py and GMF.py on the TF-1 API shim in fp64 (oracle/make_golden_graphs.py, build
container); here the same initial tables and the same four feeds go through the C ABI (crb_train_step_bpr /
crb_train_step_pointwise, host feeds, host loss) and must give the reference's per-step losses and final variables.  Harness and
tolerances are tests/golden_replay.py, which the CPU suite exercises with the fp32 restatement in the device's place.
(File name: collected last among the -m gpu files -- it was added after the round's GPU budget was spent.)"""
import numpy as np
import pytest
import torch

import golden_replay as GR

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cleverrec_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("kind", GR.KINDS)
def test_bpr_steps_reproduce_the_reference_graph_golden(eng, kind):
    from cleverrec_b200.engine import Optimizer, Table
    z = GR.load()
    opt = Optimizer(kind, float(z["lr_" + kind]), adam_mode="tf1")
    P, Q = Table(torch.tensor(z["P0"]).cuda(), kind, "tf1"), Table(torch.tensor(z["Q0"]).cuda(), kind, "tf1")
    reg = float(z["reg"])

    def step(k, u, i, j, y):
        return eng.train_step_bpr(P, Q, opt, u, i, j, reg=reg)

    def read():
        eng.adam_flush(P, opt)
        eng.adam_flush(Q, opt)
        torch.cuda.synchronize()
        return {"P": P.w.cpu().numpy(), "Q": Q.w.cpu().numpy()}
    GR.replay(z, "bpr", kind, step, read)


@pytest.mark.parametrize("kind", GR.KINDS)
def test_gmf_steps_reproduce_the_reference_graph_golden(eng, kind):
    from cleverrec_b200 import _lib
    from cleverrec_b200.engine import Optimizer, Table
    z = GR.load()
    opt = Optimizer(kind, float(z["lr_" + kind]), adam_mode="tf1")
    P, Q = Table(torch.tensor(z["P0"]).cuda(), kind, "tf1"), Table(torch.tensor(z["Q0"]).cuda(), kind, "tf1")
    hd = torch.tensor(z["h0"]).cuda()
    s1 = torch.full_like(hd, 0.1) if kind == "Adagrad" else (torch.zeros_like(hd) if kind == "Adam" else None)
    s2 = torch.zeros_like(hd) if kind == "Adam" else None
    reg = float(z["reg"])

    def step(k, u, i, j, y):
        return eng.train_step_pointwise(_lib.SCORE_GMF, P, Q, opt, u, i, y, reg, _lib.LOSS_CROSS_ENTROPY, hd, s1, s2)

    def read():
        eng.adam_flush(P, opt)
        eng.adam_flush(Q, opt)
        torch.cuda.synchronize()
        return {"P": P.w.cpu().numpy(), "Q": Q.w.cpu().numpy(), "h": hd.cpu().numpy()}
    GR.replay(z, "gmf", kind, step, read)
