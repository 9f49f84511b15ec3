"""Epoch time of BPR at the ml-1m shape through crb_train_epoch_bpr: captured whole-epoch graph vs the ordinary launch loop."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import bench_models as BM
from cleverrec_b200.engine import Engine, Optimizer, Table

data = BM.make_data(BM.ML1M, 1)
eng = Engine(0)
eng.set_history(data.ui_train, data.user_nums, data.item_nums)
d, B, R = 64, 6144, 4
rows = eng.epoch_rows(R)
n_steps = -(-rows // B)
for mode in ("graph", "loop", "loop-separate-tail", "graph-separate-tail"):
    os.environ["CRB_EPOCH_GRAPH"] = "1" if mode.startswith("graph") else "0"
    os.environ["CRB_DUP_TAIL"] = "0" if mode.endswith("separate-tail") else "1"
    g = torch.Generator().manual_seed(0)
    P = Table((torch.randn(data.user_nums, d, generator=g) * 0.01).cuda(), "Adam", "tf1")
    Q = Table((torch.randn(data.item_nums, d, generator=g) * 0.01).cuda(), "Adam", "tf1")
    opt = Optimizer("Adam", 1e-3)
    losses = torch.zeros(n_steps, dtype=torch.float64, device="cuda")
    ts = []
    for epoch in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.train_epoch_bpr(P, Q, opt, 0, epoch, 0, B, n_steps, R, 0.01, losses)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(mode, "steps", n_steps, "epoch ms", ["%.2f" % (1e3 * t) for t in ts], "loss", float(losses[-1]))

# the same epoch with the triplets sampled beforehand (device feeds): the auxiliary stream then only counts + assigns, which tells how
# much of the step the sampler's chain costs
os.environ["CRB_EPOCH_GRAPH"] = "0"
os.environ["CRB_DUP_TAIL"] = "1"
g = torch.Generator().manual_seed(0)
P = Table((torch.randn(data.user_nums, d, generator=g) * 0.01).cuda(), "Adam", "tf1")
Q = Table((torch.randn(data.item_nums, d, generator=g) * 0.01).cuda(), "Adam", "tf1")
opt = Optimizer("Adam", 1e-3)
u, i, j = eng.sample_pairwise(0, 0, 0, rows, R)
losses = torch.zeros(n_steps, dtype=torch.float64, device="cuda")
ts = []
for epoch in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.train_epoch_bpr_feeds(P, Q, opt, u, i, j, B, 0.01, losses)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print("device-feeds (no sampler)", "steps", n_steps, "epoch ms", ["%.2f" % (1e3 * t) for t in ts])
