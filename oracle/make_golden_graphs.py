"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates tests/golden/refgraph_steps.npz by EXECUTING the genuine reference model classes
(/root/reference/model/ranking/{BPR,GMF}.py, unmodified) on the TF-1 API shim (oracle/tf1_shim.py) in fp64: initial tables (fp32
values, so a device run starts from the same bits), the fed batches, the loss `sess.run([self.train, self.loss], feed)` returned
at every step and the variables after the last one, for SGD / Adagrad / Adam.  The GPU box has no /root/reference: the `-m gpu`
test tests/test_gpu_zz_golden_graphs.py replays these vectors through the C ABI.  Build container only:

    python -m oracle.make_golden_graphs
"""
import logging
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refimport as R  # noqa: E402
from oracle import tf1_shim as tf  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "refgraph_steps.npz")
U, I, D = 40, 60, 32
BATCHES = (64, 1, 257, 64)      # duplicates guaranteed (B > rows), a single triplet, an odd size
LR = {"SGD": 0.05, "Adagrad": 0.05, "Adam": 0.01}
REG = 0.01


class _Data(object):
    def __init__(self):
        self.user_nums, self.item_nums = U, I
        self.ui_train, self.ui_test = {u: [u % I] for u in range(U)}, {0: [1]}


def _model(classes, name, kind, **over):
    tf.reset_default_graph()
    tf.seed_initializers(1)
    cfg = R.default_configs(recommender=name, **dict({"init_method": "normal", "stddev": 0.1, "embed_size": D, "optimizer": kind,
                                                       "lr": LR[kind], "reg": REG, "data.split_way": "loo", "test.neg_samples": 5}, **over))
    sess = tf.Session()
    m = classes[name](sess, _Data(), cfg, logging.getLogger("golden"))
    m.build_model()
    return sess, m


def main():
    classes = tf.load_reference_models(R.REFERENCE_ROOT, ["BPR", "GMF", "MLP"])
    classes["GMF"].__init__.__globals__["get_loss"] = classes["MLP"].__init__.__globals__["get_loss"]   # GMF.py never imports it
    out = {"U": U, "I": I, "D": D, "reg": REG, "batches": np.asarray(BATCHES)}
    g = torch.Generator().manual_seed(2024)
    P0 = (torch.randn(U, D, generator=g) * 0.1).numpy().astype(np.float32)
    Q0 = (torch.randn(I, D, generator=g) * 0.1).numpy().astype(np.float32)
    h0 = (torch.randn(D, generator=g) * 0.5).numpy().astype(np.float32)
    out.update(P0=P0, Q0=Q0, h0=h0)
    rs = np.random.RandomState(7)
    feeds = []
    for k, B in enumerate(BATCHES):
        u, i, j = rs.randint(0, U, B).astype(np.int32), rs.randint(0, I, B).astype(np.int32), rs.randint(0, I, B).astype(np.int32)
        y = rs.randint(0, 2, B).astype(np.float32)
        feeds.append((u, i, j, y))
        out.update({"u%d" % k: u, "i%d" % k: i, "j%d" % k: j, "y%d" % k: y})
    for kind in ("SGD", "Adagrad", "Adam"):
        out["lr_" + kind] = LR[kind]
        # BPR.py:31-44
        sess, m = _model(classes, "BPR", kind)
        m.P.assign_value(P0); m.Q.assign_value(Q0)
        losses = [sess.run([m.train, m.loss], {m.u_idx: u, m.i_idx: i, m.j_idx: j})[1] for u, i, j, _ in feeds]
        out.update({"bpr_%s_loss" % kind: np.asarray(losses), "bpr_%s_P" % kind: m.P.numpy(), "bpr_%s_Q" % kind: m.Q.numpy()})
        # GMF.py:37-49, cross entropy
        sess, m = _model(classes, "GMF", kind, loss_func="cross_entropy")
        m.P.assign_value(P0); m.Q.assign_value(Q0); m.h_gmf.assign_value(h0)
        losses = [sess.run([m.train, m.loss], {m.u_idx: u, m.i_idx: i, m.y: y})[1] for u, i, _, y in feeds]
        out.update({"gmf_%s_loss" % kind: np.asarray(losses), "gmf_%s_P" % kind: m.P.numpy(), "gmf_%s_Q" % kind: m.Q.numpy(),
                    "gmf_%s_h" % kind: m.h_gmf.numpy()})
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
