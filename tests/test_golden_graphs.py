"""CPU side of the golden replay: the fixture made by executing the genuine reference BPR / GMF classes on the TF-1 shim
(oracle/make_golden_graphs.py) is (i) reproducible here when /root/reference is present, (ii) met by the fp32 restatement through the
SAME harness and tolerances the `-m gpu` test uses (tests/golden_replay.py) -- so a failure on the GPU box is the device's, not the
harness's."""
import numpy as np
import pytest
import torch

import golden_replay as GR
from oracle import refimport as R
from oracle import tf1_restatement as T


@pytest.mark.parametrize("kind", GR.KINDS)
@pytest.mark.parametrize("model", ["bpr", "gmf"])
def test_fp32_restatement_meets_the_golden_through_the_gpu_harness(model, kind):
    z = GR.load()
    ref = {"P": torch.tensor(z["P0"]), "Q": torch.tensor(z["Q0"])}
    if model == "gmf":
        ref["h"] = torch.tensor(z["h0"])
    opt = T.TF1Optimizer(kind, float(z["lr_" + kind]))
    reg = float(z["reg"])

    def step(k, u, i, j, y):
        t = lambda a: torch.tensor(a.astype(np.int64))      # noqa: E731
        if model == "bpr":
            return T.train_step(T.bpr_loss, ref, {"u": t(u), "i": t(i), "j": t(j)}, {"reg": reg}, opt, sparse_index={"P": ["u"], "Q": ["i", "j"]})
        return T.train_step(T.gmf_loss, ref, {"u": t(u), "i": t(i), "y": torch.tensor(y)}, {"reg": reg, "loss_func": "cross_entropy"}, opt,
                            sparse_index={"P": ["u"], "Q": ["i"]})
    out = GR.replay(z, model, kind, step, lambda: {k: v.numpy() for k, v in ref.items()})
    assert set(out) == ({"P", "Q"} if model == "bpr" else {"P", "Q", "h"})
    assert float(np.abs(out["P"] - z["P0"]).max()) > 1e-4           # the tables were trained


def test_harness_rejects_a_wrong_result():
    z = GR.load()
    with pytest.raises(AssertionError):
        GR.check_table(z["bpr_SGD_P"].astype(np.float32) * (1 + 2e-4), z["bpr_SGD_P"], "SGD", 0.05, "P")
    with pytest.raises(AssertionError):
        GR.check_table(z["bpr_Adam_P"].astype(np.float32) + 1e-3, z["bpr_Adam_P"], "Adam", 0.01, "P")
    with pytest.raises(AssertionError):
        GR.check_loss(1.0001, 1.0)


@pytest.mark.skipif(not R.available(), reason="/root/reference not present")
def test_fixture_is_what_the_genuine_reference_classes_produce_today(tmp_path, monkeypatch):
    from oracle import make_golden_graphs as M
    monkeypatch.setattr(M, "OUT", str(tmp_path / "again.npz"))
    M.main()
    a, b = GR.load(), np.load(str(tmp_path / "again.npz"))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k
