"""Shared by tests/test_golden_graphs.py (CPU: harness + fixture + tolerances, with the fp32 restatement standing where the device
stands) and tests/test_gpu_zz_golden_graphs.py (-m gpu: the same replay through the C ABI).  The fixture tests/golden/refgraph_steps.npz
holds what the GENUINE reference BPR / GMF classes produced when executed on the TF-1 shim in fp64 (oracle/make_golden_graphs.py)."""
import os

import numpy as np

from conftest import GOLDEN

KINDS = ("SGD", "Adagrad", "Adam")
LOSS_RTOL = 4e-5


def load():
    return np.load(os.path.join(GOLDEN, "refgraph_steps.npz"))


def feeds(z, k):
    return z["u%d" % k], z["i%d" % k], z["j%d" % k], z["y%d" % k]


def check_loss(got, want):
    assert abs(got - want) <= LOSS_RTOL * abs(want), (got, want)


def check_table(got, want, kind, lr, name, model="bpr"):
    """got: fp32 result; want: the fp64 golden.  The criteria have the FORM of the existing device-vs-restatement tests
    (tests/test_gpu_train_bpr.py::_compare, tests/test_gpu_train_pointwise.py) with slightly wider numbers, because the golden is
    the fp64 truth and not another fp32 run:
      bpr, SGD / Adagrad: every entry within north_star's 1e-4 relative;
      bpr, Adam: deviation D1 of DESIGN.md section 4 -- all but a vanishing fraction of the entries within 1e-4, none further
                 than 5 % of one step;
      gmf tables: <= 0.2 % of the entries outside the band, none further than 5 % of one step; gmf's dense h: every entry."""
    got = np.asarray(got, dtype=np.float64)
    assert got.shape == want.shape, name
    if model == "gmf" and name == "h":
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=4e-6, err_msg=name)
        return
    if model == "bpr" and kind != "Adam":
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-6, err_msg=name)
        return
    rtol, atol = (1e-4, 4e-5) if kind == "Adam" else (2e-5, 2e-6)
    if model == "bpr":
        rtol, atol = 1e-4, 2e-5
    bad = ~np.isclose(got, want, rtol=rtol, atol=atol)
    assert bad.mean() <= 2e-3, (name, int(bad.sum()))
    assert np.abs(got - want).max() <= 0.05 * lr, (name, float(np.abs(got - want).max()))


def replay(z, model, kind, step, read):
    """step(k, u, i, j, y) -> loss of step k; read() -> {name: array} after the last step."""
    n = int(z["batches"].shape[0])
    for k in range(n):
        u, i, j, y = feeds(z, k)
        assert len(u) == int(z["batches"][k])
        check_loss(step(k, u, i, j, y), float(z["%s_%s_loss" % (model, kind)][k]))
    out = read()
    lr = float(z["lr_" + kind])
    for name, got in out.items():
        check_table(got, z["%s_%s_%s" % (model, kind, name)], kind, lr, name, model)
    return out
