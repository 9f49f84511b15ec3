import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_ROOT = os.environ.get("CRB_REFERENCE_ROOT", "/root/reference")   # read-only mount; absent on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) when no device is present and they were not deselected with -m "not gpu"
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def unflat(keys, lens, items):
    out, p = {}, 0
    for k, n in zip(keys.tolist(), lens.tolist()):
        out[k] = items[p:p + n].tolist()
        p += n
    return out


class Data(object):
    def __init__(self, user_nums, item_nums, ui_train, ui_test):
        self.user_nums, self.item_nums, self.ui_train, self.ui_test = user_nums, item_nums, ui_train, ui_test


def load_split(name):
    z = np.load(os.path.join(GOLDEN, name))
    return Data(int(z["user_nums"]), int(z["item_nums"]), unflat(z["train_keys"], z["train_lens"], z["train_items"]),
                unflat(z["test_keys"], z["test_lens"], z["test_items"]))


@pytest.fixture(scope="session")
def split_loo():
    return load_split("split_ml100k_loo.npz")


@pytest.fixture(scope="session")
def split_rs():
    return load_split("split_ml100k_rs.npz")


@pytest.fixture(scope="session")
def golden():
    return {f[:-4]: np.load(os.path.join(GOLDEN, f)) for f in os.listdir(GOLDEN) if f.endswith(".npz")}


def synthetic_data(n_users, n_items, mean_len, seed, test_per_user=1):
    """Small synthetic data object with the reference's attribute surface."""
    rng = np.random.default_rng(seed)
    ui_train, ui_test = {}, {}
    for u in range(n_users):
        if u % 11 == 7:  # users without history exist in the reference too (u not in ui_train)
            continue
        n = int(min(n_items - 30, max(1, rng.poisson(mean_len))))
        items = rng.choice(n_items, size=n + test_per_user, replace=False).tolist()
        ui_train[u] = items[:n]
        ui_test[u] = items[n:]
    return Data(n_users, n_items, ui_train, ui_test)
