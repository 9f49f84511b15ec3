# coding: utf-8
""" Sample instances for training -- drop-in signatures of the reference utils/sampler.py:10-99.

The reference materialises an epoch with Python loops over np.random; here the epoch is drawn by the CUDA sampler
(cleverrec_b200/csrc/sampler.cu) and copied back so callers get the same tuple of NumPy int arrays.  The model
classes do not use these functions in their training loop (they call the fused sample+train entry point); they
exist so code written against the reference's sampler API keeps working."""
import math

import numpy as np

from ..engine import Engine

_state = {"engine": None, "data_id": None, "epoch": 0, "seed": 0, "mode": "philox"}


def set_seed(seed):
    _state["seed"] = int(seed)
    _state["epoch"] = 0


def set_mode(mode):
    """'philox' (default): the counter-based device sampler, same distribution as the reference, its own stream.
    'numpy_stream': the reference's output BIT FOR BIT -- the device adopts NumPy's current global RandomState
    (np.random.get_state()), draws the epoch exactly as utils/sampler.py would, and hands the advanced state back to NumPy,
    so `np.random.seed(s); pairwise_ranking_sampler(...)` returns the reference's arrays."""
    if mode not in ("philox", "numpy_stream"):
        raise ValueError("sampler mode must be 'philox' or 'numpy_stream'")
    _state["mode"] = mode


def _numpy_epoch(eng, kind, neg_ratio, with_nbr=False):
    eng.np_set_state()
    out = eng.sample_epoch_numpy(kind, neg_ratio, with_nbr=with_nbr) if kind == "pairwise" else eng.sample_epoch_numpy(kind, neg_ratio)
    np.random.set_state(eng.np_get_state())
    return out


def _engine_for(data):
    if _state["engine"] is None:
        _state["engine"] = Engine(0)
    if _state["data_id"] != id(data):
        _state["engine"].set_history(data.ui_train, data.user_nums, data.item_nums)
        _state["data_id"] = id(data)
    return _state["engine"]


def _next_epoch():
    e = _state["epoch"]
    _state["epoch"] += 1
    return e


# Get training instances for pointwise learning
def pointwise_ranking_sampler(data, neg_ratio, batch_size, fism_like=False):
    eng = _engine_for(data)
    n = eng.epoch_rows(neg_ratio, "pointwise")
    if _state["mode"] == "numpy_stream":
        out = _numpy_epoch(eng, "pointwise", neg_ratio)
        return (math.ceil(n / batch_size), out[0].cpu().numpy().astype(np.int64), out[1].cpu().numpy().astype(np.int64),
                out[2].cpu().numpy().astype(np.float64))
    out = eng.sample_pointwise(_state["seed"], _next_epoch(), 0, n, neg_ratio, with_nbr=fism_like)
    res = (math.ceil(n / batch_size), out[0].cpu().numpy().astype(np.int64), out[1].cpu().numpy().astype(np.int64),
           out[2].cpu().numpy().astype(np.float64))
    if fism_like:
        res = res + (out[3].cpu().numpy().astype(np.int64),)
    return res


# Get training instances for pairwise learning
def pairwise_ranking_sampler(data, neg_ratio, batch_size, fism_like=False):
    eng = _engine_for(data)
    n = eng.epoch_rows(neg_ratio, "pairwise")
    if _state["mode"] == "numpy_stream":
        out = _numpy_epoch(eng, "pairwise", neg_ratio, with_nbr=fism_like)
    else:
        out = eng.sample_pairwise(_state["seed"], _next_epoch(), 0, n, neg_ratio, with_nbr=fism_like)
    res = (math.ceil(n / batch_size),) + tuple(t.cpu().numpy().astype(np.int64) for t in out[:3])
    if fism_like:
        res = res + (out[3].cpu().numpy().astype(np.int64),)
    return res


# For CML
def ranking_sampler_cml(data, neg_ratio, batch_size):
    eng = _engine_for(data)
    n = eng.epoch_rows(neg_ratio, "cml")
    if _state["mode"] == "numpy_stream":
        u, i, neg = _numpy_epoch(eng, "cml", neg_ratio)
    else:
        u, i, neg = eng.sample_cml(_state["seed"], _next_epoch(), 0, n, neg_ratio)
    return (math.ceil(n / batch_size), u.cpu().numpy().astype(np.int64), i.cpu().numpy().astype(np.int64),
            neg.cpu().numpy().astype(np.int64))


# For SBPR
def ranking_sampler_sbpr(data, SPu, neg_ratio, batch_size, is_suk=True):
    """utils/sampler.py:102-141: (train_batches, u, i, i_s, i_neg[, suk]).  numpy_stream mode returns the reference's arrays bit for
    bit under NumPy's current global stream (the path SBPR.train_model_sbpr takes with sampler=numpy_stream)."""
    eng = _engine_for(data)
    if _state.get("social_id") != (id(data), id(SPu)):
        eng.set_social(data.ui_train, data.user_friends, SPu, data.user_nums)
        _state["social_id"] = (id(data), id(SPu))
    n = eng.epoch_rows(neg_ratio, "sbpr")
    if _state["mode"] == "numpy_stream":
        eng.np_set_state()
        out = eng.sample_epoch_numpy_sbpr(neg_ratio, is_suk=is_suk)
        np.random.set_state(eng.np_get_state())
    else:
        out = eng.sample_sbpr(_state["seed"], _next_epoch(), 0, n, neg_ratio, is_suk=is_suk)
    res = (math.ceil(n / batch_size),) + tuple(t.cpu().numpy().astype(np.int64) for t in out[:4])
    if is_suk:
        res = res + (out[4].cpu().numpy().astype(np.int64),)
    return res
