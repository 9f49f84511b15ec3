"""ORACLE / TEST INFRASTRUCTURE ONLY.  ctypes wrapper of oracle/crb_oracle.c (canonical-order scoring + top-K)."""
import ctypes as C

import numpy as np

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build())
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def score_pairs(kind, P, Q, u, i, hvec=None):
    P, Q = np.ascontiguousarray(P, np.float32), np.ascontiguousarray(Q, np.float32)
    u, i = np.ascontiguousarray(u, np.int32), np.ascontiguousarray(i, np.int32)
    hvec = np.ascontiguousarray(hvec, np.float32) if hvec is not None else None
    out = np.empty(u.shape[0], np.float32)
    lib().oracle_score_pairs(C.c_int(kind), _p(P), _p(Q), _p(hvec), C.c_int(P.shape[1]), _p(u), _p(i), C.c_int64(u.shape[0]), _p(out))
    return out


def topk_segments(scores, offsets, K, ascending=False):
    scores, offsets = np.ascontiguousarray(scores, np.float32), np.ascontiguousarray(offsets, np.int64)
    n = offsets.shape[0] - 1
    out = np.empty((n, K), np.int32)
    lib().oracle_topk_segments(_p(scores), _p(offsets), C.c_int64(n), C.c_int(K), C.c_int(1 if ascending else 0), _p(out))
    return out


def fullrank_topk(kind, P, Q, users, seen_rowptr, seen_cols, K, hvec=None, hist_users=None, n_items=None):
    P, Q = np.ascontiguousarray(P, np.float32), np.ascontiguousarray(Q, np.float32)
    users = np.ascontiguousarray(users, np.int32)
    hist_users = np.ascontiguousarray(hist_users, np.int32) if hist_users is not None else None
    hvec = np.ascontiguousarray(hvec, np.float32) if hvec is not None else None
    seen_rowptr, seen_cols = np.ascontiguousarray(seen_rowptr, np.int64), np.ascontiguousarray(seen_cols, np.int32)
    n_items = Q.shape[0] if n_items is None else n_items
    items = np.empty((users.shape[0], K), np.int32)
    scores = np.empty((users.shape[0], K), np.float32)
    lib().oracle_fullrank_topk(C.c_int(kind), _p(P), _p(Q), _p(hvec), C.c_int64(n_items), C.c_int(P.shape[1]), _p(users), _p(hist_users),
                               C.c_int64(users.shape[0]), _p(seen_rowptr), _p(seen_cols), C.c_int(K), _p(items), _p(scores))
    return items, scores


def score_pairs_neumf(Pg, Qg, Pm, Qm, dense, n_layers, u, i):
    """Canonical NeuMF (Pg / Qg given) or MLP (Pg = Qg = None) logit of flattened pairs; `dense` = packed W_l, b_l, h."""
    Pm, Qm = np.ascontiguousarray(Pm, np.float32), np.ascontiguousarray(Qm, np.float32)
    E = 0 if Pg is None else Pg.shape[1]
    if E:
        Pg, Qg = np.ascontiguousarray(Pg, np.float32), np.ascontiguousarray(Qg, np.float32)
    dense = np.ascontiguousarray(dense, np.float32)
    u, i = np.ascontiguousarray(u, np.int32), np.ascontiguousarray(i, np.int32)
    out = np.empty(u.shape[0], np.float32)
    lib().oracle_score_pairs_neumf(_p(Pg) if E else None, _p(Qg) if E else None, _p(Pm), _p(Qm), _p(dense), C.c_int(E), C.c_int(Pm.shape[1]),
                                   C.c_int(n_layers), _p(u), _p(i), C.c_int64(u.shape[0]), _p(out))
    return out
