"""Thin Python object layer over the C ABI: device memory is owned by torch tensors, every numeric operation is
one call into libcleverrec_b200.so.  This is the replacement for the reference's `tf.Session` (main.py:39-45):
the mirror classes under cleverrec_b200/model call it where the reference calls `self.sess.run`."""
import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import CrbOpt, CrbTable, check, ptr


class Table(object):
    """One TF variable [rows, dim] + its optimizer slot variables, on the device."""

    def __init__(self, w, optimizer, adam_mode="tf1"):
        assert w.is_cuda and w.dtype == torch.float32 and w.dim() == 2 and w.is_contiguous()
        self.w = w
        self.s1 = self.s2 = self.last = None
        if optimizer == "Adagrad":
            self.s1 = torch.full_like(w, 0.1)  # tf.train.AdagradOptimizer initial_accumulator_value
        elif optimizer == "Adam":
            self.s1, self.s2 = torch.zeros_like(w), torch.zeros_like(w)
            if adam_mode == "tf1":
                self.last = torch.zeros(w.shape[0], dtype=torch.int32, device=w.device)
        self.c = CrbTable(ptr(self.w), ptr(self.s1), ptr(self.s2), ptr(self.last), w.shape[0], w.shape[1], 0)

    @property
    def rows(self):
        return self.w.shape[0]

    @property
    def dim(self):
        return self.w.shape[1]


class Optimizer(object):
    """utils/tools.py:79-87 get_optimizer with TF-1 defaults; `t` counts applied steps."""
    KINDS = {"SGD": _lib.OPT_SGD, "Adagrad": _lib.OPT_ADAGRAD, "Adam": _lib.OPT_ADAM}

    def __init__(self, kind, lr, adam_mode="tf1", beta1=0.9, beta2=0.999, eps=1e-8):
        if kind not in self.KINDS:
            raise ValueError("optimizer must be one of SGD/Adam/Adagrad, got %r" % (kind,))
        if adam_mode not in ("tf1", "lazy"):
            raise ValueError("adam_mode must be 'tf1' or 'lazy'")
        self.kind, self.lr, self.adam_mode = kind, float(lr), adam_mode
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        self.t = 0

    def c(self, step):
        return CrbOpt(self.KINDS[self.kind], _lib.ADAM_TF1 if self.adam_mode == "tf1" else _lib.ADAM_LAZY, self.lr, self.beta1,
                      self.beta2, self.eps, step)


class Engine(object):
    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("cleverrec_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        check(self.lib.crb_create(device, C.byref(h)))
        self.h = h
        self._hist = None
        self.n_users = self.n_items = self.n_pos = 0

    def close(self):
        if self.h:
            self.lib.crb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def launches(self):
        return int(self.lib.crb_launch_count(self.h))

    def profile(self, on):
        check(self.lib.crb_profile_enable(self.h, 1 if on else 0))

    def profile_read(self, tag=0):
        """-> (milliseconds spent in the bracketed kernel, launches) since the last read.  tag 0 = the fused step kernel,
        1 = staged fetch (multi-GPU), 2 = duplicate reduce / send, 3 = the owner's inbox pass (multi-GPU)."""
        ms, n = C.c_double(), C.c_int64()
        check(self.lib.crb_profile_read_tag(self.h, tag, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ------------------------------------------------------------------ history
    def set_history_arrays(self, n_users, n_items, pos_user, pos_item, seen_rowptr, seen_cols):
        """Arrays as produced by history_from_dict (NumPy or torch, host or device)."""
        def dev(a, dt):
            t = torch.as_tensor(a)
            return t.to(device=self.device, dtype=dt).contiguous()
        pu, pi = dev(pos_user, torch.int32), dev(pos_item, torch.int32)
        rp, sc = dev(seen_rowptr, torch.int64), dev(seen_cols, torch.int32)
        if sc.numel() == 0:
            sc = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._hist = (pu, pi, rp, sc)  # keep alive: the library borrows the pointers
        self.n_users, self.n_items, self.n_pos = int(n_users), int(n_items), int(pu.numel())
        check(self.lib.crb_set_history(self.h, n_users, n_items, self.n_pos, ptr(pu) if self.n_pos else None,
                                       ptr(pi) if self.n_pos else None, ptr(rp), ptr(sc), self.stream))

    def build_history(self, users, items, n_users, n_items):
        """The training split's (user, item) rows -- int32 columns in file order, NumPy or torch, host or device -- straight to the
        device history (crb_build_history): no dict of lists, no Python sets.  Equivalent to
        set_history(train.groupby('u_id').i_id.apply(list).to_dict(), ...) (model/RankingPreprocess.py:117)."""
        def col(a):
            if isinstance(a, torch.Tensor):
                return a.to(dtype=torch.int32).contiguous()
            return np.ascontiguousarray(a, dtype=np.int32)
        users, items = col(users), col(items)
        n = int(len(users))
        dev = self.device
        pu, pi = torch.empty(max(n, 1), dtype=torch.int32, device=dev), torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        rp, sc = torch.empty(n_users + 1, dtype=torch.int64, device=dev), torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        start, ln = torch.empty(n_users + 1, dtype=torch.int64, device=dev), torch.empty(n_users, dtype=torch.int32, device=dev)
        n_seen = C.c_int64(0)
        check(self.lib.crb_build_history(self.h, ptr(users), ptr(items), n, n_users, n_items, ptr(pu), ptr(pi), ptr(rp), ptr(sc),
                                         C.byref(n_seen), ptr(start), ptr(ln), self.stream))
        sc = sc[:max(int(n_seen.value), 1)]
        self._hist = (pu[:n], pi[:n], rp, sc)
        self.n_users, self.n_items, self.n_pos = int(n_users), int(n_items), n
        check(self.lib.crb_set_history(self.h, n_users, n_items, n, ptr(pu) if n else None, ptr(pi) if n else None, ptr(rp), ptr(sc), self.stream))
        self._lists = (start, ln)
        check(self.lib.crb_set_history_lists(self.h, ptr(start), ptr(ln)))
        return self._hist

    # ------------------------------------------------------------------ preprocessing on the device (csrc/preprocess.cu)
    def prep_filter_reindex(self, raw_u, raw_i, user_min=0, item_min=0):
        """RankingPreprocess._filter_users / _filter_items / re_index on two id columns (any integer dtype, host or device).
        -> dict(u, i: int32 new ids of the kept rows in file order; row: their original row numbers; user_ids / item_ids: raw id of
        each new id; n_users, n_items)."""
        dev = self.device
        def col(a):
            if not isinstance(a, torch.Tensor):
                a = torch.from_numpy(np.array(a, dtype=np.int64))   # a copy: pandas hands out read-only views
            return a.to(device=dev, dtype=torch.int64).contiguous()
        ru, ri = col(raw_u), col(raw_i)
        n = int(ru.numel())
        u, i = torch.empty(max(n, 1), dtype=torch.int32, device=dev), torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        row, uid, iid = (torch.empty(max(n, 1), dtype=torch.int64, device=dev) for _ in range(3))
        nk, nu, ni = C.c_int64(), C.c_int64(), C.c_int64()
        check(self.lib.crb_prep_filter_reindex(self.h, ptr(ru), ptr(ri), n, int(user_min), int(item_min), ptr(u), ptr(i), ptr(row), C.byref(nk),
                                               C.byref(nu), C.byref(ni), ptr(uid), ptr(iid), self.stream))
        k = int(nk.value)
        return {"u": u[:k], "i": i[:k], "row": row[:k], "user_ids": uid[:int(nu.value)], "item_ids": iid[:int(ni.value)],
                "n_users": int(nu.value), "n_items": int(ni.value)}

    def prep_split_loo(self, u, n_users, time=None):
        """The leave-one-out split: -> (perm int64 [n], is_test bool [n]) in the reference's enumeration order (ascending user,
        then time / file order); is_test marks a user's last row when the user has more than 3 rows."""
        dev = self.device
        u = torch.as_tensor(u).to(device=dev, dtype=torch.int32).contiguous()
        t = None if time is None else torch.as_tensor(time).to(device=dev, dtype=torch.int64).contiguous()
        n = int(u.numel())
        perm = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        flag = torch.zeros(max(n, 1), dtype=torch.uint8, device=dev)
        check(self.lib.crb_prep_split_loo(self.h, ptr(u), ptr(t), n, int(n_users), ptr(perm), ptr(flag), self.stream))
        return perm[:n], flag[:n].bool()

    def prep_eval_negatives(self, seed, test_users, neg_samples):
        """neg_samples distinct items outside each test user's training items (the installed history): int32 [n_test, neg_samples]."""
        dev = self.device
        tu = torch.as_tensor(test_users).to(device=dev, dtype=torch.int32).contiguous()
        out = torch.empty((int(tu.numel()), int(neg_samples)), dtype=torch.int32, device=dev)
        check(self.lib.crb_prep_eval_negatives(self.h, int(seed), ptr(tu), int(tu.numel()), int(neg_samples), ptr(out), self.stream))
        return out

    def set_history(self, ui_train, n_users, n_items):
        self.set_history_arrays(n_users, n_items, *history_from_dict(ui_train, n_users))
        # per-user interaction lists (order and duplicates kept) inside pos_item: FISM / NAIS need them
        users = np.fromiter(ui_train.keys(), dtype=np.int64, count=len(ui_train))
        lens = np.fromiter((len(v) for v in ui_train.values()), dtype=np.int64, count=len(ui_train))
        start = np.zeros(n_users, dtype=np.int64)
        ln = np.zeros(n_users, dtype=np.int32)
        start[users] = np.concatenate([[0], np.cumsum(lens)[:-1]]) if len(users) else []
        ln[users] = lens
        self._lists = (torch.from_numpy(start).to(self.device), torch.from_numpy(ln).to(self.device))
        check(self.lib.crb_set_history_lists(self.h, ptr(self._lists[0]), ptr(self._lists[1])))

    def epoch_rows(self, neg_ratio, kind="pairwise"):
        return int(self.lib.crb_epoch_rows(self.h, neg_ratio, {"pairwise": 0, "pointwise": 1, "cml": 2, "sbpr": 3}[kind]))

    # ------------------------------------------------------------------ sampler
    def sample_pairwise(self, seed, epoch, first, count, neg_ratio, with_nbr=False):
        u, i, j = (torch.empty(count, dtype=torch.int32, device=self.device) for _ in range(3))
        nbr = torch.empty(count, dtype=torch.int32, device=self.device) if with_nbr else None
        check(self.lib.crb_sample_pairwise(self.h, seed, epoch, first, count, neg_ratio, ptr(u), ptr(i), ptr(j), ptr(nbr), self.stream))
        return (u, i, j, nbr) if with_nbr else (u, i, j)

    def sample_pointwise(self, seed, epoch, first, count, neg_ratio, with_nbr=False):
        u, i = (torch.empty(count, dtype=torch.int32, device=self.device) for _ in range(2))
        y = torch.empty(count, dtype=torch.float32, device=self.device)
        nbr = torch.empty(count, dtype=torch.int32, device=self.device) if with_nbr else None
        check(self.lib.crb_sample_pointwise(self.h, seed, epoch, first, count, neg_ratio, ptr(u), ptr(i), ptr(y), ptr(nbr), self.stream))
        return (u, i, y, nbr) if with_nbr else (u, i, y)

    def sample_cml(self, seed, epoch, first, count, neg_ratio):
        u, i = (torch.empty(count, dtype=torch.int32, device=self.device) for _ in range(2))
        neg = torch.empty((count, neg_ratio), dtype=torch.int32, device=self.device)
        check(self.lib.crb_sample_cml(self.h, seed, epoch, first, count, neg_ratio, ptr(u), ptr(i), ptr(neg), self.stream))
        return u, i, neg

    # ------------------------------------------------------------------ numpy_stream sampler mode (bit-exact reference samplers)
    def np_seed(self, seed):
        """np.random.seed(seed) for the device copy of NumPy's legacy global stream."""
        check(self.lib.crb_np_seed(self.h, int(seed) & 0xFFFFFFFF))

    def np_set_state(self, state=None):
        """Import np.random.get_state() (default: NumPy's current global state)."""
        st = np.random.get_state() if state is None else state
        key = np.ascontiguousarray(st[1], dtype=np.uint32)
        check(self.lib.crb_np_set_state(self.h, ptr(key), int(st[2])))

    def np_get_state(self):
        """The device stream as a tuple accepted by np.random.set_state."""
        key = np.zeros(624, dtype=np.uint32)
        pos = C.c_int32()
        check(self.lib.crb_np_get_state(self.h, ptr(key), C.byref(pos)))
        return ("MT19937", key, int(pos.value), 0, 0.0)

    def sample_epoch_numpy(self, kind, neg_ratio, with_nbr=False):
        """One epoch exactly as the reference sampler returns it under the current stream.  kind: pairwise|pointwise|cml|negatives."""
        dev = self.device
        n_pos = self.n_pos
        if kind == "negatives":
            negs = torch.empty((n_pos, neg_ratio), dtype=torch.int32, device=dev)
            check(self.lib.crb_sample_epoch_numpy(self.h, 3, neg_ratio, None, None, ptr(negs), None, self.stream))
            return negs
        n = {"pairwise": n_pos * neg_ratio, "pointwise": n_pos * (neg_ratio + 1), "cml": n_pos}[kind]
        u, i = torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev)
        if kind == "pairwise":
            j = torch.empty(n, dtype=torch.int32, device=dev)
            nbr = torch.empty(n, dtype=torch.int32, device=dev) if with_nbr else None
            check(self.lib.crb_sample_epoch_numpy(self.h, 0, neg_ratio, ptr(u), ptr(i), ptr(j), ptr(nbr), self.stream))
            return (u, i, j, nbr) if with_nbr else (u, i, j)
        if kind == "pointwise":
            y = torch.empty(n, dtype=torch.float32, device=dev)
            check(self.lib.crb_sample_epoch_numpy(self.h, 1, neg_ratio, ptr(u), ptr(i), ptr(y), None, self.stream))
            return u, i, y
        neg = torch.empty((n, neg_ratio), dtype=torch.int32, device=dev)
        check(self.lib.crb_sample_epoch_numpy(self.h, 2, neg_ratio, ptr(u), ptr(i), ptr(neg), None, self.stream))
        return u, i, neg

    # ------------------------------------------------------------------ training
    @staticmethod
    def _feed_i32(x):
        """Feeds may be host (NumPy / list, like a TF feed_dict) or device tensors; the library stages host buffers."""
        if isinstance(x, torch.Tensor):
            return x if x.dtype == torch.int32 and x.is_contiguous() else x.to(torch.int32).contiguous()   # e.g. a column of an [n, 3] tensor
        return np.ascontiguousarray(np.asarray(x), dtype=np.int32)

    def train_step_bpr(self, P, Q, opt, u, i, j, reg, loss_out=None):
        """One `sess.run([train, loss], {u_idx, i_idx, j_idx})` of BPR.  Returns the loss as a Python float when
        loss_out is None (host read, synchronises), else writes it to the 1-element double device tensor."""
        u, i, j = self._feed_i32(u), self._feed_i32(i), self._feed_i32(j)
        opt.t += 1
        co = opt.c(opt.t)
        host = np.zeros(1, dtype=np.float64) if loss_out is None else None
        check(self.lib.crb_train_step_bpr(self.h, C.byref(P.c), C.byref(Q.c), C.byref(co), ptr(u), ptr(i), ptr(j), len(u),
                                          float(reg), ptr(host) if loss_out is None else ptr(loss_out), self.stream))
        return float(host[0]) if loss_out is None else None

    def train_epoch_bpr(self, P, Q, opt, seed, epoch, first, batch, n_steps, neg_ratio, reg, loss_out):
        """n_steps fused sample+train steps; loss_out: double tensor [n_steps] on the device (no sync) or NumPy (sync)."""
        co = opt.c(opt.t + 1)
        check(self.lib.crb_train_epoch_bpr(self.h, C.byref(P.c), C.byref(Q.c), C.byref(co), seed, epoch, first, batch, n_steps,
                                           neg_ratio, float(reg), ptr(loss_out), self.stream))
        opt.t += n_steps

    def train_epoch_bpr_feeds(self, P, Q, opt, u, i, j, batch, reg, loss_out):
        """RankingRecommender.train_model's loop (:39-46) over caller-sampled epoch arrays u / i / j (host NumPy / pinned torch CPU
        tensors, or device tensors): ceil(len(u) / batch) steps, feeds staged one step ahead.  loss_out: double [n_steps], device
        tensor (no sync) or host (NumPy / pinned tensor; the call returns when the last loss has landed)."""
        def feed(x):
            if isinstance(x, torch.Tensor):
                return x if x.dtype == torch.int32 and x.is_contiguous() else x.to(torch.int32).contiguous()
            return np.ascontiguousarray(np.asarray(x), dtype=np.int32)
        u, i, j = feed(u), feed(i), feed(j)
        n = len(u)
        n_steps = -(-n // batch)
        co = opt.c(opt.t + 1)
        check(self.lib.crb_train_epoch_bpr_feeds(self.h, C.byref(P.c), C.byref(Q.c), C.byref(co), ptr(u), ptr(i), ptr(j), n, batch, float(reg),
                                                 ptr(loss_out), self.stream))
        opt.t += n_steps

    def train_epoch_pointwise(self, kind, P, Q, opt, seed, epoch, first, batch, n_steps, neg_ratio, reg, loss_kind, hvec=None, h_s1=None,
                              h_s2=None, loss_out=None):
        """n_steps fused sample+train steps of MF / GMF; loss_out: double tensor [n_steps] on the device (no sync) or NumPy (sync)."""
        co = opt.c(opt.t + 1)
        check(self.lib.crb_train_epoch_pointwise(self.h, kind, C.byref(P.c), C.byref(Q.c), ptr(hvec), ptr(h_s1), ptr(h_s2), C.byref(co), loss_kind,
                                                 seed, epoch, first, batch, n_steps, neg_ratio, float(reg), ptr(loss_out), self.stream))
        opt.t += n_steps

    def train_step_pointwise(self, kind, P, Q, opt, u, i, y, reg, loss_kind, hvec=None, h_s1=None, h_s2=None, loss_out=None):
        """One `sess.run([train, loss], {u_idx, i_idx, y})` of MF (kind SCORE_DOT) / GMF (kind SCORE_GMF, hvec = h_gmf and
        its optimizer slots, device tensors).  Host or device feeds; returns the loss when loss_out is None."""
        u, i = self._feed_i32(u), self._feed_i32(i)
        if not isinstance(y, torch.Tensor):
            y = np.ascontiguousarray(np.asarray(y), dtype=np.float32)
        elif y.dtype != torch.float32:
            y = y.to(torch.float32)
        opt.t += 1
        co = opt.c(opt.t)
        host = np.zeros(1, dtype=np.float64) if loss_out is None else None
        check(self.lib.crb_train_step_pointwise(self.h, kind, C.byref(P.c), C.byref(Q.c), ptr(hvec), ptr(h_s1), ptr(h_s2), C.byref(co),
                                                loss_kind, ptr(u), ptr(i), ptr(y), len(u), float(reg),
                                                ptr(host) if loss_out is None else ptr(loss_out), self.stream))
        return float(host[0]) if loss_out is None else None

    def train_step_cml(self, P, Q, opt, u, i, neg, margin, reg, item_nums, loss_out=None):
        """One `sess.run([train, loss], {u_idx, i_idx, neg_items})` of CML (dense optimizer apply on both tables)."""
        u, i = self._feed_i32(u), self._feed_i32(i)
        neg = neg.to(torch.int32).contiguous() if isinstance(neg, torch.Tensor) else np.ascontiguousarray(np.asarray(neg), dtype=np.int32)
        for T_ in (P, Q):
            if getattr(T_, "grad", None) is None:
                T_.grad = torch.zeros_like(T_.w)
        opt.t += 1
        co = opt.c(opt.t)
        host = np.zeros(1, dtype=np.float64) if loss_out is None else None
        check(self.lib.crb_train_step_cml(self.h, C.byref(P.c), C.byref(Q.c), ptr(P.grad), ptr(Q.grad), C.byref(co), ptr(u), ptr(i), ptr(neg),
                                          len(u), neg.shape[1], float(margin), float(reg), int(item_nums),
                                          ptr(host) if loss_out is None else ptr(loss_out), self.stream))
        return float(host[0]) if loss_out is None else None

    def train_step_fism(self, P, Q, B, opt, u, i, j, nbr, alpha, reg, reg_bias, conf_batch_size, loss_out=None):
        """One `sess.run([train, loss], {u_idx, i_idx, j_idx, u_neighbors_num})` of FISM (pairwise)."""
        u, i, j, nbr = (self._feed_i32(x) for x in (u, i, j, nbr))
        for T_ in (P, Q, B):
            if getattr(T_, "grad", None) is None:
                T_.grad = torch.zeros_like(T_.w)
        opt.t += 1
        co = opt.c(opt.t)
        host = np.zeros(1, dtype=np.float64) if loss_out is None else None
        check(self.lib.crb_train_step_fism(self.h, C.byref(P.c), C.byref(Q.c), C.byref(B.c), ptr(P.grad), ptr(Q.grad), ptr(B.grad), C.byref(co),
                                           ptr(u), ptr(i), ptr(j), ptr(nbr), len(u), float(alpha), float(reg), float(reg_bias),
                                           int(conf_batch_size), ptr(host) if loss_out is None else ptr(loss_out), self.stream))
        return float(host[0]) if loss_out is None else None

    # ------------------------------------------------------------------ SBPR
    def set_social(self, ui_train, user_friends, SPu, n_users):
        """Device form of what ranking_sampler_sbpr (utils/sampler.py:102-141) walks: see crb_set_social in the header."""
        arrs = social_arrays(ui_train, user_friends, SPu, n_users)
        dts = (torch.int32, torch.int32, torch.int64, torch.int32, torch.float32, torch.int64, torch.int32)
        self._social = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(self.device) for a, dt in zip(arrs, dts))
        t = self._social
        check(self.lib.crb_set_social(self.h, int(t[0].numel()), ptr(t[0]), ptr(t[1]), ptr(t[2]), ptr(t[3]), ptr(t[4]), ptr(t[5]), ptr(t[6])))

    def sample_sbpr(self, seed, epoch, first, count, neg_ratio, is_suk=True):
        u, i, k, j = (torch.empty(count, dtype=torch.int32, device=self.device) for _ in range(4))
        suk = torch.empty(count, dtype=torch.float32, device=self.device) if is_suk else None
        check(self.lib.crb_sample_sbpr(self.h, seed, epoch, first, count, neg_ratio, ptr(u), ptr(i), ptr(k), ptr(j), ptr(suk), self.stream))
        return (u, i, k, j, suk) if is_suk else (u, i, k, j)

    def sample_epoch_numpy_sbpr(self, neg_ratio, is_suk=True):
        """One epoch exactly as ranking_sampler_sbpr returns it under the device copy of NumPy's stream (np_seed / np_set_state)."""
        n = self.epoch_rows(neg_ratio, "sbpr")
        u, i, k, j = (torch.empty(n, dtype=torch.int32, device=self.device) for _ in range(4))
        suk = torch.empty(n, dtype=torch.float32, device=self.device) if is_suk else None
        check(self.lib.crb_sample_epoch_numpy_sbpr(self.h, neg_ratio, ptr(u), ptr(i), ptr(k), ptr(j), ptr(suk), self.stream))
        return (u, i, k, j, suk) if is_suk else (u, i, k, j)

    def train_step_sbpr(self, P, Q, B, opt, u, i, k, j, suk, reg, loss_out=None):
        """One `sess.run([train, loss], {u_idx, i_idx, i_s_idx, i_neg_idx, suk})` of SBPR (SBPR.py:51-57)."""
        u, i, k, j = (self._feed_i32(x) for x in (u, i, k, j))
        if not isinstance(suk, torch.Tensor):
            suk = np.ascontiguousarray(np.asarray(suk), dtype=np.float32)
        elif suk.dtype != torch.float32:
            suk = suk.to(torch.float32)
        for T_ in (P, Q, B):
            if getattr(T_, "grad", None) is None:
                T_.grad = torch.zeros_like(T_.w)
        opt.t += 1
        co = opt.c(opt.t)
        host = np.zeros(1, dtype=np.float64) if loss_out is None else None
        check(self.lib.crb_train_step_sbpr(self.h, C.byref(P.c), C.byref(Q.c), C.byref(B.c), ptr(P.grad), ptr(Q.grad), ptr(B.grad), C.byref(co),
                                           ptr(u), ptr(i), ptr(k), ptr(j), ptr(suk), len(u), float(reg),
                                           ptr(host) if loss_out is None else ptr(loss_out), self.stream))
        return float(host[0]) if loss_out is None else None

    # ------------------------------------------------------------------ TransCF
    def set_item_lists(self):
        """Item-side lists of TransCF's iu_sp_mat (utils/tools.py:100-113) from the current history: the users of every item,
        duplicates kept -- crb_build_history on the swapped columns."""
        pu, pi = self._hist[0], self._hist[1]
        n, dev = int(pu.numel()), self.device
        members = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        grouped_item = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        rp, sc = torch.empty(self.n_items + 1, dtype=torch.int64, device=dev), torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        start, ln = torch.empty(self.n_items + 1, dtype=torch.int64, device=dev), torch.empty(self.n_items, dtype=torch.int32, device=dev)
        n_seen = C.c_int64(0)
        # rows in the order utils/tools.py:103-109 appends them: users in dict order, items in list order = pos_user / pos_item order
        check(self.lib.crb_build_history(self.h, ptr(pi), ptr(pu), n, self.n_items, self.n_users, ptr(grouped_item), ptr(members), ptr(rp), ptr(sc),
                                         C.byref(n_seen), ptr(start), ptr(ln), self.stream))
        self._item_lists = (start, ln, members)
        check(self.lib.crb_set_item_lists(self.h, ptr(start), ptr(ln), ptr(members)))
        return self._item_lists

    def train_step_transcf(self, P, Q, opt, u, i, j, margin, reg1, reg2, loss_out=None):
        """One `sess.run([train, loss], {u_idx, i_idx, j_idx})` of TransCF (TransCF.py:38-71)."""
        u, i, j = (self._feed_i32(x) for x in (u, i, j))
        for T_ in (P, Q):
            if getattr(T_, "grad", None) is None:
                T_.grad = torch.zeros_like(T_.w)
        opt.t += 1
        co = opt.c(opt.t)
        host = np.zeros(1, dtype=np.float64) if loss_out is None else None
        check(self.lib.crb_train_step_transcf(self.h, C.byref(P.c), C.byref(Q.c), ptr(P.grad), ptr(Q.grad), C.byref(co), ptr(u), ptr(i), ptr(j),
                                              len(u), float(margin), float(reg1), float(reg2), ptr(host) if loss_out is None else ptr(loss_out),
                                              self.stream))
        return float(host[0]) if loss_out is None else None

    def transcf_neighbourhood(self, which, table, n_rows):
        """which = 0: alpha of all users (table = Q); which = 1: beta of all items (table = P)."""
        out = torch.empty((n_rows, table.shape[1]), dtype=torch.float32, device=self.device)
        check(self.lib.crb_transcf_neighbourhood(self.h, which, ptr(table), table.shape[1], None, n_rows, ptr(out), self.stream))
        return out

    def score_pairs_transcf(self, P, Q, A, B, u, i):
        u = torch.as_tensor(np.asarray(u), dtype=torch.int32).to(self.device) if not isinstance(u, torch.Tensor) else u.to(torch.int32).contiguous()
        i = torch.as_tensor(np.asarray(i), dtype=torch.int32).to(self.device) if not isinstance(i, torch.Tensor) else i.to(torch.int32).contiguous()
        out = torch.empty(u.numel(), dtype=torch.float32, device=self.device)
        check(self.lib.crb_score_pairs_transcf(self.h, ptr(P), ptr(Q), ptr(A), ptr(B), P.shape[1], ptr(u), ptr(i), u.numel(), ptr(out), self.stream))
        return out

    def fism_user_vectors(self, P, users, nbr, alpha):
        users = torch.as_tensor(np.asarray(users), dtype=torch.int32).to(self.device) if not isinstance(users, torch.Tensor) else users.to(torch.int32)
        nbr = torch.as_tensor(np.asarray(nbr), dtype=torch.int32).to(self.device) if not isinstance(nbr, torch.Tensor) else nbr.to(torch.int32)
        out = torch.empty((users.numel(), P.shape[1]), dtype=torch.float32, device=self.device)
        check(self.lib.crb_fism_user_vectors(self.h, ptr(P), P.shape[1], ptr(users), ptr(nbr), users.numel(), float(alpha), ptr(out), self.stream))
        return out

    def clip_rows(self, src, max_norm=1.0):
        dst = torch.empty_like(src)
        check(self.lib.crb_clip_rows(self.h, ptr(src), ptr(dst), src.shape[0], src.shape[1], float(max_norm), self.stream))
        return dst

    def train_step_neumf(self, tabs, dense, dense_s1, dense_s2, n_layers, opt, u, i, y, reg1, reg2, loss_kind, loss_out=None):
        """One `sess.run([train, loss], {u_idx, i_idx, y})` of NeuMF.  tabs = (P_gmf, Q_gmf, P_mlp, Q_mlp) Tables."""
        u, i = self._feed_i32(u), self._feed_i32(i)
        if not isinstance(y, torch.Tensor):
            y = np.ascontiguousarray(np.asarray(y), dtype=np.float32)
        for T_ in tabs:   # tabs[0] = tabs[1] = None: the MLP model (no GMF branch)
            if T_ is not None and getattr(T_, "grad", None) is None:
                T_.grad = torch.zeros_like(T_.w)
        opt.t += 1
        co = opt.c(opt.t)
        host = np.zeros(1, dtype=np.float64) if loss_out is None else None
        check(self.lib.crb_train_step_neumf(self.h, *(C.byref(T_.c) if T_ is not None else None for T_ in tabs),
                                            *(ptr(T_.grad) if T_ is not None else None for T_ in tabs), ptr(dense), ptr(dense_s1),
                                            ptr(dense_s2), n_layers, C.byref(co), loss_kind, ptr(u), ptr(i), ptr(y), len(u), float(reg1),
                                            float(reg2), ptr(host) if loss_out is None else ptr(loss_out), self.stream))
        return float(host[0]) if loss_out is None else None

    def score_pairs_neumf(self, tabs, dense, n_layers, u, i):
        u = torch.as_tensor(np.asarray(u), dtype=torch.int32).to(self.device) if not isinstance(u, torch.Tensor) else u.to(torch.int32).contiguous()
        i = torch.as_tensor(np.asarray(i), dtype=torch.int32).to(self.device) if not isinstance(i, torch.Tensor) else i.to(torch.int32).contiguous()
        out = torch.empty(u.numel(), dtype=torch.float32, device=self.device)
        gmf = tabs[0] is not None
        check(self.lib.crb_score_pairs_neumf(self.h, ptr(tabs[0].w) if gmf else None, ptr(tabs[1].w) if gmf else None, ptr(tabs[2].w), ptr(tabs[3].w),
                                             ptr(dense), tabs[0].dim if gmf else 0, tabs[2].dim, n_layers, ptr(u), ptr(i), u.numel(), ptr(out), self.stream))
        return out

    def train_step_lrml(self, P, Q, dense, dense_s1, dense_s2, mem_size, opt, u, i, j, margin, reg, loss_out=None):
        """One `sess.run([train, loss], {u_idx, i_idx, j_idx})` of LRML (LRML.py:53-64).  dense = K [d, mem] then M [mem, d]."""
        u, i, j = (self._feed_i32(x) for x in (u, i, j))
        for T_ in (P, Q):
            if getattr(T_, "grad", None) is None:
                T_.grad = torch.zeros_like(T_.w)
        opt.t += 1
        co = opt.c(opt.t)
        host = np.zeros(1, dtype=np.float64) if loss_out is None else None
        check(self.lib.crb_train_step_lrml(self.h, C.byref(P.c), C.byref(Q.c), ptr(P.grad), ptr(Q.grad), ptr(dense), ptr(dense_s1), ptr(dense_s2),
                                           int(mem_size), C.byref(co), ptr(u), ptr(i), ptr(j), len(u), float(margin), float(reg),
                                           ptr(host) if loss_out is None else ptr(loss_out), self.stream))
        return float(host[0]) if loss_out is None else None

    def score_pairs_lrml(self, P, Q, dense, mem_size, u, i):
        u = torch.as_tensor(np.asarray(u), dtype=torch.int32).to(self.device) if not isinstance(u, torch.Tensor) else u.to(torch.int32).contiguous()
        i = torch.as_tensor(np.asarray(i), dtype=torch.int32).to(self.device) if not isinstance(i, torch.Tensor) else i.to(torch.int32).contiguous()
        out = torch.empty(u.numel(), dtype=torch.float32, device=self.device)
        check(self.lib.crb_score_pairs_lrml(self.h, ptr(P), ptr(Q), ptr(dense), P.shape[1], int(mem_size), ptr(u), ptr(i), u.numel(), ptr(out),
                                            self.stream))
        return out

    def mask_seen(self, scores, users, value=float("-inf")):
        users = users.to(torch.int32).contiguous()
        check(self.lib.crb_mask_seen(self.h, ptr(scores), ptr(users), scores.shape[0], scores.shape[1], float(value), self.stream))
        return scores

    def _ensure_grads(self, tabs):
        for T_ in tabs:
            if getattr(T_, "grad", None) is None:
                T_.grad = torch.zeros_like(T_.w)

    def train_step_nais(self, P, Q, B, dense, dense_s1, dense_s2, atten_size, opt, hist, targets, y, beta, reg, loss_out=None, concat=False):
        """One `sess.run([train, loss], {u_idx: history, i_idx: targets, y})` of NAIS_single (one user).  concat: atten_type 'concat'
        (W has 2*dim rows) instead of 'prod'."""
        dev = self.device
        hist = torch.as_tensor(np.asarray(hist), dtype=torch.int32).to(dev) if not isinstance(hist, torch.Tensor) else hist.to(torch.int32)
        targets = torch.as_tensor(np.asarray(targets), dtype=torch.int32).to(dev) if not isinstance(targets, torch.Tensor) else targets.to(torch.int32)
        y = torch.as_tensor(np.asarray(y), dtype=torch.float32).to(dev) if not isinstance(y, torch.Tensor) else y.to(torch.float32)
        self._ensure_grads((P, Q, B))
        opt.t += 1
        co = opt.c(opt.t)
        lo = torch.zeros(1, dtype=torch.float64, device=dev) if loss_out is None else loss_out
        check(self.lib.crb_train_step_nais(self.h, C.byref(P.c), C.byref(Q.c), C.byref(B.c), ptr(P.grad), ptr(Q.grad), ptr(B.grad), ptr(dense),
                                           ptr(dense_s1), ptr(dense_s2), atten_size, 1 if concat else 0, C.byref(co), ptr(hist), hist.numel(), ptr(targets),
                                           ptr(y), targets.numel(), float(beta), float(reg), ptr(lo), self.stream))
        return float(lo.item()) if loss_out is None else None

    def train_epoch_nais(self, P, Q, B, dense, dense_s1, dense_s2, atten_size, opt, seed, epoch, list_start, list_len, neg_ratio, beta, reg, losses,
                         concat=False):
        self._ensure_grads((P, Q, B))
        list_start = np.ascontiguousarray(list_start, dtype=np.int64)
        list_len = np.ascontiguousarray(list_len, dtype=np.int32)
        co = opt.c(opt.t + 1)
        check(self.lib.crb_train_epoch_nais(self.h, C.byref(P.c), C.byref(Q.c), C.byref(B.c), ptr(P.grad), ptr(Q.grad), ptr(B.grad), ptr(dense),
                                            ptr(dense_s1), ptr(dense_s2), atten_size, 1 if concat else 0, C.byref(co), seed, epoch, ptr(list_start),
                                            ptr(list_len), len(list_len), neg_ratio, float(beta), float(reg), ptr(losses), self.stream))
        opt.t += len(list_len)

    def sample_nais(self, seed, epoch, pos_first, n_pos_user, neg_ratio):
        m = n_pos_user * (neg_ratio + 1)
        tg = torch.empty(m, dtype=torch.int32, device=self.device)
        y = torch.empty(m, dtype=torch.float32, device=self.device)
        check(self.lib.crb_sample_nais(self.h, seed, epoch, pos_first, n_pos_user, neg_ratio, ptr(tg), ptr(y), self.stream))
        return tg, y

    def score_nais(self, P, Q, bias, dense, atten_size, hist, targets, beta, out=None, concat=False):
        """out: optional contiguous float32 device tensor [len(targets)] to write into (e.g. a slice of a larger buffer)."""
        dev = self.device
        hist = torch.as_tensor(np.asarray(hist), dtype=torch.int32).to(dev) if not isinstance(hist, torch.Tensor) else hist.to(torch.int32)
        targets = torch.as_tensor(np.asarray(targets), dtype=torch.int32).to(dev) if not isinstance(targets, torch.Tensor) else targets.to(torch.int32)
        if out is None:
            out = torch.empty(targets.numel(), dtype=torch.float32, device=dev)
        check(self.lib.crb_score_nais(self.h, ptr(P), ptr(Q), ptr(bias), ptr(dense), P.shape[1], atten_size, 1 if concat else 0, ptr(hist), hist.numel(),
                                      ptr(targets), targets.numel(), float(beta), ptr(out), self.stream))
        return out

    def sampler_errors(self):
        """Rows for which the device sampler ran out of attempts since the last call (synchronises; clears the counter).  The
        reference loops forever in that case (utils/sampler.py:58-61); here it is an error the epoch loop raises."""
        n = C.c_uint32()
        check(self.lib.crb_sampler_errors(self.h, C.byref(n), self.stream))
        return int(n.value)

    def adam_flush(self, table, opt):
        if opt.kind == "Adam" and opt.adam_mode == "tf1" and opt.t > 0:
            co = opt.c(opt.t)
            check(self.lib.crb_adam_flush(self.h, C.byref(table.c), C.byref(co), self.stream))

    # ------------------------------------------------------------------ evaluation
    def score_pairs(self, kind, P, Q, u, i, hvec=None, out=None):
        """Scores of flattened (user, item) pairs in canonical fp32.  u/i host or device; returns NumPy when the
        feeds are host arrays (like sess.run), else a device tensor."""
        u, i = self._feed_i32(u), self._feed_i32(i)
        n = len(u)
        host = not isinstance(u, torch.Tensor)
        if out is None:
            out = np.empty(n, dtype=np.float32) if host else torch.empty(n, dtype=torch.float32, device=self.device)
        check(self.lib.crb_score_pairs(self.h, kind, ptr(P), ptr(Q), ptr(hvec), P.shape[1], ptr(u), ptr(i), n, ptr(out), self.stream))
        return out

    def topk_segments(self, scores, offsets, K, ascending=False):
        n_users = len(offsets) - 1
        host = not isinstance(scores, torch.Tensor)
        if host:
            scores = np.ascontiguousarray(scores, dtype=np.float32)
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            out = np.empty((n_users, K), dtype=np.int32)
        else:
            offsets = torch.as_tensor(offsets, dtype=torch.int64, device=self.device)
            out = torch.empty((n_users, K), dtype=torch.int32, device=self.device)
        check(self.lib.crb_topk_segments(self.h, ptr(scores), ptr(offsets), n_users, K, 1 if ascending else 0, ptr(out), self.stream))
        return out

    def score_pairs_topk(self, kind, P, Q, seg_users, items, offsets, K, hvec=None, ascending=False):
        """test_model_loo's predict + argsort in one kernel: for segment k (user row seg_users[k], candidates
        items[offsets[k]:offsets[k+1]]) the K best positions inside the segment, -1 padded.  Host feeds -> NumPy, device -> tensor.
        Shapes the fused kernel does not take (K > 32, dim % 4 != 0) run crb_score_pairs + crb_topk_segments: same results."""
        n_users = len(offsets) - 1
        host = not isinstance(items, torch.Tensor)
        if K > 32 or P.shape[1] % 4 != 0 or P.shape[1] > 512:
            lens = np.diff(np.asarray(offsets.cpu() if isinstance(offsets, torch.Tensor) else offsets, dtype=np.int64))
            if host:
                u = np.repeat(np.asarray(seg_users, dtype=np.int32), lens)
            else:
                u = torch.repeat_interleave(torch.as_tensor(seg_users, dtype=torch.int32, device=self.device), torch.from_numpy(lens).to(self.device))
            return self.topk_segments(self.score_pairs(kind, P, Q, u, items, hvec=hvec), offsets, K, ascending=ascending)
        if host:
            seg_users, items = np.ascontiguousarray(seg_users, dtype=np.int32), np.ascontiguousarray(items, dtype=np.int32)
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            out = np.empty((n_users, K), dtype=np.int32)
        else:
            seg_users = torch.as_tensor(seg_users, dtype=torch.int32, device=self.device).contiguous()
            items = items.to(torch.int32).contiguous()
            offsets = torch.as_tensor(offsets, dtype=torch.int64, device=self.device).contiguous()
            out = torch.empty((n_users, K), dtype=torch.int32, device=self.device)
        check(self.lib.crb_score_pairs_topk(self.h, kind, ptr(P), ptr(Q), ptr(hvec), P.shape[1], ptr(seg_users), ptr(items), ptr(offsets), n_users, K,
                                            1 if ascending else 0, ptr(out), self.stream))
        return out

    def score_topk(self, kind, P, Q, users, K, hvec=None, hist_users=None, exact=False, n_items=None, return_scores=False):
        """Best K unseen item ids of each user (test_model_rs).  users host -> NumPy out; device -> tensors."""
        users = self._feed_i32(users)
        n = len(users)
        host = not isinstance(users, torch.Tensor)
        if hist_users is not None:
            hist_users = self._feed_i32(hist_users)
        n_items = Q.shape[0] if n_items is None else n_items
        self._eval_cache_guard(Q, hvec)
        if host:
            items = np.empty((n, K), dtype=np.int32)
            scores = np.empty((n, K), dtype=np.float32) if return_scores else None
        else:
            items = torch.empty((n, K), dtype=torch.int32, device=self.device)
            scores = torch.empty((n, K), dtype=torch.float32, device=self.device) if return_scores else None
        check(self.lib.crb_score_topk(self.h, kind, ptr(P), ptr(Q), ptr(hvec), n_items, P.shape[1], ptr(users), ptr(hist_users), n, K,
                                      1 if exact else 0, ptr(items), ptr(scores), self.stream))
        return (items, scores) if return_scores else items

    def _eval_cache_guard(self, Q, hvec):
        """The library keys its cached bf16 item table by ADDRESS (Q, hvec, sizes, kind) and drops it whenever one of its own calls writes
        a table.  What it cannot see is torch: an in-place op on Q (or on any view of it) and a table that was freed and whose address the
        caching allocator handed to the next one.  Both are visible here -- the version counter of the tensor (shared by its views) and
        the identity of the object that owns the memory -- so the copy is dropped for them too.  Anything unexpected drops it as well:
        a spurious conversion costs 2 ms at 2M items, a stale one returns another table's ranking."""
        try:
            sig = []
            for t in (Q, hvec):
                if t is None:
                    sig.append(None)
                    continue
                base = t._base if t._base is not None else t
                sig.append((weakref.ref(base), base._version))
            prev = getattr(self, "_evq_sig", None)
            same = prev is not None and all(
                (a is None and b is None) or (a is not None and b is not None and a[0]() is b[0]() and a[0]() is not None and a[1] == b[1])
                for a, b in zip(prev, sig))
            self._evq_sig = sig
        except Exception:
            same, self._evq_sig = False, None
        if not same:
            self.invalidate_eval_cache()

    def invalidate_eval_cache(self):
        """score_topk keeps the bf16 copy of the item table between calls; library calls that write tables drop it themselves.
        Call this after writing Q / hvec with torch (e.g. restoring a checkpoint into the same buffer)."""
        check(self.lib.crb_eval_cache_invalidate(self.h))

    def score_topk_stats(self):
        st = (C.c_int64 * 4)()
        check(self.lib.crb_score_topk_stats(self.h, C.byref(st)))
        return {"certified": st[0], "exact_rerun": st[1], "max_candidates": st[2]}


def history_from_dict(ui_train, n_users):
    """data.ui_train (dict[int -> list[int]], RankingPreprocess.py:117) -> flat positives in the enumeration order of
    utils/sampler.py:50-52 + per-user sorted-unique CSR (the `seen_items` sets)."""
    users = list(ui_train.keys())
    lens = np.fromiter((len(ui_train[u]) for u in users), dtype=np.int64, count=len(users))
    pos_user = np.repeat(np.asarray(users, dtype=np.int32), lens)
    pos_item = np.fromiter((i for u in users for i in ui_train[u]), dtype=np.int32, count=int(lens.sum()))
    key = np.unique(pos_user.astype(np.int64) * (1 << 31) + pos_item.astype(np.int64))
    su, sc = (key >> 31).astype(np.int64), (key & ((1 << 31) - 1)).astype(np.int32)
    seen_rowptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(su, minlength=n_users), out=seen_rowptr[1:])
    return pos_user, pos_item, seen_rowptr, sc


def social_arrays(ui_train, user_friends, SPu, n_users):
    """-> (sp_pos_user, sp_pos_item, spu_start, spu_items, spu_suk, excl_rowptr, excl_cols): the flat form of the loops of
    ranking_sampler_sbpr (utils/sampler.py:105-131).  suk of an SPu entry counts u's friends -- as often as they are listed --
    that have a training list containing the item (:126-130)."""
    pos_u, pos_i = [], []
    spu_start = np.zeros(n_users + 1, dtype=np.int64)
    excl_len = np.zeros(n_users + 1, dtype=np.int64)
    sp_items, sp_suk, excl = [], [], []
    lens = np.zeros(n_users, dtype=np.int64)
    per_user = {}
    for u, items in ui_train.items():
        if u not in SPu:
            continue
        pos_u.extend([u] * len(items))
        pos_i.extend(items)
        spu = np.asarray(SPu[u], dtype=np.int64)
        counts = {}
        for friend in user_friends[u]:
            if friend not in ui_train:
                continue
            for it in set(ui_train[friend]):
                counts[it] = counts.get(it, 0) + 1
        per_user[u] = (spu, np.asarray([counts.get(int(it), 0) for it in spu], dtype=np.float32),
                       np.union1d(np.asarray(items, dtype=np.int64), spu))
        lens[u] = spu.shape[0]
    for u in range(n_users):
        if u in per_user:
            spu, suk, ex = per_user[u]
            sp_items.append(spu); sp_suk.append(suk); excl.append(ex)
            excl_len[u + 1] = ex.shape[0]
    np.cumsum(lens, out=spu_start[1:])
    excl_rowptr = np.cumsum(excl_len)
    cat = lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dtype=dt)
    return (np.asarray(pos_u, dtype=np.int32), np.asarray(pos_i, dtype=np.int32), spu_start, cat(sp_items, np.int32), cat(sp_suk, np.float32),
            excl_rowptr.astype(np.int64), cat(excl, np.int32))
