"""Multi-GPU (one process per GPU, one box) BPR training over NVLink peer memory -- the host side of csrc/train_sharded.cu.

Partitioning: users are split into contiguous ranges, one per rank (rows of P, their histories and their sampling
are local); the item table Q is row-sharded by `item % world` and every rank maps every shard, every gradient inbox and every
barrier flag array through CUDA IPC.  torch.distributed is used for exactly one thing on the data path: exchanging the 64-byte
IPC handles once.  The per-step barriers are flag barriers in peer memory (crb_shard_barrier, on the compute stream, no host
synchronisation); there is no data-path collective.  `barrier="host"` (stream synchronise + a torch.distributed barrier) serves
process groups whose ranks share ONE device (the 1-GPU parity tests: gloo control plane, CUDA IPC between two processes on the
same GPU -- a spinning flag barrier would only burn the other process's time slice there)."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import CrbShard, CrbTable, check, ptr
from .engine import Engine, Optimizer, Table


# ---------------------------------------------------------------------------------------------- pure partition logic (CPU-testable)
def user_range(n_users, rank, world):
    """Contiguous user range [lo, hi) owned by `rank`."""
    return (n_users * rank) // world, (n_users * (rank + 1)) // world


def item_owner(item, world):
    return item % world


def item_local_row(item, world):
    return item // world


def shard_rows(n_items, rank, world):
    """Number of item rows stored on `rank` under owner = item % world."""
    return (n_items - rank + world - 1) // world


def shard_history(ui_train, n_users, rank, world):
    """The part of data.ui_train owned by `rank`, re-keyed to local user rows (item ids stay global)."""
    lo, hi = user_range(n_users, rank, world)
    return {u - lo: items for u, items in ui_train.items() if lo <= u < hi}, hi - lo


class _DeviceBuffer(object):
    """Library-allocated (cudaMalloc) device memory viewed as a torch tensor -- exportable through CUDA IPC."""

    def __init__(self, engine, shape, dtype):
        self.engine, self.shape, self.dtype = engine, tuple(shape), dtype
        self.nbytes = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        p = C.c_void_p()
        check(engine.lib.crb_malloc(engine.h, max(self.nbytes, 16), C.byref(p)))
        self.ptr = p.value
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint32: "<u4"}[dtype]
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": typestr, "data": (self.ptr, False), "version": 2, "strides": None}
        self.tensor = torch.as_tensor(self, device=engine.device)

    def export(self):
        buf = C.create_string_buffer(64)
        check(self.engine.lib.crb_ipc_export(self.engine.h, self.ptr, buf))
        return bytes(buf.raw)


def _is_gloo(group):
    return dist.get_backend(group) == "gloo"


def broadcast_dev(t, src, group=None):
    """dist.broadcast of a CUDA tensor that also works on a gloo group (staged through the host)."""
    if _is_gloo(group):
        c = t.cpu()
        dist.broadcast(c, src=src, group=group)
        t.copy_(c)
    else:
        dist.broadcast(t, src=src, group=group)


def all_reduce_dev(t, op=None, group=None):
    op = dist.ReduceOp.SUM if op is None else op
    if _is_gloo(group):
        c = t.cpu()
        dist.all_reduce(c, op=op, group=group)
        t.copy_(c)
    else:
        dist.all_reduce(t, op=op, group=group)


def gather_dev(t, dst, group=None):
    """dist.gather of equally shaped CUDA tensors -> list at `dst` (None elsewhere); host-staged on a gloo group."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if _is_gloo(group):
        c = t.cpu()
        parts = [torch.empty_like(c) for _ in range(world)] if rank == dst else None
        dist.gather(c, parts, dst=dst, group=group)
        return [p.to(t.device) for p in parts] if rank == dst else None
    parts = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, parts, dst=dst, group=group)
    return parts


class ShardedBPR(object):
    """BPR (model/ranking/BPR.py:31-44) with P partitioned by user and Q row-sharded over the ranks of one box."""

    def __init__(self, engine, n_users, n_items, dim, optimizer, lr, adam_mode, batch, init_P=None, init_Q=None, seed=0, group=None,
                 barrier=None, barrier_timeout_ms=20000):
        self.engine, self.group = engine, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world <= _lib.MAX_RANKS
        self.n_users, self.n_items, self.dim, self.batch = n_users, n_items, dim, batch
        self.u_lo, self.u_hi = user_range(n_users, self.rank, self.world)
        self.opt = Optimizer(optimizer, lr, adam_mode=adam_mode)
        self.barrier_mode = barrier or ("host" if _is_gloo(group) else "flag")
        assert self.barrier_mode in ("flag", "host")
        self.barrier_timeout_ms, self._ticket = int(barrier_timeout_ms), 0
        dev = engine.device
        g = torch.Generator(device=dev).manual_seed(seed)
        # user rows: ordinary torch memory
        if init_P is None:
            init_P = torch.randn(self.u_hi - self.u_lo, dim, device=dev, generator=g) * 0.01
        self.P = Table(torch.as_tensor(init_P, dtype=torch.float32).to(dev).contiguous(), optimizer, adam_mode)
        # item shard + inbox + flags: IPC-exportable memory
        rows = shard_rows(n_items, self.rank, self.world)
        self.q_rows = rows
        kinds = ["w"] + (["s1"] if optimizer != "SGD" else []) + (["s2"] if optimizer == "Adam" else [])
        self.q = {k: _DeviceBuffer(engine, (rows, dim), torch.float32) for k in kinds}
        if optimizer == "Adam" and adam_mode == "tf1":
            self.q["last"] = _DeviceBuffer(engine, (rows,), torch.int32)
        if init_Q is None:
            gq = torch.Generator(device=dev).manual_seed(seed + 1 + self.rank)
            self.q["w"].tensor.copy_(torch.randn(rows, dim, device=dev, generator=gq) * 0.01)
        else:  # init_Q is the FULL table (parity runs): keep this rank's rows
            self.q["w"].tensor.copy_(torch.as_tensor(init_Q, dtype=torch.float32)[self.rank::self.world].to(dev))
        if optimizer == "Adagrad":
            self.q["s1"].tensor.fill_(0.1)
        # direct-mapped inbox: one gradient slot per (source rank, local row) -- a source sums its own duplicates before it sends,
        # so the inbox cannot overflow whatever the skew; world * rows_cap * dim floats = the size of the whole item table
        self.rows_cap = shard_rows(n_items, 0, self.world)
        self.inbox = {"grad": _DeviceBuffer(engine, (self.world * self.rows_cap, dim), torch.float32),
                      "stamp": _DeviceBuffer(engine, (self.world * self.rows_cap,), torch.int32),
                      "flags": _DeviceBuffer(engine, (_lib.SHARD_FLAGS,), torch.int32),
                      "dense": _DeviceBuffer(engine, (self.world * _lib.SHARD_DENSE,), torch.float32)}
        self._map_peers()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def _map_peers(self):
        mine = {"rows": self.q_rows}
        mine.update({"q_" + k: b.export() for k, b in self.q.items()})
        mine.update({"in_" + k: b.export() for k, b in self.inbox.items()})
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        self._opened = []
        sh = CrbShard()
        sh.n_ranks, sh.rank, sh.rows_cap = self.world, self.rank, self.rows_cap

        def open_(handle):
            p = C.c_void_p()
            check(self.engine.lib.crb_ipc_open(self.engine.h, handle, C.byref(p)))
            self._opened.append(p.value)
            return p.value
        for r in range(self.world):
            info = everyone[r]
            local = r == self.rank
            def pointer(kind, store):
                if kind not in store:
                    return None
                return store[kind].ptr if local else open_(info[("q_" if store is self.q else "in_") + kind])
            sh.q[r] = CrbTable(pointer("w", self.q), pointer("s1", self.q), pointer("s2", self.q), pointer("last", self.q), info["rows"], self.dim, 0)
            sh.inbox_grad[r], sh.inbox_stamp[r], sh.flags[r] = pointer("grad", self.inbox), pointer("stamp", self.inbox), pointer("flags", self.inbox)
            sh.dense_inbox[r] = pointer("dense", self.inbox)
        self.shard = sh

    def barrier(self):
        """Cross-rank barrier ordered with the compute stream."""
        if self.world == 1:
            return
        if self.barrier_mode == "host":
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            return
        self._ticket += 1
        check(self.engine.lib.crb_shard_barrier(self.engine.h, C.byref(self.shard), self._ticket, self.barrier_timeout_ms, self.engine.stream))

    def set_history(self, ui_train_local, n_users_local):
        """ui_train_local: this rank's users keyed by LOCAL row (see shard_history), item ids global."""
        self.engine.set_history(ui_train_local, n_users_local, self.n_items)

    def step(self, reg, neg_ratio=None, seed=0, epoch=0, first=0, batch=None, feed=None, loss_out=None):
        """One synchronous step over the union batch.  feed = (u_local, i_global, j_global) or None to sample on the device."""
        eng, lib = self.engine, self.engine.lib
        self.opt.t += 1
        co = self.opt.c(self.opt.t)
        batch = self.batch if batch is None else batch
        u = i = j = None
        if feed is not None:
            u, i, j = (eng._feed_i32(x) for x in feed)
            batch = len(u)
        host = np.zeros(1, dtype=np.float64) if loss_out is None else None
        check(lib.crb_shard_step_compute(eng.h, C.byref(self.P.c), C.byref(self.shard), C.byref(co), ptr(u), ptr(i), ptr(j), seed, epoch, first,
                                         neg_ratio or 1, batch, float(reg), ptr(host) if loss_out is None else ptr(loss_out), eng.stream))
        self.barrier()
        check(lib.crb_shard_apply_inbox(eng.h, C.byref(self.shard), C.byref(co), eng.stream))
        self.barrier()
        return float(host[0]) if loss_out is None else None

    def run_steps(self, n_steps, reg, neg_ratio, seed, epoch, first=0, batch=None, loss_out=None, bounds=None, feeds=None, host_losses=None,
                  phase_ms=None):
        """n synchronous steps over consecutive batches of this rank's epoch, sampled on the device.  Step k+1's index work
        (sampler, user-row counting and slot assignment) is prepared on the engine's auxiliary stream while step k's barriers and
        inbox phase run (crb_shard_step_prepare).  loss_out: device float64 [n_steps] or None.  bounds: optional n_steps+1 row
        offsets (step k covers rows [bounds[k], bounds[k+1]) of this rank's epoch) instead of a fixed batch.
        feeds: optional list of n_steps (u_local, i_global, j_global) int32 arrays (host -- ideally pinned -- or device): the caller's
        own triplets are staged one step ahead instead of being sampled (the reference's epoch loop over its sampler's output).
        host_losses: optional pinned float64 CPU tensor [n_steps]: every step's loss is copied to it asynchronously as the step ends.
        phase_ms: optional dict; CUDA events on the compute stream then split every step into compute (fetch + fused step + duplicate
        reduce / send + loss) / barrier1 / apply (the owner's inbox pass) / barrier2 and the totals are added to it (profiling)."""
        eng, lib = self.engine, self.engine.lib
        batch = self.batch if batch is None else batch
        if bounds is None:
            bounds = [first + k * batch for k in range(n_steps + 1)]
        assert len(bounds) == n_steps + 1 and all(b > a for a, b in zip(bounds[:-1], bounds[1:])), "every step needs at least one row"

        if feeds is not None:
            feeds = [tuple(eng._feed_i32(x) for x in f) for f in feeds]
            assert len(feeds) == n_steps
            bounds = [0]
            for f in feeds:
                bounds.append(bounds[-1] + len(f[0]))
            neg_ratio = neg_ratio or 1
        if host_losses is not None and loss_out is None:
            loss_out = torch.zeros(n_steps, dtype=torch.float64, device=eng.device)

        def prepare(k):
            f = feeds[k] if feeds is not None else (None, None, None)
            check(lib.crb_shard_step_prepare(eng.h, C.byref(self.P.c), seed, epoch, bounds[k], neg_ratio, bounds[k + 1] - bounds[k], self.n_items,
                                             ptr(f[0]), ptr(f[1]), ptr(f[2]), eng.stream))
        prepare(0)
        host = np.zeros(1, dtype=np.float64)
        marks = [] if phase_ms is not None else None

        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        for k in range(n_steps):
            self.opt.t += 1
            co = self.opt.c(self.opt.t)
            lo = ptr(host) if loss_out is None else ptr(loss_out[k:k + 1])
            mark()
            check(lib.crb_shard_step_compute(eng.h, C.byref(self.P.c), C.byref(self.shard), C.byref(co), None, None, None, seed, epoch,
                                             bounds[k], neg_ratio, bounds[k + 1] - bounds[k], float(reg), lo, eng.stream))
            if host_losses is not None:
                host_losses[k:k + 1].copy_(loss_out[k:k + 1], non_blocking=True)
            if k + 1 < n_steps:
                prepare(k + 1)
            mark()
            self.barrier()
            mark()
            check(lib.crb_shard_apply_inbox(eng.h, C.byref(self.shard), C.byref(co), eng.stream))
            mark()
            self.barrier()
        if marks is not None:
            mark()
            torch.cuda.synchronize()
            names = ("compute", "barrier1", "apply", "barrier2")
            for k in range(n_steps):
                for q, name in enumerate(names):
                    a, b = marks[4 * k + q], marks[4 * k + q + 1]
                    phase_ms[name] = phase_ms.get(name, 0.0) + a.elapsed_time(b)
            phase_ms["steps"] = phase_ms.get("steps", 0) + n_steps

    def check(self):
        """Synchronises and raises CrbError if a cross-rank barrier timed out or the sampler gave up on a row since the last check
        (call once per epoch: the step calls themselves never synchronise)."""
        check(self.engine.lib.crb_shard_check(self.engine.h, C.byref(self.shard), self.engine.stream))

    def flush(self):
        """Bring CRB_ADAM_TF1 tables up to date before reading them."""
        self.engine.adam_flush(self.P, self.opt)
        if self.opt.kind == "Adam" and self.opt.adam_mode == "tf1" and self.opt.t > 0:
            co = self.opt.c(self.opt.t)
            check(self.engine.lib.crb_adam_flush(self.engine.h, C.byref(self.shard.q[self.rank]), C.byref(co), self.engine.stream))

    def gather_P(self):
        """All user rows on every rank, in global user order (checkpoints / tests)."""
        self.flush()
        parts = [torch.zeros(user_range(self.n_users, r, self.world)[1] - user_range(self.n_users, r, self.world)[0], self.dim,
                             device=self.engine.device) for r in range(self.world)]
        for r in range(self.world):
            if r == self.rank:
                parts[r].copy_(self.P.w)
            if self.world > 1:
                broadcast_dev(parts[r], r, self.group)
        return torch.cat(parts)

    def gather_Q(self):
        """Full item table on every rank (tests / evaluation set-up)."""
        self.flush()
        parts = [torch.zeros(shard_rows(self.n_items, r, self.world), self.dim, device=self.engine.device) for r in range(self.world)]
        if self.world > 1:
            if len({p.shape for p in parts}) == 1 and not _is_gloo(self.group):
                dist.all_gather(parts, self.q["w"].tensor.contiguous(), group=self.group)
            else:
                self._gather_ragged(parts)
        else:
            parts[0] = self.q["w"].tensor
        full = torch.zeros(self.n_items, self.dim, device=self.engine.device)
        for r in range(self.world):
            full[r::self.world] = parts[r]
        return full

    def _gather_ragged(self, parts):
        for r in range(self.world):
            if r == self.rank:
                parts[r].copy_(self.q["w"].tensor)
            broadcast_dev(parts[r], r, self.group)

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for p in self._opened:
            self.engine.lib.crb_ipc_close(self.engine.h, p)
        for b in list(self.q.values()) + list(self.inbox.values()):
            b.tensor = None
            self.engine.lib.crb_free(self.engine.h, b.ptr)


class ShardedPointwise(ShardedBPR):
    """MF / GMF (model/ranking/GMF.py:37-49; MF per SURVEY F6) over the same partition: P by user, Q row-sharded, item rows and
    gradients over peer memory exactly as for BPR.  GMF's h is replicated; its gradient is summed over the ranks through the dense
    inbox (crb_shard_apply_dense), the same update on every rank."""

    def __init__(self, engine, n_users, n_items, dim, optimizer, lr, adam_mode, batch, kind=_lib.SCORE_DOT, loss_kind=_lib.LOSS_CROSS_ENTROPY,
                 init_P=None, init_Q=None, init_h=None, seed=0, group=None, barrier=None, barrier_timeout_ms=20000):
        ShardedBPR.__init__(self, engine, n_users, n_items, dim, optimizer, lr, adam_mode, batch, init_P=init_P, init_Q=init_Q, seed=seed, group=group,
                            barrier=barrier, barrier_timeout_ms=barrier_timeout_ms)
        self.kind, self.loss_kind = kind, loss_kind
        self.h = self.h_s1 = self.h_s2 = None
        if kind == _lib.SCORE_GMF:
            dev = engine.device
            if init_h is None:   # the same vector on every rank
                init_h = torch.randn(dim, generator=torch.Generator().manual_seed(seed + 977)) * 0.01
            self.h = torch.as_tensor(init_h, dtype=torch.float32).to(dev).contiguous()
            if optimizer == "Adagrad":
                self.h_s1 = torch.full_like(self.h, 0.1)
            elif optimizer == "Adam":
                self.h_s1, self.h_s2 = torch.zeros_like(self.h), torch.zeros_like(self.h)

    def _one(self, co, u, i, y, seed, epoch, first, neg_ratio, batch, reg, lo):
        eng, lib = self.engine, self.engine.lib
        check(lib.crb_shard_step_compute_pointwise(eng.h, self.kind, C.byref(self.P.c), C.byref(self.shard), ptr(self.h), C.byref(co), self.loss_kind,
                                                   ptr(u), ptr(i), ptr(y), seed, epoch, first, neg_ratio or 1, batch, float(reg), lo, eng.stream))
        self.barrier()
        check(lib.crb_shard_apply_inbox(eng.h, C.byref(self.shard), C.byref(co), eng.stream))
        if self.h is not None:
            check(lib.crb_shard_apply_dense(eng.h, C.byref(self.shard), C.byref(co), ptr(self.h), ptr(self.h_s1), ptr(self.h_s2), self.dim, eng.stream))
        self.barrier()

    def step(self, reg, neg_ratio=None, seed=0, epoch=0, first=0, batch=None, feed=None, loss_out=None):
        """One synchronous step over the union batch.  feed = (u_local, i_global, y) or None to sample on the device."""
        eng = self.engine
        self.opt.t += 1
        co = self.opt.c(self.opt.t)
        batch = self.batch if batch is None else batch
        u = i = y = None
        if feed is not None:
            u, i = (eng._feed_i32(x) for x in feed[:2])
            y = feed[2] if isinstance(feed[2], torch.Tensor) else np.ascontiguousarray(np.asarray(feed[2]), dtype=np.float32)
            batch = len(u)
        host = np.zeros(1, dtype=np.float64) if loss_out is None else None
        self._one(co, u, i, y, seed, epoch, first, neg_ratio, batch, reg, ptr(host) if loss_out is None else ptr(loss_out))
        return float(host[0]) if loss_out is None else None

    def run_steps(self, n_steps, reg, neg_ratio, seed, epoch, first=0, batch=None, loss_out=None, bounds=None, **_):
        batch = self.batch if batch is None else batch
        if bounds is None:
            bounds = [first + k * batch for k in range(n_steps + 1)]
        host = np.zeros(1, dtype=np.float64)
        for k in range(n_steps):
            self.opt.t += 1
            co = self.opt.c(self.opt.t)
            lo = ptr(host) if loss_out is None else ptr(loss_out[k:k + 1])
            self._one(co, None, None, None, seed, epoch, bounds[k], neg_ratio, bounds[k + 1] - bounds[k], reg, lo)


# ---------------------------------------------------------------------------------------------- evaluation across the item shards
def transpose_history(seen_rowptr, seen_cols, u_lo, n_users_total, world, rank, group=None):
    """Each rank holds the histories of ITS users (global item ids).  Evaluation needs the opposite cut: for ALL users, the seen
    items that live on THIS rank's item shard (owner = item % world), as local rows.  One all-to-all of (user, local row) pairs at
    set-up; works on CPU tensors with gloo (tests) and CUDA tensors with NCCL.  -> (rowptr int64 [n_users_total+1], cols int32)."""
    out_dev = seen_cols.device
    if world > 1 and seen_cols.is_cuda and _is_gloo(group):   # ranks sharing one device (tests): the exchange runs on host copies
        seen_rowptr, seen_cols = seen_rowptr.cpu(), seen_cols.cpu()
    dev = seen_cols.device
    counts = (seen_rowptr[1:] - seen_rowptr[:-1])
    users = torch.repeat_interleave(torch.arange(counts.numel(), device=dev, dtype=torch.int64) + u_lo, counts)
    cols = seen_cols.to(torch.int64)
    owner = cols % world
    order = torch.argsort(owner, stable=True)
    send_counts = torch.bincount(owner, minlength=world)
    send_u, send_i = users[order].contiguous(), (cols // world)[order].contiguous()
    recv_counts = torch.empty_like(send_counts)
    if world > 1:
        dist.all_to_all_single(recv_counts, send_counts, group=group)
    else:
        recv_counts.copy_(send_counts)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    recv_u = torch.empty(sum(rc), dtype=torch.int64, device=dev)
    recv_i = torch.empty(sum(rc), dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_to_all_single(recv_u, send_u, rc, sc, group=group)
        dist.all_to_all_single(recv_i, send_i, rc, sc, group=group)
    else:
        recv_u.copy_(send_u); recv_i.copy_(send_i)
    stride = int(recv_i.max().item()) + 1 if recv_i.numel() else 1
    key = torch.unique(recv_u * stride + recv_i)  # sorted by (user, local row)
    ku = torch.div(key, stride, rounding_mode="floor")
    rowptr = torch.zeros(n_users_total + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(ku, minlength=n_users_total), 0, out=rowptr[1:])
    return rowptr.to(out_dev), (key - ku * stride).to(torch.int32).to(out_dev)


def merge_topk(ids_per_rank, scores_per_rank, K, ascending=False):
    """Exact global top-K from every shard's exact local top-K: candidates ordered by (score, global id) with the documented tie
    rule.  ids/scores: lists of [n, K] tensors (ids global, -1 = padding).  Pure tensor logic (CPU-testable)."""
    ids = torch.cat(ids_per_rank, dim=1).to(torch.int64)
    sc = torch.cat(scores_per_rank, dim=1).to(torch.float32)
    pad = ids < 0
    worst = float("inf") if ascending else float("-inf")
    sc = torch.where(pad, torch.full_like(sc, worst), sc)
    ids_key = torch.where(pad, torch.full_like(ids, 1 << 40), ids)
    order = torch.argsort(ids_key, dim=1, stable=True)               # id ascending first ...
    ids, sc = torch.gather(ids, 1, order), torch.gather(sc, 1, order)
    order = torch.argsort(sc, dim=1, descending=not ascending, stable=True)[:, :K]   # ... then a stable sort by score
    return torch.gather(ids, 1, order).to(torch.int32), torch.gather(sc, 1, order)


class ShardedEval(object):
    """Full-rank top-K of this rank's users (test_model_rs across ranks).  Two layouts:

    mode="replicate" (default): the item shards are all-gathered once per evaluation into a full copy of Q on every rank
        (2M x 128 fp32 = 1 GB, a few ms over NVLink) and every rank ranks ITS OWN users with the single-GPU tensor-core path
        against its own history CSR.  No per-batch collective, no merge; throughput scales with the number of ranks, and the ids
        are those of the single-GPU path by construction.
    mode="shard": for catalogues that do not fit one GPU.  The users' vectors are broadcast batch by batch, every rank ranks
        them against ITS item shard (seen items masked from the transposed history), and the G x K candidates per user are merged
        at the user's owner."""

    def __init__(self, model, seen_rowptr, seen_cols, kind=_lib.SCORE_DOT, mode="replicate"):
        assert mode in ("replicate", "shard")
        self.m, self.kind, self.mode = model, kind, mode
        self.world, self.rank = model.world, model.rank
        dev = model.engine.device
        if mode == "replicate":
            self.eng = None   # the model's own engine already holds this rank's history (local user rows, global item ids)
            return
        rowptr, cols = transpose_history(seen_rowptr.to(dev), seen_cols.to(dev), model.u_lo, model.n_users, self.world, self.rank, model.group)
        self.eng = Engine(dev.index)   # a second handle on the same device: its history is the transposed one
        pu = torch.zeros(1, dtype=torch.int32, device=dev)
        self.eng.set_history_arrays(model.n_users, model.q_rows, pu[:0], pu[:0], rowptr, cols)

    def _topk_replicated(self, K, batch_users, exact, limit):
        m, dev = self.m, self.m.engine.device
        Qfull = m.gather_Q()   # flushes pending Adam decay, then one all-gather
        m.engine.adam_flush(m.P, m.opt)
        n = m.u_hi - m.u_lo if limit is None else min(m.u_hi - m.u_lo, limit)
        out = torch.full((m.u_hi - m.u_lo, K), -1, dtype=torch.int32, device=dev)
        for a in range(0, n, batch_users):
            b = min(n, a + batch_users)
            users = torch.arange(a, b, dtype=torch.int32, device=dev)
            out[a:b] = m.engine.score_topk(self.kind, m.P.w, Qfull, users, K, exact=exact)
        return out

    def topk(self, K, batch_users=1 << 16, exact=False, limit=None):
        """limit: evaluate only the first `limit` users of every rank (benchmarks)."""
        if self.mode == "replicate":
            return self._topk_replicated(K, batch_users, exact, limit)
        m, dev = self.m, self.m.engine.device
        m.flush()
        out = torch.full((m.u_hi - m.u_lo, K), -1, dtype=torch.int32, device=dev)
        Qw = m.q["w"].tensor
        for owner in range(self.world):
            lo, hi = user_range(m.n_users, owner, self.world)
            if limit is not None:
                hi = min(hi, lo + limit)
            for a in range(lo, hi, batch_users):
                b = min(hi, a + batch_users)
                rows = m.P.w[a - lo:b - lo].contiguous() if owner == self.rank else torch.empty(b - a, m.dim, device=dev)
                if self.world > 1:
                    broadcast_dev(rows, owner, m.group)
                local_u = torch.arange(b - a, dtype=torch.int32, device=dev)
                hist_u = torch.arange(a, b, dtype=torch.int32, device=dev)
                ids, sc = self.eng.score_topk(self.kind, rows, Qw, local_u, K, hist_users=hist_u, exact=exact, n_items=m.q_rows, return_scores=True)
                gids = torch.where(ids >= 0, ids * self.world + self.rank, ids)   # local row -> global item id
                if self.world > 1:
                    all_ids, all_sc = gather_dev(gids, owner, m.group), gather_dev(sc, owner, m.group)
                else:
                    all_ids, all_sc = [gids], [sc]
                if owner == self.rank:
                    out[a - lo:b - lo] = merge_topk(all_ids, all_sc, K, ascending=self.kind == _lib.SCORE_SQDIST)[0]
        return out
