"""Multi-GPU (one process per GPU, one box) BPR training over NVLink peer memory -- the host side of csrc/train_sharded.cu.

Partitioning: users are split into contiguous ranges, one per rank (rows of P, their histories and their sampling
are local); the item table Q is row-sharded by `item % world` and every rank maps every shard and every gradient inbox through
CUDA IPC.  torch.distributed (NCCL) is used for exactly two things: exchanging the 64-byte IPC handles once, and the per-step
barrier (a one-element all-reduce enqueued on the compute stream).  There is no data-path collective."""
import ctypes as C
import math

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


class ShardedBPR(object):
    """BPR (model/ranking/BPR.py:31-44) with P partitioned by user and Q row-sharded over the ranks of one box."""

    def __init__(self, engine, n_users, n_items, dim, optimizer, lr, adam_mode, batch, init_P=None, init_Q=None, seed=0, group=None):
        self.engine, self.group = engine, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world <= _lib.MAX_RANKS
        self.n_users, self.n_items, self.dim, self.batch = n_users, n_items, dim, batch
        self.u_lo, self.u_hi = user_range(n_users, self.rank, self.world)
        self.opt = Optimizer(optimizer, lr, adam_mode=adam_mode)
        dev = engine.device
        g = torch.Generator(device=dev).manual_seed(seed)
        # user rows: ordinary torch memory
        if init_P is None:
            init_P = torch.randn(self.u_hi - self.u_lo, dim, device=dev, generator=g) * 0.01
        self.P = Table(torch.as_tensor(init_P, dtype=torch.float32).to(dev).contiguous(), optimizer, adam_mode)
        # item shard + inbox: IPC-exportable memory
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
        # every rank may receive up to 2 * batch * world gradients per step (2 per triplet); hubs make the split uneven
        self.inbox_cap = int(2 * batch * min(self.world, 2)) + 32 * 4096 * self.world  # + slack for partly used reservations
        self.inbox = {"grad": _DeviceBuffer(engine, (self.inbox_cap, dim), torch.float32), "row": _DeviceBuffer(engine, (self.inbox_cap,), torch.int32),
                      "key": _DeviceBuffer(engine, (self.inbox_cap,), torch.int32), "cnt": _DeviceBuffer(engine, (4,), torch.int32)}
        self.inbox["row"].tensor.fill_(-1)   # every slot is a hole until a gradient is written into it
        self._map_peers()
        self._flag = torch.zeros(1, device=dev)
        torch.cuda.synchronize()
        self.barrier()

    def _map_peers(self):
        mine = {"rows": self.q_rows}
        mine.update({"q_" + k: b.export() for k, b in self.q.items()})
        mine.update({"in_" + k: b.export() for k, b in self.inbox.items()})
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        self._opened = []
        sh = CrbShard()
        sh.n_ranks, sh.rank, sh.inbox_cap = self.world, self.rank, self.inbox_cap

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
            sh.inbox_grad[r], sh.inbox_row[r] = pointer("grad", self.inbox), pointer("row", self.inbox)
            sh.inbox_key[r], sh.inbox_cnt[r] = pointer("key", self.inbox), pointer("cnt", self.inbox)
        self.shard = sh

    def barrier(self):
        """Cross-rank barrier ordered with the compute stream (no host synchronisation)."""
        if self.world > 1:
            dist.all_reduce(self._flag, group=self.group)

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

    def inbox_overflowed(self):
        v = C.c_int32()
        check(self.engine.lib.crb_shard_inbox_overflow(self.engine.h, C.byref(self.shard), C.byref(v), self.engine.stream))
        return bool(v.value)

    def flush(self):
        """Bring CRB_ADAM_TF1 tables up to date before reading them."""
        self.engine.adam_flush(self.P, self.opt)
        if self.opt.kind == "Adam" and self.opt.adam_mode == "tf1" and self.opt.t > 0:
            co = self.opt.c(self.opt.t)
            check(self.engine.lib.crb_adam_flush(self.engine.h, C.byref(self.shard.q[self.rank]), C.byref(co), self.engine.stream))

    def gather_Q(self):
        """Full item table on every rank (tests / evaluation set-up)."""
        self.flush()
        parts = [torch.zeros(shard_rows(self.n_items, r, self.world), self.dim, device=self.engine.device) for r in range(self.world)]
        if self.world > 1:
            dist.all_gather(parts, self.q["w"].tensor.contiguous(), group=self.group) if len({p.shape for p in parts}) == 1 else self._gather_ragged(parts)
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
            dist.broadcast(parts[r], src=r, group=self.group)

    def close(self):
        torch.cuda.synchronize()
        self.barrier()
        torch.cuda.synchronize()
        for p in self._opened:
            self.engine.lib.crb_ipc_close(self.engine.h, p)
        for b in list(self.q.values()) + list(self.inbox.values()):
            b.tensor = None
            self.engine.lib.crb_free(self.engine.h, b.ptr)
