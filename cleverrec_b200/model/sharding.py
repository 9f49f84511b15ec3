# coding: utf-8
""" Multi-GPU plumbing shared by the model classes that run under WORLD_SIZE > 1 (BPR, MF, GMF): one process per GPU of one box
(torchrun), users partitioned into contiguous ranges, the item table row-sharded over NVLink peer memory (cleverrec_b200/dist.py).
The reference has no distributed code; the contract here is that the SAME class, methods, returns and log lines are reached --
every step is one synchronous step on the union of the ranks' batches (batch_size stays the GLOBAL batch,
model/RankingRecommender.py:38-46), and every rank returns the full HR / MRR / NDCG lists in self.test_users order (:243-247). """
import math
import os
from collections import defaultdict

import numpy as np
import torch


class ShardedModelMixin(object):
    supports_sharding = True
    _sampler_kind = 'pairwise'     # which epoch the class trains on under WORLD_SIZE > 1

    @property
    def sharded(self):
        return self.world > 1

    def _install_history(self):
        if not self.sharded:
            return super(ShardedModelMixin, self)._install_history()
        import torch.distributed as dist
        from ..dist import shard_history, user_range
        if not dist.is_initialized():
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            if os.environ.get('CRB_SHARED_DEVICE', '0') == '1':
                dist.init_process_group('gloo')     # ranks share one GPU: NCCL refuses that, the data path (CUDA IPC) does not care
            else:
                dist.init_process_group('nccl', device_id=self.engine.device)
        self.u_lo, self.u_hi = user_range(self.data.user_nums, self.rank, self.world)
        mine, n_local = shard_history(self.data.ui_train, self.data.user_nums, self.rank, self.world)
        self.engine.set_history(mine, n_local, self.data.item_nums)   # local user rows, global item ids

    def _full_init(self, names_shapes, init):
        """Every rank draws the same full variables in the single-GPU order (same seed): identical to the single-GPU initialisation."""
        full = {}
        for name, shape in names_shapes:
            full[name] = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer(shape)
        return full

    def _train_model_sharded(self):
        import torch.distributed as dist
        from ..dist import all_reduce_dev
        if self.sampler_mode == 'numpy_stream':
            raise NotImplementedError("sampler=numpy_stream reproduces ONE process's np.random stream; it is single-GPU only")
        rows = self.engine.epoch_rows(self.neg_ratio, self._sampler_kind)
        every = [None] * self.world
        dist.all_gather_object(every, int(rows))
        total, fewest = sum(every), min(every)
        n_steps = math.ceil(total / self.batch_size)                # the global epoch in global batches (RankingRecommender.py:38-39)
        if fewest < n_steps:                                         # decided from gathered values: every rank raises, none hangs
            raise ValueError('a rank holds %d training rows for %d steps: too little data for %d ranks' % (fewest, n_steps, self.world))
        bounds = [rows * k // n_steps for k in range(n_steps + 1)]   # this rank's share of every union batch
        losses = torch.zeros(n_steps, dtype=torch.float64, device=self.engine.device)
        self._shm.run_steps(n_steps, self.reg, self.neg_ratio, self.seed, self.epoch, bounds=bounds, loss_out=losses)
        self._shm.check()                                            # barrier time-outs / sampler give-ups surface here, once per epoch
        all_reduce_dev(losses)                                       # the step's loss is the sum over the union batch
        self.epoch += 1
        self._Qfull = None
        return float(losses.sum().item()) / n_steps

    def _sharded_eval(self, run):
        import torch.distributed as dist
        everyone = self.test_users
        mine = [u for u in everyone if self.u_lo <= u < self.u_hi]
        self.test_users, self.test_batches, self._test_cache = mine, math.ceil(len(mine) / self.batch_size_t), None
        try:
            self._before_eval()
            HR, MRR, NDCG = run() if mine else (defaultdict(list), defaultdict(list), defaultdict(list))
        finally:
            self.test_users, self.test_batches, self._test_cache = everyone, math.ceil(len(everyone) / self.batch_size_t), None
        parts = [None] * self.world
        dist.all_gather_object(parts, (mine, dict(HR), dict(MRR), dict(NDCG)))
        per_user = {}
        for users, hr, mrr, ndcg in parts:
            for k, u in enumerate(users):
                per_user[u] = {kid: (hr[kid][k], mrr[kid][k], ndcg[kid][k]) for kid in hr}
        out = (defaultdict(list), defaultdict(list), defaultdict(list))
        for u in everyone:                       # one float per test user in self.test_users order (RankingRecommender.py:243-247)
            for kid, vals in per_user[u].items():
                for m in range(3):
                    out[m][kid].append(vals[m])
        return out

    def test_model_loo(self):
        run = super(ShardedModelMixin, self).test_model_loo
        return self._sharded_eval(run) if self.sharded else run()

    def test_model_rs(self):
        run = super(ShardedModelMixin, self).test_model_rs
        return self._sharded_eval(run) if self.sharded else run()

    def _pair_users(self, u_idx):
        return u_idx - self.u_lo if self.sharded else super(ShardedModelMixin, self)._pair_users(u_idx)

    def _fullrank_users(self, cur_users):
        if not self.sharded:
            return super(ShardedModelMixin, self)._fullrank_users(cur_users)
        rows = np.asarray(cur_users, dtype=np.int32) - self.u_lo
        return rows, None   # the engine's history is keyed by local user rows too

    def _gather_item_table(self):
        if self._Qfull is None:
            self._Qfull = self._shm.gather_Q()   # flushes pending Adam decay, then one all-gather of the item shards
        return self._Qfull

    def save_model(self, step=None):
        if not self.sharded:
            return super(ShardedModelMixin, self).save_model(step)
        from ..utils.tools import save_checkpoint
        variables = self._variables()          # collective: every rank takes part, rank 0 writes
        if self.rank != 0:
            return None
        return save_checkpoint(os.path.join(self.saved_model_dir, self.model), self.model, variables, step)
