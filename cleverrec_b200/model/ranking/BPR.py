# coding: utf-8
" Bayesian Personalized Ranking (2009) -- mirror of the reference model/ranking/BPR.py. "
import numpy as np
import torch

from .. import RankingRecommender as _rr
from ... import _lib
from ...engine import Table


class BPR(_rr.RankingRecommender):
    """Single GPU: the fused epoch of csrc/train.cu.  Under torchrun (WORLD_SIZE > 1, one process per GPU of one box) the same class
    runs the multi-GPU path of cleverrec_b200/dist.py behind the same methods: users are partitioned into contiguous ranges, the
    item table is row-sharded and read / updated over NVLink peer memory, every step is one synchronous step on the union of the
    ranks' batches (batch_size stays the GLOBAL batch), and evaluation ranks each rank's own test users against the all-gathered
    item table.  Every rank returns the full metric lists."""
    supports_sharding = True

    def __init__(self, sess, data, configs, logger):
        super(BPR, self).__init__(sess, data, configs, logger)
        self.embed_size, self.reg = int(configs['embed_size']), float(configs['reg'])
        logger.info(' model_params: embed_size=%d, reg=%s' % (self.embed_size, self.reg) + ', ' + self.model_params)
        if self.loss_func != 'bpr' or self.is_pairwise != 'True':
            # BPR.py:42 calls get_loss(self.loss_func, ui - uj) without margin / logits: only 'bpr' builds in the reference
            raise ValueError("BPR is defined with is_pairwise=True, loss_func=bpr (conf/BPR.properties), got %r / %r" % (self.is_pairwise, self.loss_func))

    # ------------------------------------------------------------------------------------------ multi-GPU plumbing
    @property
    def sharded(self):
        return self.world > 1

    def _install_history(self):
        if not self.sharded:
            return super(BPR, self)._install_history()
        import os
        import torch.distributed as dist
        from ...dist import shard_history, user_range
        if not dist.is_initialized():
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            if os.environ.get('CRB_SHARED_DEVICE', '0') == '1':
                dist.init_process_group('gloo')     # ranks share one GPU: NCCL refuses that, the data path (CUDA IPC) does not care
            else:
                dist.init_process_group('nccl', device_id=self.engine.device)
        self.u_lo, self.u_hi = user_range(self.data.user_nums, self.rank, self.world)
        mine, n_local = shard_history(self.data.ui_train, self.data.user_nums, self.rank, self.world)
        self.engine.set_history(mine, n_local, self.data.item_nums)   # local user rows, global item ids

    def _create_params_sharded(self, init):
        from ...dist import ShardedBPR
        shapes = {'P': [self.data.user_nums, self.embed_size], 'Q': [self.data.item_nums, self.embed_size]}
        full = {}
        for name in ('P', 'Q'):   # every rank draws the same full tables (same seed): identical to the single-GPU initialisation
            full[name] = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer(shapes[name])
        per_rank = -(-self.batch_size // self.world)
        self._shm = ShardedBPR(self.engine, self.data.user_nums, self.data.item_nums, self.embed_size, self.optimizer.kind, self.optimizer.lr,
                               self.optimizer.adam_mode, per_rank, init_P=full['P'][self.u_lo:self.u_hi], init_Q=full['Q'])
        self.optimizer = self._shm.opt
        self.P = self._shm.P          # this rank's user rows
        self.tables = {'P': self.P}
        self._Qfull = None

    def train_model(self):
        if not self.sharded:
            return super(BPR, self).train_model()
        import math
        import torch.distributed as dist
        if self.sampler_mode == 'numpy_stream':
            raise NotImplementedError("sampler=numpy_stream reproduces ONE process's np.random stream; it is single-GPU only")
        from ...dist import all_reduce_dev
        rows = self.engine.epoch_rows(self.neg_ratio, 'pairwise')
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
        import math
        import torch.distributed as dist
        from collections import defaultdict
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
        if not self.sharded:
            return super(BPR, self).test_model_loo()
        return self._sharded_eval(super(BPR, self).test_model_loo)

    def test_model_rs(self):
        if not self.sharded:
            return super(BPR, self).test_model_rs()
        return self._sharded_eval(super(BPR, self).test_model_rs)

    def _pair_users(self, u_idx):
        return u_idx - self.u_lo if self.sharded else u_idx

    def _fullrank_users(self, cur_users):
        if not self.sharded:
            return super(BPR, self)._fullrank_users(cur_users)
        rows = np.asarray(cur_users, dtype=np.int32) - self.u_lo
        return rows, None   # the engine's history is keyed by local user rows too

    def _create_params(self, init=None):
        """BPR.py:23-29.  `init` = {'P': array, 'Q': array} injects initial tables (parity runs)."""
        dev = self.engine.device
        shapes = {'P': [self.data.user_nums, self.embed_size], 'Q': [self.data.item_nums, self.embed_size]}
        self.tables = {}
        for name in ('P', 'Q'):
            w = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer(shapes[name])
            assert list(w.shape) == shapes[name]
            self.tables[name] = Table(w.to(dev).contiguous(), self.optimizer.kind, self.optimizer.adam_mode)
        self.P, self.Q = self.tables['P'], self.tables['Q']

    def build_model(self, init=None):
        if self.sharded:
            return self._create_params_sharded(init)
        self._create_params(init)

    def _variables(self):   # BPR.py:53-58
        if self.sharded:
            return {'BPR_params/P': self._shm.gather_P(), 'BPR_params/Q': self._shm.gather_Q()}
        return {'BPR_params/P': self.P.w, 'BPR_params/Q': self.Q.w}

    def save_model(self, step=None):
        if not self.sharded:
            return super(BPR, self).save_model(step)
        import os
        from ...utils.tools import save_checkpoint
        variables = self._variables()          # collective: every rank takes part, rank 0 writes
        if self.rank != 0:
            return None
        return save_checkpoint(os.path.join(self.saved_model_dir, self.model), self.model, variables, step)

    # sess.run([self.train, self.loss], ...) for every batch of the epoch, sampler fused in (BPR.py:31-44)
    def _train_epoch_pairwise(self, epoch, n_rows, n_batches, losses):
        self.engine.train_epoch_bpr(self.P, self.Q, self.optimizer, self.seed, epoch, 0, self.batch_size, n_batches,
                                    self.neg_ratio, self.reg, losses)

    def train_step(self, u_idx, i_idx, j_idx, loss_out=None):
        """Feed-style single step (the reference's `sess.run([train, loss], {u_idx, i_idx, j_idx})`).  Multi-GPU: u_idx are this
        rank's LOCAL user rows, i_idx / j_idx global item ids; returns this rank's part of the loss."""
        if self.sharded:
            return self._shm.step(self.reg, feed=(u_idx, i_idx, j_idx), loss_out=loss_out)
        return self.engine.train_step_bpr(self.P, self.Q, self.optimizer, u_idx, i_idx, j_idx, self.reg, loss_out=loss_out)

    def _before_eval(self):
        if self.sharded:
            if self._Qfull is None:
                self._Qfull = self._shm.gather_Q()   # flushes pending Adam decay, then one all-gather of the item shards
            return
        self.engine.adam_flush(self.P, self.optimizer)
        self.engine.adam_flush(self.Q, self.optimizer)

    def _score_spec(self):  # BPR.py:46-51
        if self.sharded:
            return _lib.SCORE_DOT, self.P.w, self._Qfull, None
        return _lib.SCORE_DOT, self.P.w, self.Q.w, None
