# coding: utf-8
" Bayesian Personalized Ranking (2009) -- mirror of the reference model/ranking/BPR.py. "
import numpy as np
import torch

from .. import RankingRecommender as _rr
from ..sharding import ShardedModelMixin
from ... import _lib
from ...engine import Table


class BPR(ShardedModelMixin, _rr.RankingRecommender):
    """Single GPU: the fused epoch of csrc/train.cu.  Under torchrun (WORLD_SIZE > 1, one process per GPU of one box) the same class
    runs the multi-GPU path of cleverrec_b200/dist.py behind the same methods (cleverrec_b200/model/sharding.py): users are
    partitioned into contiguous ranges, the item table is row-sharded and read / updated over NVLink peer memory, every step is one
    synchronous step on the union of the ranks' batches (batch_size stays the GLOBAL batch), and evaluation ranks each rank's own
    test users against the all-gathered item table.  Every rank returns the full metric lists."""
    _sampler_kind = 'pairwise'

    def __init__(self, sess, data, configs, logger):
        super(BPR, self).__init__(sess, data, configs, logger)
        self.embed_size, self.reg = int(configs['embed_size']), float(configs['reg'])
        logger.info(' model_params: embed_size=%d, reg=%s' % (self.embed_size, self.reg) + ', ' + self.model_params)
        if self.loss_func != 'bpr' or self.is_pairwise != 'True':
            # BPR.py:42 calls get_loss(self.loss_func, ui - uj) without margin / logits: only 'bpr' builds in the reference
            raise ValueError("BPR is defined with is_pairwise=True, loss_func=bpr (conf/BPR.properties), got %r / %r" % (self.is_pairwise, self.loss_func))

    # ------------------------------------------------------------------------------------------ multi-GPU plumbing
    def _create_params_sharded(self, init):
        from ...dist import ShardedBPR
        full = self._full_init((('P', [self.data.user_nums, self.embed_size]), ('Q', [self.data.item_nums, self.embed_size])), init)
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
        return self._train_model_sharded()

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
            self._gather_item_table()
            return
        self.engine.adam_flush(self.P, self.optimizer)
        self.engine.adam_flush(self.Q, self.optimizer)

    def _score_spec(self):  # BPR.py:46-51
        if self.sharded:
            return _lib.SCORE_DOT, self.P.w, self._Qfull, None
        return _lib.SCORE_DOT, self.P.w, self.Q.w, None
