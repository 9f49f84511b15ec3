# coding: utf-8
" Bayesian Personalized Ranking (2009) -- mirror of the reference model/ranking/BPR.py. "
import numpy as np
import torch

from .. import RankingRecommender as _rr
from ... import _lib
from ...engine import Table


class BPR(_rr.RankingRecommender):
    def __init__(self, sess, data, configs, logger):
        super(BPR, self).__init__(sess, data, configs, logger)
        self.embed_size, self.reg = int(configs['embed_size']), float(configs['reg'])
        logger.info(' model_params: embed_size=%d, reg=%s' % (self.embed_size, self.reg) + ', ' + self.model_params)

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
        self._create_params(init)

    def _variables(self):   # BPR.py:53-58
        return {'BPR_params/P': self.P.w, 'BPR_params/Q': self.Q.w}

    # sess.run([self.train, self.loss], ...) for every batch of the epoch, sampler fused in (BPR.py:31-44)
    def _train_epoch_pairwise(self, epoch, n_rows, n_batches, losses):
        self.engine.train_epoch_bpr(self.P, self.Q, self.optimizer, self.seed, epoch, 0, self.batch_size, n_batches,
                                    self.neg_ratio, self.reg, losses)

    def train_step(self, u_idx, i_idx, j_idx, loss_out=None):
        """Feed-style single step (the reference's `sess.run([train, loss], {u_idx, i_idx, j_idx})`)."""
        return self.engine.train_step_bpr(self.P, self.Q, self.optimizer, u_idx, i_idx, j_idx, self.reg, loss_out=loss_out)

    def _before_eval(self):
        self.engine.adam_flush(self.P, self.optimizer)
        self.engine.adam_flush(self.Q, self.optimizer)

    def _score_spec(self):  # BPR.py:46-51
        return _lib.SCORE_DOT, self.P.w, self.Q.w, None
