# coding: utf-8
" Generalized Matrix Factorization (2017 NCF) -- mirror of the reference model/ranking/GMF.py. "
import numpy as np
import torch

from .. import RankingRecommender as _rr
from ... import _lib
from ...engine import Table

_LOSS = {'cross_entropy': _lib.LOSS_CROSS_ENTROPY, 'square': _lib.LOSS_SQUARE}


class GMF(_rr.RankingRecommender):
    score_kind = _lib.SCORE_GMF

    def __init__(self, sess, data, configs, logger):
        super(GMF, self).__init__(sess, data, configs, logger)
        # conf/GMF.properties defines reg_gmf while GMF.py:12 reads 'reg' (SURVEY 2.3): accept both
        reg = configs['reg'] if 'reg' in configs else configs['reg_gmf']
        self.embed_size, self.reg = int(configs['embed_size']), float(reg)
        if self.loss_func not in _LOSS:
            raise ValueError('pointwise loss_func must be cross_entropy or square, got %r' % self.loss_func)
        logger.info(' model_params: embed_size=%d, reg=%s' % (self.embed_size, self.reg) + ', ' + self.model_params)

    def _create_params(self, init=None):
        """GMF.py:22-31.  `init` = {'P','Q','h'} injects initial values (parity runs)."""
        dev = self.engine.device
        shapes = {'P': [self.data.user_nums, self.embed_size], 'Q': [self.data.item_nums, self.embed_size]}
        for name in ('P', 'Q'):
            w = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer(shapes[name])
            setattr(self, name, Table(w.to(dev).contiguous(), self.optimizer.kind, self.optimizer.adam_mode))
        self.h_gmf = self.h_s1 = self.h_s2 = None
        if self.score_kind == _lib.SCORE_GMF:
            h = torch.as_tensor(np.asarray(init['h']), dtype=torch.float32) if init and 'h' in init else self.initializer([self.embed_size])
            self.h_gmf = h.to(dev).contiguous()
            if self.optimizer.kind == 'Adagrad':
                self.h_s1 = torch.full_like(self.h_gmf, 0.1)
            elif self.optimizer.kind == 'Adam':
                self.h_s1, self.h_s2 = torch.zeros_like(self.h_gmf), torch.zeros_like(self.h_gmf)

    def build_model(self, init=None):
        self._create_params(init)

    def _variables(self):   # GMF.py:59-64 (MF, which has no reference source, saves under its own class name)
        out = {'%s_params/P' % self.model: self.P.w, '%s_params/Q' % self.model: self.Q.w}
        if self.h_gmf is not None:
            out['%s_params/h_gmf' % self.model] = self.h_gmf
        return out

    def train_step(self, u_idx, i_idx, y, loss_out=None):
        """Feed-style step: sess.run([train, loss], {u_idx, i_idx, y})  (GMF.py:45-49)."""
        return self.engine.train_step_pointwise(self.score_kind, self.P, self.Q, self.optimizer, u_idx, i_idx, y, self.reg,
                                                _LOSS[self.loss_func], self.h_gmf, self.h_s1, self.h_s2, loss_out=loss_out)

    def _train_epoch_pointwise(self, epoch, n_rows, n_batches, losses):
        # the whole loop of RankingRecommender.py:48-60 in one call (sampler fused in, no host round trip per batch)
        self.engine.train_epoch_pointwise(self.score_kind, self.P, self.Q, self.optimizer, self.seed, epoch, 0, self.batch_size, n_batches,
                                          self.neg_ratio, self.reg, _LOSS[self.loss_func], self.h_gmf, self.h_s1, self.h_s2, loss_out=losses)

    def _before_eval(self):
        self.engine.adam_flush(self.P, self.optimizer)
        self.engine.adam_flush(self.Q, self.optimizer)

    def _score_spec(self):  # GMF.py:51-57: ranking on the logit (sigmoid is monotone; ties at saturation stated in DESIGN.md)
        return self.score_kind, self.P.w, self.Q.w, self.h_gmf
