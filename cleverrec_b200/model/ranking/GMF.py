# coding: utf-8
" Generalized Matrix Factorization (2017 NCF) -- mirror of the reference model/ranking/GMF.py. "
import numpy as np
import torch

from .. import RankingRecommender as _rr
from ..sharding import ShardedModelMixin
from ... import _lib
from ...engine import Table

_LOSS = {'cross_entropy': _lib.LOSS_CROSS_ENTROPY, 'square': _lib.LOSS_SQUARE}


class GMF(ShardedModelMixin, _rr.RankingRecommender):
    """Single GPU: csrc/train.cu's pointwise step.  Under torchrun (WORLD_SIZE > 1) the same class (and MF, its subclass) runs the
    multi-GPU path (cleverrec_b200/dist.py::ShardedPointwise): users partitioned, item table row-sharded over NVLink peer memory,
    h replicated with its gradient summed over the ranks in rank order."""
    score_kind = _lib.SCORE_GMF
    _sampler_kind = 'pointwise'

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

    def _create_params_sharded(self, init):
        from ...dist import ShardedPointwise
        shapes = [('P', [self.data.user_nums, self.embed_size]), ('Q', [self.data.item_nums, self.embed_size])]
        if self.score_kind == _lib.SCORE_GMF:
            shapes.append(('h', [self.embed_size]))
        full = self._full_init(shapes, init)
        per_rank = -(-self.batch_size // self.world)
        self._shm = ShardedPointwise(self.engine, self.data.user_nums, self.data.item_nums, self.embed_size, self.optimizer.kind, self.optimizer.lr,
                                     self.optimizer.adam_mode, per_rank, kind=self.score_kind, loss_kind=_LOSS[self.loss_func],
                                     init_P=full['P'][self.u_lo:self.u_hi], init_Q=full['Q'], init_h=full.get('h'))
        self.optimizer = self._shm.opt
        self.P = self._shm.P
        self.h_gmf, self.h_s1, self.h_s2 = self._shm.h, self._shm.h_s1, self._shm.h_s2
        self._Qfull = None

    def build_model(self, init=None):
        if self.sharded:
            return self._create_params_sharded(init)
        self._create_params(init)

    def train_model(self):
        if not self.sharded:
            return super(GMF, self).train_model()
        if self.is_pairwise == 'True':
            raise NotImplementedError('%s under WORLD_SIZE > 1 trains pointwise (is_pairwise=False)' % self.model)
        return self._train_model_sharded()

    def _variables(self):   # GMF.py:59-64 (MF, which has no reference source, saves under its own class name)
        if self.sharded:
            out = {'%s_params/P' % self.model: self._shm.gather_P(), '%s_params/Q' % self.model: self._shm.gather_Q()}
        else:
            out = {'%s_params/P' % self.model: self.P.w, '%s_params/Q' % self.model: self.Q.w}
        if self.h_gmf is not None:
            out['%s_params/h_gmf' % self.model] = self.h_gmf
        return out

    def train_step(self, u_idx, i_idx, y, loss_out=None):
        """Feed-style step: sess.run([train, loss], {u_idx, i_idx, y})  (GMF.py:45-49).  Multi-GPU: u_idx are this rank's LOCAL user
        rows, i_idx global item ids; returns this rank's part of the loss."""
        if self.sharded:
            return self._shm.step(self.reg, feed=(u_idx, i_idx, y), loss_out=loss_out)
        return self.engine.train_step_pointwise(self.score_kind, self.P, self.Q, self.optimizer, u_idx, i_idx, y, self.reg,
                                                _LOSS[self.loss_func], self.h_gmf, self.h_s1, self.h_s2, loss_out=loss_out)

    def _train_epoch_pointwise(self, epoch, n_rows, n_batches, losses):
        # the whole loop of RankingRecommender.py:48-60 in one call (sampler fused in, no host round trip per batch)
        self.engine.train_epoch_pointwise(self.score_kind, self.P, self.Q, self.optimizer, self.seed, epoch, 0, self.batch_size, n_batches,
                                          self.neg_ratio, self.reg, _LOSS[self.loss_func], self.h_gmf, self.h_s1, self.h_s2, loss_out=losses)

    def _before_eval(self):
        if self.sharded:
            self._gather_item_table()
            return
        self.engine.adam_flush(self.P, self.optimizer)
        self.engine.adam_flush(self.Q, self.optimizer)

    def _score_spec(self):  # GMF.py:51-57: ranking on the logit (sigmoid is monotone; ties at saturation stated in DESIGN.md)
        if self.sharded:
            return self.score_kind, self.P.w, self._Qfull, self.h_gmf
        return self.score_kind, self.P.w, self.Q.w, self.h_gmf
