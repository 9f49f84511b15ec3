# coding: utf-8
" CML: Collaborative Metric Learning (2017) -- mirror of the reference model/ranking/CML.py. "
import math

import numpy as np
import torch

from .. import RankingRecommender as _rr
from ... import _lib
from ...engine import Table


class CML(_rr.RankingRecommender):
    def __init__(self, sess, data, configs, logger):
        super(CML, self).__init__(sess, data, configs, logger)
        self.embed_size, self.reg, self.margin = int(configs['embed_size']), float(configs['reg']), float(configs['margin'])
        logger.info(' model_params: embed_size=%d, reg=%s, margin=%s' % (self.embed_size, self.reg, self.margin) + ', ' + self.model_params)
        # Specify training/testing model (CML.py:16)
        self.train_model = self.train_model_cml
        # `clip_rows=True` would clip the tables after every step; the reference never does (its _unit_clipping rebinds Python
        # attributes to clipped temporaries), so parity mode leaves it off.
        self.clip_tables = configs.get('clip_rows', 'False') == 'True'

    def _create_params(self, init=None):
        dev = self.engine.device
        shapes = {'P': [self.data.user_nums, self.embed_size], 'Q': [self.data.item_nums, self.embed_size]}
        for name in ('P', 'Q'):
            w = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer(shapes[name])
            # both gradients are dense (covariance term): TF's dense apply -> plain Adam slots, no `last`
            setattr(self, name, Table(w.to(dev).contiguous(), self.optimizer.kind, 'lazy'))

    def build_model(self, init=None):
        self._create_params(init)

    def _variables(self):   # CML.py:86-91
        return {'cml_params/P': self.P.w, 'cml_params/Q': self.Q.w}

    def train_step(self, u_idx, i_idx, neg_items, loss_out=None):
        """sess.run([train, loss], {u_idx, i_idx, neg_items})  (CML.py:39-61)."""
        out = self.engine.train_step_cml(self.P, self.Q, self.optimizer, u_idx, i_idx, neg_items, self.margin, self.reg,
                                         self.data.item_nums, loss_out=loss_out)
        if self.clip_tables:
            self.P.w.copy_(self.engine.clip_rows(self.P.w))
            self.Q.w.copy_(self.engine.clip_rows(self.Q.w))
        return out

    # For CML (RankingRecommender.py:90-100)
    def train_model_cml(self):
        if self.sampler_mode == 'numpy_stream':
            # the epoch ranking_sampler_cml (utils/sampler.py:77-99) returns under NumPy's CURRENT global stream, bit for bit
            import numpy as np_
            eng = self.engine
            eng.np_set_state()
            u, i, neg = eng.sample_epoch_numpy('cml', self.neg_ratio)
            np_.random.set_state(eng.np_get_state())
            n_rows = u.numel()
            n_batches = math.ceil(n_rows / self.batch_size)
            losses = torch.zeros(n_batches, dtype=torch.float64, device=eng.device)
            for k in range(n_batches):
                sl = slice(k * self.batch_size, min((k + 1) * self.batch_size, n_rows))
                self.train_step(u[sl], i[sl], neg[sl], loss_out=losses[k:k + 1])
            self.epoch += 1
            return float(losses.sum().item()) / n_batches
        n_rows = self.engine.epoch_rows(self.neg_ratio, 'cml')
        n_batches = math.ceil(n_rows / self.batch_size)
        losses = torch.zeros(n_batches, dtype=torch.float64, device=self.engine.device)
        for k in range(n_batches):
            lo = k * self.batch_size
            u, i, neg = self.engine.sample_cml(self.seed, self.epoch, lo, min(self.batch_size, n_rows - lo), self.neg_ratio)
            self.train_step(u, i, neg, loss_out=losses[k:k + 1])
        self.epoch += 1
        return float(losses.sum().item()) / n_batches

    def _score_spec(self):
        # CML.py:80-84: loo -> squared distance of the fed pairs; full rank -> the *clipped* user rows against unclipped Q
        P = self.P.w
        if not (self.configs['data.split_way'] == 'loo' or self.neg_samples > 0):
            P = self.engine.clip_rows(P)
        return _lib.SCORE_SQDIST, P, self.Q.w, None
