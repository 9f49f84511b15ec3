# coding: utf-8
" SBPR: Social Bayesian Personalized Ranking (2014) -- mirror of the reference model/ranking/SBPR.py. "
import math

import numpy as np
import torch

from .. import RankingRecommender as _rr
from ... import _lib
from ...engine import Table
from ...utils.tools import get_SPu


class SBPR(_rr.RankingRecommender):
    def __init__(self, sess, data, configs, logger):
        super(SBPR, self).__init__(sess, data, configs, logger)
        self.embed_size, self.reg = int(configs['embed_size']), float(configs['reg'])
        logger.info(' model_params: embed_size=%d, reg=%s' % (self.embed_size, self.reg) + ', ' + self.model_params)
        if self.loss_func != 'bpr':
            raise ValueError('SBPR is defined with loss_func=bpr (conf/SBPR.properties), got %r' % self.loss_func)
        # Get SPu (SBPR.py:16) and hand the sampler's view of the social data to the device
        self.SPu = get_SPu(data)
        if not self.SPu:
            raise ValueError('SBPR: no user has social items (data.user_friends / social_file) -- the reference would train on nothing')
        self.engine.set_social(data.ui_train, data.user_friends, self.SPu, data.user_nums)
        # Specify training model (SBPR.py:18)
        self.train_model = self.train_model_sbpr

    def _create_params(self, init=None):
        """SBPR.py:29-36: P, Q by the initializer, bias = zeros(item_nums + 1) (padded to a multiple of 4 for the dense apply)."""
        dev, kind = self.engine.device, self.optimizer.kind
        shapes = {'P': [self.data.user_nums, self.embed_size], 'Q': [self.data.item_nums, self.embed_size]}
        for name in ('P', 'Q'):
            w = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer(shapes[name])
            setattr(self, name, Table(w.to(dev).contiguous(), kind, 'lazy'))
        n = self.data.item_nums + 1
        b = torch.as_tensor(np.asarray(init['bias']), dtype=torch.float32) if init and 'bias' in init else torch.zeros(n)
        self.B = Table(torch.cat([b, torch.zeros((-n) % 4)]).reshape(-1, 1).to(dev).contiguous(), kind, 'lazy')
        self.n_bias = n

    def build_model(self, init=None):
        self._create_params(init)

    @property
    def bias(self):
        return self.B.w.reshape(-1)[:self.n_bias]

    def _variables(self):   # SBPR.py:65-70
        return {'sbpr_params/P': self.P.w, 'sbpr_params/Q': self.Q.w, 'sbpr_params/bias': self.bias}

    def train_step(self, u_idx, i_idx, i_s_idx, i_neg_idx, suk, loss_out=None):
        """sess.run([train, loss], {u_idx, i_idx, i_s_idx, i_neg_idx, suk})  (RankingRecommender.py:108-116, SBPR.py:51-57)."""
        return self.engine.train_step_sbpr(self.P, self.Q, self.B, self.optimizer, u_idx, i_idx, i_s_idx, i_neg_idx, suk, self.reg,
                                           loss_out=loss_out)

    # For SBPR (RankingRecommender.py:103-117).  is_suk=False would leave the `suk` placeholder unfed in the reference (the graph
    # always divides by it, SBPR.py:54) and fail; it is kept as "coefficient 1".
    def train_model_sbpr(self, is_suk=True):
        if self.sampler_mode == 'numpy_stream':
            # the epoch ranking_sampler_sbpr (utils/sampler.py:102-141) returns under NumPy's CURRENT global stream, bit for bit
            eng = self.engine
            eng.np_set_state()
            feeds = eng.sample_epoch_numpy_sbpr(self.neg_ratio, is_suk=is_suk)
            np.random.set_state(eng.np_get_state())
            n_rows = feeds[0].numel()
            n_batches = math.ceil(n_rows / self.batch_size)
            losses = torch.zeros(n_batches, dtype=torch.float64, device=eng.device)
            for b in range(n_batches):
                sl = slice(b * self.batch_size, min((b + 1) * self.batch_size, n_rows))
                suk = feeds[4][sl] if is_suk else torch.ones(sl.stop - sl.start, dtype=torch.float32, device=eng.device)
                self.train_step(feeds[0][sl], feeds[1][sl], feeds[2][sl], feeds[3][sl], suk, loss_out=losses[b:b + 1])
            self.epoch += 1
            return float(losses.sum().item()) / n_batches
        n_rows = self.engine.epoch_rows(self.neg_ratio, 'sbpr')
        n_batches = math.ceil(n_rows / self.batch_size)
        losses = torch.zeros(n_batches, dtype=torch.float64, device=self.engine.device)
        for b in range(n_batches):
            lo = b * self.batch_size
            cnt = min(self.batch_size, n_rows - lo)
            feeds = self.engine.sample_sbpr(self.seed, self.epoch, lo, cnt, self.neg_ratio, is_suk=is_suk)
            suk = feeds[4] if is_suk else torch.ones(cnt, dtype=torch.float32, device=self.engine.device)
            self.train_step(feeds[0], feeds[1], feeds[2], feeds[3], suk, loss_out=losses[b:b + 1])
        self.epoch += 1
        return float(losses.sum().item()) / n_batches

    def _score_spec(self):
        # SBPR.py:59-63: fed pairs -> p.q + bias; the full-rank branch is the plain matmul WITHOUT the bias
        if self.configs['data.split_way'] == 'loo' or self.neg_samples > 0:
            return _lib.SCORE_DOT_BIAS, self.P.w, self.Q.w, self.bias.contiguous()
        return _lib.SCORE_DOT, self.P.w, self.Q.w, None
