# coding: utf-8
" FISM: Factorized Item Similarity Model (2013) -- mirror of the reference model/ranking/FISM.py (pairwise branch). "

import numpy as np
import torch

from .. import RankingRecommender as _rr
from ... import _lib
from ...engine import Table


class FISM(_rr.RankingRecommender):
    def __init__(self, sess, data, configs, logger):
        super(FISM, self).__init__(sess, data, configs, logger)
        self.embed_size, self.reg, self.reg_bias = int(configs['embed_size']), float(configs['reg']), float(configs['reg_bias'])
        self.alpha = float(configs['alpha'])
        if self.is_pairwise != 'True':
            # FISM.py:61 uses an undefined `y` and the pointwise sampler feeds mismatched lengths
            raise NotImplementedError('FISM pointwise is broken in the reference; use the shipped pairwise configuration')
        logger.info(' model_params: embed_size=%d, alpha=%s, reg=%s, reg_bias=%s' % (self.embed_size, self.alpha, self.reg, self.reg_bias) + ', ' + self.model_params)
        if self.loss_func != 'bpr':
            raise ValueError('FISM (pairwise) is defined with loss_func=bpr (conf/FISM.properties), got %r' % self.loss_func)

    def _create_params(self, init=None):
        """FISM.py:32-38: P, Q [(I+1), d] and b ~ U(-0.1, 0.1) [(I+1)] (padded to a multiple of 4 for the dense apply)."""
        dev = self.engine.device
        n = self.data.item_nums + 1
        for name in ('P', 'Q'):
            w = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer([n, self.embed_size])
            setattr(self, name, Table(w.to(dev).contiguous(), self.optimizer.kind, 'lazy'))
        b = torch.as_tensor(np.asarray(init['b']), dtype=torch.float32) if init and 'b' in init else \
            (torch.rand(n, generator=self.init_generator) * 0.2 - 0.1)
        pad = (-n) % 4
        bw = torch.cat([b, torch.zeros(pad)]).reshape(-1, 1).to(dev).contiguous()
        self.B = Table(bw, self.optimizer.kind, 'lazy')
        self.n_bias = n

    def build_model(self, init=None):
        self._create_params(init)

    def _variables(self):   # FISM.py:72-77 -- 'FISM_paras/P' is the reference's spelling; NAIS_single.py:36 restores by it
        return {'FISM_paras/P': self.P.w, 'FISM_params/Q': self.Q.w, 'FISM_params/b': self.b}

    @property
    def b(self):
        return self.B.w.reshape(-1)[:self.n_bias]

    def train_step(self, u_idx, i_idx, j_idx, u_neighbors_num, loss_out=None):
        """sess.run([train, loss], {u_idx, i_idx, j_idx, u_neighbors_num})  (FISM.py:49-63)."""
        return self.engine.train_step_fism(self.P, self.Q, self.B, self.optimizer, u_idx, i_idx, j_idx, u_neighbors_num, self.alpha, self.reg,
                                           self.reg_bias, self.batch_size, loss_out=loss_out)

    def _train_epoch_pairwise(self, epoch, n_rows, n_batches, losses):
        for k in range(n_batches):
            lo = k * self.batch_size
            u, i, j, nbr = self.engine.sample_pairwise(self.seed, epoch, lo, min(self.batch_size, n_rows - lo), self.neg_ratio, with_nbr=True)
            self.train_step(u, i, j, nbr, loss_out=losses[k:k + 1])

    # ---- evaluation: users are represented by coeff * mean(P[history]) (FISM.py:65-70); the test-time u_neighbors_num is
    # len(ui_train[u]) -- the list length, duplicates counted (RankingRecommender.py:208,263) ----
    def _user_matrix(self):
        users = np.asarray(self.test_users, dtype=np.int32)
        nbr = np.asarray([len(self.data.ui_train[u]) if u in self.data.ui_train else 0 for u in self.test_users], dtype=np.int32)
        self._test_row = {u: k for k, u in enumerate(self.test_users)}
        return self.engine.fism_user_vectors(self.P.w, users, nbr, self.alpha)

    def _before_eval(self):
        self._S = self._user_matrix()

    def _score_spec(self):
        return _lib.SCORE_DOT_BIAS, self._S, self.Q.w, self.b.contiguous()

    def _pair_users(self, u_idx):  # rows of the user matrix instead of user ids
        lut = torch.full((self.data.user_nums,), -1, dtype=torch.int32, device=self.engine.device)
        users = torch.as_tensor(np.asarray(self.test_users), dtype=torch.int64, device=self.engine.device)
        lut[users] = torch.arange(users.numel(), dtype=torch.int32, device=self.engine.device)
        return lut[u_idx.long()]

    def _fullrank_users(self, cur_users):
        rows = np.asarray([self._test_row[u] for u in cur_users], dtype=np.int32)
        return rows, np.asarray(cur_users, dtype=np.int32)
