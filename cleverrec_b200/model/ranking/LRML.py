# coding: utf-8
" LRML: Latent Relational Metric Learning (2018) -- mirror of the reference model/ranking/LRML.py. "
from collections import defaultdict

import numpy as np
import torch

from .. import RankingRecommender as _rr
from ...engine import Table
from ...utils.metrics import batch_ranking_metrics


class LRML(_rr.RankingRecommender):
    def __init__(self, sess, data, configs, logger):
        super(LRML, self).__init__(sess, data, configs, logger)
        self.embed_size, self.reg, self.margin = int(configs['embed_size']), float(configs['reg']), float(configs['margin'])
        self.mem_size = int(configs['mem_size'])  # Number of memory slots
        logger.info(' model_params: embed_size=%d, mem_size=%d, reg=%s, margin=%s' % (self.embed_size, self.mem_size, self.reg, self.margin) +
                    ', ' + self.model_params)
        if self.loss_func != 'hinge':
            raise ValueError('LRML is defined with loss_func=hinge (conf/LRML.properties), got %r' % self.loss_func)

    def _create_params(self, init=None):
        """LRML.py:24-33.  P / Q are reached only through gathers (IndexedSlices): under tf.train.AdamOptimizer every row's
        moments decay and every row moves each step -- a dense apply with zero gradient on untouched rows, which for
        SGD / Adagrad is a no-op.  K [embed_size, mem_size] and M [mem_size, embed_size] are packed into one dense vector."""
        dev, kind = self.engine.device, self.optimizer.kind
        d, m = self.embed_size, self.mem_size
        shapes = {'P': [self.data.user_nums, d], 'Q': [self.data.item_nums, d], 'K': [d, m], 'M': [m, d]}
        vals = {}
        for name in ('P', 'Q', 'K', 'M'):   # creation order of LRML.py:29-33 (consumes the initializer's generator in that order)
            vals[name] = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer(shapes[name])
        for name in ('P', 'Q'):
            setattr(self, name, Table(vals[name].to(dev).contiguous(), kind, 'lazy'))
        self.dense = torch.cat([vals['K'].reshape(-1), vals['M'].reshape(-1)]).to(dev).contiguous()
        self.dense_s1 = torch.full_like(self.dense, 0.1) if kind == 'Adagrad' else (torch.zeros_like(self.dense) if kind == 'Adam' else None)
        self.dense_s2 = torch.zeros_like(self.dense) if kind == 'Adam' else None

    def build_model(self, init=None):
        self._create_params(init)

    @property
    def K(self):
        return self.dense[:self.embed_size * self.mem_size].reshape(self.embed_size, self.mem_size)

    @property
    def M(self):
        return self.dense[self.embed_size * self.mem_size:].reshape(self.mem_size, self.embed_size)

    def _variables(self):   # LRML.py:79-84
        return {'lrml_params/P': self.P.w, 'lrml_params/Q': self.Q.w, 'lrml_params/K': self.K, 'lrml_params/M': self.M}

    def train_step(self, u_idx, i_idx, j_idx, loss_out=None):
        """sess.run([train, loss], {u_idx, i_idx, j_idx})  (LRML.py:53-64).  No clipping: _unit_clipping (:66-69) is never called by
        build_model (:86-92) and would only rebind Python attributes anyway."""
        return self.engine.train_step_lrml(self.P, self.Q, self.dense, self.dense_s1, self.dense_s2, self.mem_size, self.optimizer,
                                           u_idx, i_idx, j_idx, self.margin, self.reg, loss_out=loss_out)

    def _train_epoch_pairwise(self, epoch, n_rows, n_batches, losses):
        for k in range(n_batches):
            lo = k * self.batch_size
            u, i, j = self.engine.sample_pairwise(self.seed, epoch, lo, min(self.batch_size, n_rows - lo), self.neg_ratio)
            self.train_step(u, i, j, loss_out=losses[k:k + 1])

    # ---- evaluation (LRML.py:70-78): relation-translated distances, ascending (cml_like, RankingRecommender.py:222,285) ----
    def _dist(self, u, i):
        return self.engine.score_pairs_lrml(self.P.w, self.Q.w, self.dense, self.mem_size, u, i)

    def test_model_loo(self):
        HR, MRR, NDCG = defaultdict(list), defaultdict(list), defaultdict(list)
        offsets, u_dev, i_dev, i_host = self._loo_feed()
        K = self.topk[-1]
        scores = self._dist(u_dev, i_dev)
        args = self.engine.topk_segments(scores, offsets, K, True).cpu().numpy()
        real_lists, rec = [], np.full((len(self.test_users), K), -1, dtype=np.int64)
        for k, u in enumerate(self.test_users):
            real_lists.append(self.data.ui_test[u][self.neg_samples:])
            valid = args[k] >= 0
            rec[k, valid] = i_host[offsets[k] + args[k][valid]]
        for kid in range(len(self.topk)):
            hr, mrr, ndcg = batch_ranking_metrics(real_lists, rec, self.topk[kid])
            HR[kid].extend(hr.tolist()); MRR[kid].extend(mrr.tolist()); NDCG[kid].extend(ndcg.tolist())
        return HR, MRR, NDCG

    def test_model_rs(self):
        HR, MRR, NDCG = defaultdict(list), defaultdict(list), defaultdict(list)
        K, I, dev = self.topk[-1], self.data.item_nums, self.engine.device
        items = torch.arange(I, dtype=torch.int32, device=dev)
        bt = max(1, min(self.batch_size_t, (1 << 26) // max(1, I)))
        for a in range(0, len(self.test_users), bt):
            cur = self.test_users[a:a + bt]
            users = torch.as_tensor(np.asarray(cur), dtype=torch.int32, device=dev)
            scores = self._dist(users.repeat_interleave(I), items.repeat(len(cur)))
            scores = self.engine.mask_seen(scores.reshape(len(cur), I), users, float('inf'))   # ascending: a seen item is infinitely far
            seg = torch.arange(len(cur) + 1, dtype=torch.int64, device=dev) * I
            topk_items = self.engine.topk_segments(scores.reshape(-1), seg, K, True).cpu().numpy()
            real_lists = [self.data.ui_test[u] for u in cur]
            for kid in range(len(self.topk)):
                hr, mrr, ndcg = batch_ranking_metrics(real_lists, topk_items, self.topk[kid])
                HR[kid].extend(hr.tolist()); MRR[kid].extend(mrr.tolist()); NDCG[kid].extend(ndcg.tolist())
        return HR, MRR, NDCG
