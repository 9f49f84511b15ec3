# coding: utf-8
" TransCF: Translational Collaborative Filtering (2018) -- mirror of the reference model/ranking/TransCF.py. "
from collections import defaultdict

import numpy as np
import torch

from .. import RankingRecommender as _rr
from ...engine import Table
from ...utils.metrics import batch_ranking_metrics


class TransCF(_rr.RankingRecommender):
    def __init__(self, sess, data, configs, logger):
        super(TransCF, self).__init__(sess, data, configs, logger)
        self.embed_size, self.reg1, self.reg2, self.margin = int(configs['embed_size']), float(configs['reg1']), float(configs['reg2']), \
            float(configs['margin'])
        logger.info(' model_params: embed_size=%d, reg1=%s, reg2=%s, margin=%s' % (self.embed_size, self.reg1, self.reg2, self.margin) +
                    ', ' + self.model_params)
        if self.loss_func != 'hinge':
            raise ValueError('TransCF is defined with loss_func=hinge (conf/TransCF.properties), got %r' % self.loss_func)
        # ui_sp_mat / iu_sp_mat (TransCF.py:16, utils/tools.py:100-113) as two CSR-style lists on the device: the user-side one is
        # the history the base class installed, the item-side one is its transpose (duplicates kept, 1/count weights implied)
        self.engine.set_item_lists()

    def _create_params(self, init=None):
        """TransCF.py:25-31.  Both table gradients are dense in the reference (they flow through the two SpMMs of :41-42), so TF
        applies the dense optimizer: plain Adam slots, no `last`."""
        dev = self.engine.device
        shapes = {'P': [self.data.user_nums, self.embed_size], 'Q': [self.data.item_nums, self.embed_size]}
        for name in ('P', 'Q'):
            w = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer(shapes[name])
            setattr(self, name, Table(w.to(dev).contiguous(), self.optimizer.kind, 'lazy'))

    def build_model(self, init=None):
        self._create_params(init)

    def _variables(self):   # TransCF.py:87-92
        return {'transcf_params/P': self.P.w, 'transcf_params/Q': self.Q.w}

    def train_step(self, u_idx, i_idx, j_idx, loss_out=None):
        """sess.run([train, loss], {u_idx, i_idx, j_idx})  (TransCF.py:38-62).  No clipping: the reference's _unit_clipping
        rebinds Python attributes to clipped temporaries after the train op is built."""
        return self.engine.train_step_transcf(self.P, self.Q, self.optimizer, u_idx, i_idx, j_idx, self.margin, self.reg1, self.reg2,
                                              loss_out=loss_out)

    def _train_epoch_pairwise(self, epoch, n_rows, n_batches, losses):
        for k in range(n_batches):
            lo = k * self.batch_size
            u, i, j = self.engine.sample_pairwise(self.seed, epoch, lo, min(self.batch_size, n_rows - lo), self.neg_ratio)
            self.train_step(u, i, j, loss_out=losses[k:k + 1])

    # ---- evaluation (TransCF.py:79-85): distances, ascending (cml_like, RankingRecommender.py:222,285) ----
    def _neighbourhoods(self):
        A = self.engine.transcf_neighbourhood(0, self.Q.w, self.data.user_nums)   # all_u_nbr_embed  (:41)
        B = self.engine.transcf_neighbourhood(1, self.P.w, self.data.item_nums)   # all_i_nbr_embed  (:42)
        return A, B

    def test_model_loo(self):
        HR, MRR, NDCG = defaultdict(list), defaultdict(list), defaultdict(list)
        offsets, u_dev, i_dev, i_host = self._loo_feed()
        K = self.topk[-1]
        A, B = self._neighbourhoods()
        scores = self.engine.score_pairs_transcf(self.P.w, self.Q.w, A, B, u_dev, i_dev)
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
        A, B = self._neighbourhoods()
        # TransCF.py:59-62,83-85: _unit_clipping rebinds self.u_embed to clip_by_norm(u_embed, 1.0, axes=1) BEFORE _predict is built,
        # so the all-item branch scores the *clipped* user row; the neighbourhood means (A, B: from the variables) and Q are not
        # clipped, and the candidate-pair branch (pre_scores = ui_dist, built before the clipping) is not either.  Same quirk as CML's
        # full-rank branch; found by running the genuine class on the TF-1 shim (tests/test_reference_graphs.py).
        P_eval = self.engine.clip_rows(self.P.w)
        items = torch.arange(I, dtype=torch.int32, device=dev)
        bt = max(1, min(self.batch_size_t, (1 << 26) // max(1, I)))
        for a in range(0, len(self.test_users), bt):
            cur = self.test_users[a:a + bt]
            users = torch.as_tensor(np.asarray(cur), dtype=torch.int32, device=dev)
            scores = self.engine.score_pairs_transcf(P_eval, self.Q.w, A, B, users.repeat_interleave(I), items.repeat(len(cur)))
            scores = self.engine.mask_seen(scores.reshape(len(cur), I), users, float('inf'))   # ascending: a seen item is infinitely far
            seg = torch.arange(len(cur) + 1, dtype=torch.int64, device=dev) * I
            topk_items = self.engine.topk_segments(scores.reshape(-1), seg, K, True).cpu().numpy()
            real_lists = [self.data.ui_test[u] for u in cur]
            for kid in range(len(self.topk)):
                hr, mrr, ndcg = batch_ranking_metrics(real_lists, topk_items, self.topk[kid])
                HR[kid].extend(hr.tolist()); MRR[kid].extend(mrr.tolist()); NDCG[kid].extend(ndcg.tolist())
        return HR, MRR, NDCG
