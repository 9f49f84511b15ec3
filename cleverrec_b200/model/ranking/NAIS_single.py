# coding: utf-8
" NAIS: Neural Attentive Item Similarity Model (2018) -- mirror of the reference model/ranking/NAIS_single.py (one user per step). "
from collections import defaultdict

import numpy as np
import torch

from .. import RankingRecommender as _rr
from ...engine import Table
from ...utils.metrics import cal_ranking_metrics


class NAIS_single(_rr.RankingRecommender):
    def __init__(self, sess, data, configs, logger):
        super(NAIS_single, self).__init__(sess, data, configs, logger)
        self.embed_size, self.atten_size, self.reg = int(configs['embed_size']), int(configs['atten_size']), float(configs['reg'])
        self.beta = float(configs['beta'])  # The smoothing coefficient of Softmax
        self.atten_type = configs['atten_type']  # concat/prod -- compared with == 'concat' as in NAIS_single.py:52,67: the shipped conf's
        # quoted value atten_type='prod' (and anything else) takes the product branch
        self.concat = self.atten_type == 'concat'
        logger.info(' model_params: embed_size=%d, atten_size=%d, atten_type=%s, reg=%s, beta=%s' % (self.embed_size, self.atten_size,
                    self.atten_type, self.reg, self.beta) + ', ' + self.model_params)
        # Specify training and testing model (NAIS_single.py:20-21)
        self.train_model = self.train_model_nais
        self.test_model_rs, self.test_model_loo = self.test_model_rs_nais, self.test_model_loo_nais
        users = list(data.ui_train.keys())
        lens = np.asarray([len(data.ui_train[u]) for u in users], dtype=np.int64)
        self._list_start = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        self._list_len = lens.astype(np.int32)
        self._train_users = users

    def _create_params(self, init=None):
        """NAIS_single.py:40-57: P, Q [(I+1), d], bias/b/h ~ U(-0.1, 0.1), W [d, atten_size]."""
        dev, kind, n = self.engine.device, self.optimizer.kind, self.data.item_nums + 1
        g = self.init_generator

        def get(name, fn):
            return torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else fn()
        unif = lambda *s: torch.rand(*s, generator=g) * 0.2 - 0.1
        self.P = Table(get('P', lambda: self.initializer([n, self.embed_size])).to(dev).contiguous(), kind, 'lazy')
        self.Q = Table(get('Q', lambda: self.initializer([n, self.embed_size])).to(dev).contiguous(), kind, 'lazy')
        b = get('bias', lambda: unif(n))
        self.B = Table(torch.cat([b, torch.zeros((-n) % 4)]).reshape(-1, 1).to(dev).contiguous(), kind, 'lazy')
        self.n_bias = n
        W = get('W', lambda: self.initializer([(2 if self.concat else 1) * self.embed_size, self.atten_size]))   # NAIS_single.py:52-55
        self.dense = torch.cat([W.reshape(-1), get('b', lambda: unif(self.atten_size)), get('h', lambda: unif(self.atten_size))]).to(dev)
        self.dense_s1 = torch.full_like(self.dense, 0.1) if kind == 'Adagrad' else (torch.zeros_like(self.dense) if kind == 'Adam' else None)
        self.dense_s2 = torch.zeros_like(self.dense) if kind == 'Adam' else None

    def build_model(self, init=None):
        self._create_params(self._with_pretrained(init))

    def _with_pretrained(self, init):
        """NAIS_single.py:35-38 `_load_fism_params`: P, Q and the item bias start from a trained FISM (config key fism_pretrain =
        the FISM checkpoint directory), as the NAIS paper prescribes.  The reference defines the restore but never calls it and
        its shipped conf has no such key; here it runs when the key is present and a checkpoint exists.  Explicit `init` wins."""
        from ...utils.tools import latest_checkpoint, load_checkpoint
        ckpt = latest_checkpoint(self.configs.get('fism_pretrain'))
        if ckpt is None:
            if 'fism_pretrain' in self.configs:
                self.logger.info(' fism_pretrain=%s holds no checkpoint: training from scratch' % self.configs['fism_pretrain'])
            return init
        v = load_checkpoint(ckpt)
        out = {'P': v['FISM_paras/P'], 'Q': v['FISM_params/Q'], 'bias': v['FISM_params/b']}
        out.update(init or {})
        self.logger.info(' restored P, Q, bias from %s' % ckpt)
        return out

    def _variables(self):   # NAIS_single.py:99-106
        d, a = (2 if self.concat else 1) * self.embed_size, self.atten_size
        return {'NAIS_paras/P': self.P.w, 'NAIS_params/Q': self.Q.w, 'NAIS_params/bias': self.bias, 'NAIS_params/W': self.dense[:d * a].reshape(d, a),
                'NAIS_params/b': self.dense[d * a:d * a + a], 'NAIS_params/h': self.dense[d * a + a:d * a + 2 * a]}

    @property
    def bias(self):
        return self.B.w.reshape(-1)[:self.n_bias].contiguous()

    def train_step(self, u_idx, i_idx, y, loss_out=None):
        """sess.run([train, loss], {u_idx: history, u_nbrs_num, i_idx, i_nums, y})  (NAIS_single.py:82-90)."""
        return self.engine.train_step_nais(self.P, self.Q, self.B, self.dense, self.dense_s1, self.dense_s2, self.atten_size, self.optimizer,
                                           u_idx, i_idx, y, self.beta, self.reg, loss_out=loss_out, concat=self.concat)

    # Form mini-batch by user (RankingRecommender.py:64-87): one optimizer step per user, in data.ui_train order
    def train_model_nais(self):
        if self.sampler_mode == 'numpy_stream':
            # the negatives RankingRecommender.py:73-80 draws under NumPy's CURRENT global stream (per positive: neg_ratio distinct unseen
            # items, users and items in data.ui_train order, no permutation), bit for bit; then one step per user as in the reference
            eng, R = self.engine, self.neg_ratio
            eng.np_set_state()
            negs = eng.sample_epoch_numpy('negatives', R)                      # [n_pos, R]
            np.random.set_state(eng.np_get_state())
            pos_item = eng._hist[1]
            targets_all = torch.cat([pos_item.reshape(-1, 1), negs], dim=1)    # per positive: the item, then its negatives (:70-80)
            y_row = torch.zeros(1 + R, dtype=torch.float32, device=eng.device)
            y_row[0] = 1.0
            losses = torch.zeros(len(self._train_users), dtype=torch.float64, device=eng.device)
            for k in range(len(self._train_users)):
                a, n = int(self._list_start[k]), int(self._list_len[k])
                self.train_step(pos_item[a:a + n], targets_all[a:a + n].reshape(-1), y_row.repeat(n), loss_out=losses[k:k + 1])
            self.epoch += 1
            return float(losses.sum().item()) / len(self._train_users)
        losses = torch.zeros(len(self._train_users), dtype=torch.float64, device=self.engine.device)
        self.engine.train_epoch_nais(self.P, self.Q, self.B, self.dense, self.dense_s1, self.dense_s2, self.atten_size, self.optimizer, self.seed,
                                     self.epoch, self._list_start, self._list_len, self.neg_ratio, self.beta, self.reg, losses, concat=self.concat)
        self.epoch += 1
        return float(losses.sum().item()) / len(self._train_users)

    def _scores(self, u, targets):
        hist = self.data.ui_train[u] if u in self.data.ui_train else [self.data.item_nums]
        return self.engine.score_nais(self.P.w, self.Q.w, self.bias, self.dense, self.atten_size, hist, targets, self.beta, concat=self.concat).cpu().numpy()

    def _device_history(self, u):
        """The user's interaction list as a view of the engine's device copy of data.ui_train (no host -> device copy per user)."""
        k = self._train_pos.get(u)
        if k is None:
            return self._no_history
        a = int(self._list_start[k])
        return self.engine._hist[1][a:a + int(self._list_len[k])]

    def _eval_setup(self):
        if getattr(self, '_train_pos', None) is None:
            self._train_pos = {u: k for k, u in enumerate(self._train_users)}
            self._no_history = torch.tensor([self.data.item_nums], dtype=torch.int32, device=self.engine.device)
        return self.bias   # one contiguous copy of the (padded) bias vector per evaluation

    def test_model_loo_nais(self):  # RankingRecommender.py:330-348
        """One scoring call per user like the reference's one sess.run per user, but every call reads the history and the candidates
        from device memory and writes into one score buffer; the host sees the scores once, then ranks per user as the reference does."""
        HR, MRR, NDCG = defaultdict(list), defaultdict(list), defaultdict(list)
        bias = self._eval_setup()
        offsets, u_dev, i_dev, i_host = self._loo_feed()
        scores_dev = torch.empty(int(offsets[-1]), dtype=torch.float32, device=self.engine.device)
        for k, u in enumerate(self.test_users):
            a, b = int(offsets[k]), int(offsets[k + 1])
            self.engine.score_nais(self.P.w, self.Q.w, bias, self.dense, self.atten_size, self._device_history(u), i_dev[a:b], self.beta,
                                   out=scores_dev[a:b], concat=self.concat)
        scores = scores_dev.cpu().numpy()
        for k, u in enumerate(self.test_users):
            pre_scores = scores[offsets[k]:offsets[k + 1]]
            args_u = np.argsort(-pre_scores, kind='stable')[:self.topk[-1]]
            real_items = self.data.ui_test[u][self.neg_samples:]
            for kid in range(len(self.topk)):
                rec_items = np.take(self.data.ui_test[u], args_u[:self.topk[kid]])
                hr_u, mrr_u, ndcg_u = cal_ranking_metrics(real_items, rec_items, self.topk[kid])
                HR[kid].append(hr_u); MRR[kid].append(mrr_u); NDCG[kid].append(ndcg_u)
        return HR, MRR, NDCG

    def test_model_rs_nais(self):  # RankingRecommender.py:301-328
        """All-item scores per user (id = item_nums, which the reference scores and then drops, is not scored), seen items masked,
        first topk[-1] by (score descending, id ascending) -- np.argsort(-scores, kind='stable') followed by the reference's skip loop --
        taken on the device for a block of users at a time."""
        from ...utils.metrics import batch_ranking_metrics
        HR, MRR, NDCG = defaultdict(list), defaultdict(list), defaultdict(list)
        bias = self._eval_setup()
        K, I, dev = self.topk[-1], self.data.item_nums, self.engine.device
        all_items = torch.arange(I, dtype=torch.int32, device=dev)
        block = max(1, min(self.batch_size_t, (1 << 26) // max(1, I)))
        for a in range(0, len(self.test_users), block):
            cur = self.test_users[a:a + block]
            scores = torch.empty((len(cur), I), dtype=torch.float32, device=dev)
            for k, u in enumerate(cur):
                self.engine.score_nais(self.P.w, self.Q.w, bias, self.dense, self.atten_size, self._device_history(u), all_items, self.beta,
                                       out=scores[k], concat=self.concat)
            users = torch.as_tensor(np.asarray(cur), dtype=torch.int32, device=dev)
            scores = self.engine.mask_seen(scores, users)
            seg = torch.arange(len(cur) + 1, dtype=torch.int64, device=dev) * I
            topk_items = self.engine.topk_segments(scores.reshape(-1), seg, K).cpu().numpy()
            real_lists = [self.data.ui_test[u] for u in cur]
            for kid in range(len(self.topk)):
                hr, mrr, ndcg = batch_ranking_metrics(real_lists, topk_items, self.topk[kid])
                HR[kid].extend(hr.tolist()); MRR[kid].extend(mrr.tolist()); NDCG[kid].extend(ndcg.tolist())
        return HR, MRR, NDCG
