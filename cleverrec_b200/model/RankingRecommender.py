" Recommender for Ranking Model -- mirror of the reference model/RankingRecommender.py (hot-path parts). "
# Same public surface as the reference class:
#   train_model() -> float                       mean of the per-step *summed* losses   (:33-61)
#   test_model_loo() / test_model_rs()           -> (HR, MRR, NDCG) defaultdict(list)   (:250-299 / :198-247)
#   run_model()                                   epoch loop, log format, best-by-NDCG@topk[0] (:395-440)
# What changed underneath: the sampler, every sess.run and the argsort/top-K run in libcleverrec_b200.so.
import math
import time
from collections import defaultdict

import numpy as np
import torch

from .Recommender import Recommender
from ..utils.metrics import batch_ranking_metrics
from ..utils.tools import timer


class RankingRecommender(Recommender):
    def __init__(self, sess, data, configs, logger):
        super(RankingRecommender, self).__init__(sess, data, configs, logger)
        self.neg_ratio = int(configs['neg_ratio'])
        self.model_params += ', neg_ratio=%d' % self.neg_ratio
        # Testing data
        self.test_users = list(data.ui_test.keys())
        self.test_batches = math.ceil(len(self.test_users) / self.batch_size_t)
        self.epoch = 0            # epochs sampled so far (the sampler's `epoch` counter)
        if self.world > 1 and not self.supports_sharding:
            raise NotImplementedError('%s under WORLD_SIZE=%d: the multi-GPU path (users partitioned, item table row-sharded over '
                                      'NVLink peer memory) is implemented for BPR' % (self.model, self.world))
        self._install_history()
        self._test_cache = None

    supports_sharding = False

    def _install_history(self):
        data = self.data
        if getattr(data, 'train_rows', None) is not None:
            # the packaged RankingPreprocess also keeps the training split as two int32 columns in the dict's enumeration
            # order: the device history is built from them natively (csrc/history.cu) instead of walking the dict of lists
            self.engine.build_history(data.train_rows[0], data.train_rows[1], self._history_rows(), data.item_nums)
        else:
            self.engine.set_history(data.ui_train, self._history_rows(), data.item_nums)

    def _history_rows(self):
        return self.data.user_nums

    # ---- what a model provides -------------------------------------------------------------------------------
    def _train_epoch_pairwise(self, epoch, n_rows, n_batches, losses):
        """Run all steps of one pairwise epoch; losses: double device tensor [n_batches]."""
        raise NotImplementedError

    def _train_epoch_pointwise(self, epoch, n_rows, n_batches, losses):
        raise NotImplementedError

    def _score_spec(self):
        """-> (kind, P_tensor, Q_tensor, hvec_or_None) for crb_score_pairs / crb_score_topk."""
        raise NotImplementedError

    def _before_eval(self):
        pass

    # ---- numpy_stream mode: the epoch the reference's sampler would return for NumPy's CURRENT global state (np.random.seed(s)
    # before run_model reproduces the reference's triplet sequence bit for bit), fed batch by batch like RankingRecommender.py:39-46
    def _train_epoch_numpy_stream(self):
        import numpy as np_
        eng = self.engine
        eng.np_set_state()
        if self.is_pairwise == 'True':
            feeds = eng.sample_epoch_numpy('pairwise', self.neg_ratio, with_nbr=self.fism_like)
        else:
            feeds = eng.sample_epoch_numpy('pointwise', self.neg_ratio)
        np_.random.set_state(eng.np_get_state())
        n_rows = feeds[0].numel()
        n_batches = math.ceil(n_rows / self.batch_size)
        losses = torch.zeros(n_batches, dtype=torch.float64, device=eng.device)
        for k in range(n_batches):
            sl = slice(k * self.batch_size, min((k + 1) * self.batch_size, n_rows))
            self.train_step(*(f[sl] for f in feeds), loss_out=losses[k:k + 1])
        self.epoch += 1
        return float(losses.sum().item()) / n_batches

    # ---- Train the model (Single epoch) ----------------------------------------------------------------------
    def train_model(self):
        if self.sampler_mode == 'numpy_stream':
            return self._train_epoch_numpy_stream()
        if self.is_pairwise == 'True':
            n_rows = self.engine.epoch_rows(self.neg_ratio, 'pairwise')
            n_batches = math.ceil(n_rows / self.batch_size)
            losses = torch.zeros(n_batches, dtype=torch.float64, device=self.engine.device)
            self._train_epoch_pairwise(self.epoch, n_rows, n_batches, losses)
        else:
            n_rows = self.engine.epoch_rows(self.neg_ratio, 'pointwise')
            n_batches = math.ceil(n_rows / self.batch_size)
            losses = torch.zeros(n_batches, dtype=torch.float64, device=self.engine.device)
            self._train_epoch_pointwise(self.epoch, n_rows, n_batches, losses)
        self.epoch += 1
        total_loss = float(losses.sum().item())  # one device->host read per epoch
        bad = self.engine.sampler_errors()       # the stream is already drained by the read above
        if bad:
            raise RuntimeError('sampler: %d rows found no admissible negative (a user has fewer than neg_ratio unseen items; '
                               'the reference would loop forever, utils/sampler.py:58-61)' % bad)
        return total_loss / n_batches

    # ---- Leave-One-Out / Random split with 1000 negative items ---------------------------------------------
    def _loo_feed(self):
        """Flattened (u_idx, i_idx) of all test users in test_users order (RankingRecommender.py:257-264), built once."""
        if self._test_cache is None:
            ui_test = self.data.ui_test
            lens = np.fromiter((len(ui_test[u]) for u in self.test_users), dtype=np.int64, count=len(self.test_users))
            offsets = np.zeros(len(self.test_users) + 1, dtype=np.int64)
            np.cumsum(lens, out=offsets[1:])
            u_idx = np.repeat(np.asarray(self.test_users, dtype=np.int32), lens)
            i_idx = np.fromiter((i for u in self.test_users for i in ui_test[u]), dtype=np.int32, count=int(offsets[-1]))
            dev = self.engine.device
            self._test_cache = (offsets, torch.from_numpy(u_idx).to(dev), torch.from_numpy(i_idx).to(dev), i_idx)
        return self._test_cache

    def _pair_users(self, u_idx):
        """user-side row of each flattened pair (FISM-like models override: rows of the precomputed user matrix)."""
        return u_idx

    def _loo_lists(self):
        """Per test user: the real items (ui_test[u][neg_samples:], RankingRecommender.py:283-285), built once."""
        if getattr(self, '_real_cache', None) is None or self._real_cache[0] is not self.test_users:
            real_lists = []
            for u in self.test_users:
                real_items = self.data.ui_test[u][self.neg_samples:]
                if not isinstance(real_items, list):  # loo
                    real_items = [real_items]
                real_lists.append(real_items)
            dev = self.engine.device
            self._real_cache = (self.test_users, real_lists, torch.as_tensor(np.asarray(self.test_users, dtype=np.int32), device=dev))
        return self._real_cache[1], self._real_cache[2]

    def test_model_loo(self):
        """The reference walks the test users in batches of test.batch_size (one sess.run + a Python loop per batch, :255-298); every
        user's result is independent of the batching, so here ALL test users go through one fused score + top-K call
        (crb_score_pairs_topk) and one vectorised metric pass; the returned lists are in self.test_users order as in the reference."""
        HR, MRR, NDCG = defaultdict(list), defaultdict(list), defaultdict(list)  # evaluation metrics
        self._before_eval()
        offsets, u_dev, i_dev, i_host = self._loo_feed()
        real_lists, users_dev = self._loo_lists()
        kind, P, Q, hvec = self._score_spec()
        K = self.topk[-1]
        n = len(self.test_users)
        if n == 0:
            return HR, MRR, NDCG
        # Predict + evaluate: args_u = np.argsort(-pre_scores_u)[:topk[-1]]  (ascending for cml_like)
        args = self.engine.score_pairs_topk(kind, P, Q, self._pair_users(users_dev), i_dev, offsets, K, hvec=hvec, ascending=self.cml_like)
        args = args.cpu().numpy().astype(np.int64)
        valid = args >= 0
        idx = np.minimum(offsets[:-1, None] + np.where(valid, args, 0), max(int(offsets[-1]) - 1, 0))
        rec = np.where(valid, i_host[idx] if i_host.shape[0] else -1, -1)   # np.take(ui_test[u], args_u)
        for kid in range(len(self.topk)):
            hr, mrr, ndcg = batch_ranking_metrics(real_lists, rec, self.topk[kid])
            HR[kid].extend(hr.tolist())
            MRR[kid].extend(mrr.tolist())
            NDCG[kid].extend(ndcg.tolist())
        return HR, MRR, NDCG

    # ---- Random split with all ------------------------------------------------------------------------------
    def _fullrank_users(self, cur_users):
        """-> (user rows into P, history user ids or None)"""
        return np.asarray(cur_users, dtype=np.int32), None

    def test_model_rs(self):
        """One fused call for ALL test users (the reference's per-batch loop :201-246 gives each user the same result whatever
        test.batch_size is): the item table is converted for the tensor cores once per evaluation instead of once per batch."""
        HR, MRR, NDCG = defaultdict(list), defaultdict(list), defaultdict(list)  # evaluation metrics
        self._before_eval()
        kind, P, Q, hvec = self._score_spec()
        K = self.topk[-1]
        if not self.test_users:
            return HR, MRR, NDCG
        rows, hist = self._fullrank_users(self.test_users)
        # pre_scores = sess.run(...); argsort; skip ui_train[u]; first topk[-1]  -> one fused call
        topk_items = self.engine.score_topk(kind, P, Q, rows, K, hvec=hvec, hist_users=hist, exact=self.score_exact,
                                            n_items=self.data.item_nums)
        real_lists = [self.data.ui_test[u] for u in self.test_users]
        for kid in range(len(self.topk)):
            hr, mrr, ndcg = batch_ranking_metrics(real_lists, topk_items, self.topk[kid])
            HR[kid].extend(hr.tolist())
            MRR[kid].extend(mrr.tolist())
            NDCG[kid].extend(ndcg.tolist())
        return HR, MRR, NDCG

    @timer('run_model')
    def run_model(self):
        self.build_model()  # allocate + initialise tables (the reference builds the graph and runs the initializer)

        best_ndcg10, best_epoch = 0, 0  # Record best metrics w.r.t. NDCG@10
        best_metrics = {}
        for epoch in range(self.epoches):
            # Train
            t1 = time.time()
            avg_loss = self.train_model()
            self.logger.info(' epoch %d\n  Training loss: %.4f, time: %s' % (epoch + 1, avg_loss, time.strftime('%H:%M:%S', time.gmtime(time.time() - t1))))

            # Test
            t2 = time.time()
            if (epoch + 1) % self.T:  # Test every T epoches
                continue
            if self.configs['data.split_way'] == 'loo' or self.neg_samples > 0:  # loo/random sampling with 1000 ...
                HR, MRR, NDCG = self.test_model_loo()
            else:  # random split
                HR, MRR, NDCG = self.test_model_rs()
            self.logger.info('  Testing time: %s' % time.strftime('%H:%M:%S', time.gmtime(time.time() - t2)))

            # Record best performance
            best_flag = False
            for id in range(len(self.topk)):
                hr, mrr, ndcg = np.mean(HR[id]), np.mean(MRR[id]), np.mean(NDCG[id])
                self.logger.info('  (k=%d) HR=%.4f, MRR=%.4f, NDCG=%.4f' % (self.topk[id], hr, mrr, ndcg))
                if id == 0 and ndcg > best_ndcg10:
                    best_flag = True
                    best_ndcg10 = ndcg
                if best_flag:
                    best_metrics[id] = (hr, mrr, ndcg)
                    best_epoch = epoch + 1
                    if id == len(self.topk) - 1 and self.configs.get('save_model', 'False') == 'True':
                        self.save_model()   # the reference's commented-out saver.save (:432-433), opt-in

        # Final results
        self.logger.info('best_epoch: %d' % best_epoch)
        for id in range(len(self.topk)):
            hr, mrr, ndcg = best_metrics[id]
            self.logger.info('  (k=%d) HR=%.4f, MRR=%.4f, NDCG=%.4f' % (self.topk[id], hr, mrr, ndcg))
        return best_epoch, best_metrics
