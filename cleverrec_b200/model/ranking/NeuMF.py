# coding: utf-8
" Neural Collaborative Filtering (2017) -- mirror of the reference model/ranking/NeuMF.py (GMF + MLP tower fused by h_neumf). "
from collections import defaultdict

import numpy as np
import torch

from .. import RankingRecommender as _rr
from ... import _lib
from ...engine import Table
from ...utils.metrics import batch_ranking_metrics

_LOSS = {'cross_entropy': _lib.LOSS_CROSS_ENTROPY, 'square': _lib.LOSS_SQUARE}


class NeuMF(_rr.RankingRecommender):
    def __init__(self, sess, data, configs, logger):
        super(NeuMF, self).__init__(sess, data, configs, logger)
        self.embed_size = int(configs['embed_size'])
        self.layers = list(map(int, configs['layers'][1:-1].split(',')))
        for a, b in zip(self.layers[:-1], self.layers[1:]):
            if b * 2 != a:  # W_k is [layers[k], layers[k]//2] (NeuMF.py:41): the tower only chains when each layer halves
                raise ValueError('layers must halve at every step, got %r' % (self.layers,))
        # conf/NeuMF.properties defines reg_gmf / reg_mlp while NeuMF.py:15 reads reg1 / reg2
        self.reg1 = float(configs['reg1'] if 'reg1' in configs else configs['reg_gmf'])
        self.reg2 = float(configs['reg2'] if 'reg2' in configs else configs['reg_mlp'])
        if self.loss_func not in _LOSS:
            raise ValueError('pointwise loss_func must be cross_entropy or square, got %r' % self.loss_func)
        logger.info(' model_params: embed_size=%s, layers=%s, reg1=%s, reg2=%s' % (self.embed_size, self.layers, self.reg1, self.reg2) +
                    ', ' + self.model_params)

    def dense_layout(self):
        """name -> (offset, shape) inside the packed dense vector (W_k, b_k per layer, then h_neumf)."""
        out, off = {}, 0
        for k, n in enumerate(self.layers):
            out['W_%d' % k] = (off, (n, n // 2)); off += n * (n // 2)
            out['b_%d' % k] = (off, (n // 2,)); off += n // 2
        out['h_neumf'] = (off, (self.embed_size + self.layers[-1] // 2,)); off += self.embed_size + self.layers[-1] // 2
        return out, off

    def _create_params(self, init=None):
        """NeuMF.py:27-51 (no pretraining: the shipped pretrain restore cannot run)."""
        dev, kind = self.engine.device, self.optimizer.kind
        shapes = {'P_gmf': [self.data.user_nums, self.embed_size], 'Q_gmf': [self.data.item_nums, self.embed_size],
                  'P_mlp': [self.data.user_nums, self.layers[0] // 2], 'Q_mlp': [self.data.item_nums, self.layers[0] // 2]}
        self.tabs = []
        for name in ('P_gmf', 'Q_gmf', 'P_mlp', 'Q_mlp'):
            w = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer(shapes[name])
            t = Table(w.to(dev).contiguous(), kind, 'lazy')  # slots only; the dense apply realises TF's dense-moment semantics
            setattr(self, name, t)
            self.tabs.append(t)
        layout, n = self.dense_layout()
        dense = torch.zeros(n)
        for name, (off, shape) in layout.items():
            v = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer(list(shape))
            dense[off:off + v.numel()] = v.reshape(-1)
        self.dense = dense.to(dev)
        self.dense_s1 = torch.full_like(self.dense, 0.1) if kind == 'Adagrad' else (torch.zeros_like(self.dense) if kind == 'Adam' else None)
        self.dense_s2 = torch.zeros_like(self.dense) if kind == 'Adam' else None

    def build_model(self, init=None):
        self._create_params(self._with_pretrained(init))

    def _with_pretrained(self, init):
        """NeuMF.py:46-56,127-139: when both gmf_pretrain and mlp_pretrain name checkpoint directories, the GMF branch starts from a
        trained GMF ('GMF_params/{P,Q,h_gmf}'), the MLP branch from a trained MLP ('MLP_params/{P,Q,h_mlp,W_k,b_k}'), and
        h_neumf = 0.5 * concat(h_gmf, h_mlp).  The shipped conf names ./saved_model/{GMF,MLP}, which nothing ever writes in the
        reference (saver.save is commented out) so its restore raises; here missing checkpoints are logged and skipped.
        (Read literally, NeuMF.py:46-51 then REPLACES that h_neumf by a freshly initialised variable -- the `# NeuMF` line runs after
        `_load_pretrained_model()`; the NCF initialisation the method spells out is what is implemented here.)
        Explicit `init` entries win."""
        from ...utils.tools import latest_checkpoint, load_checkpoint
        if not ('gmf_pretrain' in self.configs and 'mlp_pretrain' in self.configs):
            return init
        g, m = latest_checkpoint(self.configs['gmf_pretrain']), latest_checkpoint(self.configs['mlp_pretrain'])
        if g is None or m is None:
            self.logger.info(' gmf_pretrain / mlp_pretrain hold no checkpoint: training from scratch')
            return init
        vg, vm = load_checkpoint(g), load_checkpoint(m)
        out = {'P_gmf': vg['GMF_params/P'], 'Q_gmf': vg['GMF_params/Q'], 'P_mlp': vm['MLP_params/P'], 'Q_mlp': vm['MLP_params/Q']}
        for k in range(len(self.layers)):
            out['W_%d' % k], out['b_%d' % k] = vm['MLP_params/W_%d' % k], vm['MLP_params/b_%d' % k]
        out['h_neumf'] = 0.5 * np.concatenate([vg['GMF_params/h_gmf'], vm['MLP_params/h_mlp']])
        out.update(init or {})
        self.logger.info(' restored the GMF branch from %s and the MLP branch from %s' % (g, m))
        return out

    def _variables(self):   # NeuMF.py:107-116; plus the two halves of h_neumf under the GMF / MLP names so that a NeuMF
        # checkpoint can itself seed a later run's gmf_pretrain / mlp_pretrain
        layout, _ = self.dense_layout()
        out = {'NeuMF_params/P_gmf': self.P_gmf.w, 'NeuMF_params/Q_gmf': self.Q_gmf.w, 'NeuMF_params/P_mlp': self.P_mlp.w,
               'NeuMF_params/Q_mlp': self.Q_mlp.w}
        for name, (off, shape) in layout.items():
            out['NeuMF_params/' + name] = self.dense[off:off + int(np.prod(shape))].reshape(shape)
        h = out['NeuMF_params/h_neumf']
        out['NeuMF_params/h_gmf'], out['NeuMF_params/h_mlp'] = h[:self.embed_size], h[self.embed_size:]
        return out

    def train_step(self, u_idx, i_idx, y, loss_out=None):
        """sess.run([train, loss], {u_idx, i_idx, y})  (NeuMF.py:87-95)."""
        return self.engine.train_step_neumf(self.tabs, self.dense, self.dense_s1, self.dense_s2, len(self.layers), self.optimizer, u_idx, i_idx, y,
                                            self.reg1, self.reg2, _LOSS[self.loss_func], loss_out=loss_out)

    def _train_epoch_pointwise(self, epoch, n_rows, n_batches, losses):
        for k in range(n_batches):
            lo = k * self.batch_size
            u, i, y = self.engine.sample_pointwise(self.seed, epoch, lo, min(self.batch_size, n_rows - lo), self.neg_ratio)
            self.train_step(u, i, y, loss_out=losses[k:k + 1])

    # ---- evaluation (NeuMF.py:97-105): pair logits through the tower; full rank = every (user, item) pair ----
    def test_model_loo(self):
        HR, MRR, NDCG = defaultdict(list), defaultdict(list), defaultdict(list)
        offsets, u_dev, i_dev, i_host = self._loo_feed()
        K = self.topk[-1]
        scores = self.engine.score_pairs_neumf(self.tabs, self.dense, len(self.layers), u_dev, i_dev)
        args = self.engine.topk_segments(scores, offsets, K).cpu().numpy()
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
            scores = self.engine.score_pairs_neumf(self.tabs, self.dense, len(self.layers), users.repeat_interleave(I), items.repeat(len(cur)))
            scores = self.engine.mask_seen(scores.reshape(len(cur), I), users)
            seg = torch.arange(len(cur) + 1, dtype=torch.int64, device=dev) * I
            topk_items = self.engine.topk_segments(scores.reshape(-1), seg, K).cpu().numpy()
            real_lists = [self.data.ui_test[u] for u in cur]
            for kid in range(len(self.topk)):
                hr, mrr, ndcg = batch_ranking_metrics(real_lists, topk_items, self.topk[kid])
                HR[kid].extend(hr.tolist()); MRR[kid].extend(mrr.tolist()); NDCG[kid].extend(ndcg.tolist())
        return HR, MRR, NDCG
