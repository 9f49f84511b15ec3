# coding: utf-8
" Multi-Layer Perceptron (2017 NCF) -- mirror of the reference model/ranking/MLP.py: NeuMF's tower on its own. "
import numpy as np
import torch

from .NeuMF import NeuMF, _LOSS
from .. import RankingRecommender as _rr
from ...engine import Table


class MLP(NeuMF):
    def __init__(self, sess, data, configs, logger):
        _rr.RankingRecommender.__init__(self, sess, data, configs, logger)
        self.layers = list(map(int, configs['layers'][1:-1].split(',')))   # MLP.py:13
        for a, b in zip(self.layers[:-1], self.layers[1:]):
            if b * 2 != a:
                raise ValueError('layers must halve at every step, got %r' % (self.layers,))
        # conf/MLP.properties defines reg_mlp while MLP.py:14 reads 'reg' (SURVEY 2.3): accept both
        self.reg = float(configs['reg'] if 'reg' in configs else configs['reg_mlp'])
        self.embed_size, self.reg1, self.reg2 = 0, 0.0, self.reg   # no GMF branch: the fused kernels run with E = 0
        if self.loss_func not in _LOSS:
            raise ValueError('pointwise loss_func must be cross_entropy or square, got %r' % self.loss_func)
        logger.info(' model_params: layers=%s, reg=%s' % (self.layers, self.reg) + ', ' + self.model_params)

    def _create_params(self, init=None):
        """MLP.py:25-41: P, Q [*, layers[0]//2], W_k / b_k per layer, h_mlp [layers[-1]//2] (stored where NeuMF keeps h_neumf)."""
        dev, kind = self.engine.device, self.optimizer.kind
        shapes = {'P': [self.data.user_nums, self.layers[0] // 2], 'Q': [self.data.item_nums, self.layers[0] // 2]}
        for name, attr in (('P', 'P_mlp'), ('Q', 'Q_mlp')):
            w = torch.as_tensor(np.asarray(init[name]), dtype=torch.float32) if init and name in init else self.initializer(shapes[name])
            setattr(self, attr, Table(w.to(dev).contiguous(), kind, 'lazy'))
        self.P, self.Q = self.P_mlp, self.Q_mlp
        self.tabs = [None, None, self.P_mlp, self.Q_mlp]
        layout, n = self.dense_layout()
        dense = torch.zeros(n)
        for name, (off, shape) in layout.items():
            key = 'h_mlp' if name == 'h_neumf' else name
            v = torch.as_tensor(np.asarray(init[key]), dtype=torch.float32) if init and key in init else self.initializer(list(shape))
            dense[off:off + v.numel()] = v.reshape(-1)
        self.dense = dense.to(dev)
        self.dense_s1 = torch.full_like(self.dense, 0.1) if kind == 'Adagrad' else (torch.zeros_like(self.dense) if kind == 'Adam' else None)
        self.dense_s2 = torch.zeros_like(self.dense) if kind == 'Adam' else None

    def build_model(self, init=None):
        self._create_params(init)

    def _variables(self):   # MLP.py:79-86
        layout, _ = self.dense_layout()
        out = {'MLP_params/P': self.P.w, 'MLP_params/Q': self.Q.w}
        for name, (off, shape) in layout.items():
            out['MLP_params/' + ('h_mlp' if name == 'h_neumf' else name)] = self.dense[off:off + int(np.prod(shape))].reshape(shape)
        return out
