# coding: utf-8
""" Base Recommender for Ranking/Rating Model -- mirror of the reference model/Recommender.py:9-40.

Same constructor `(sess, data, configs, logger)`, same config keys, same abstract methods.  `sess` is the
reference's tf.Session slot: pass None (an Engine is created on `gpu.id`... see engine_device) or an existing
cleverrec_b200.engine.Engine to share one device handle between models."""
import torch

from ..engine import Engine
from ..utils.tools import get_initializer, get_optimizer


class Recommender(object):
    def __init__(self, sess, data, configs, logger):
        self.model = configs['recommender']
        self.sess, self.data, self.configs, self.logger = sess, data, configs, logger
        # Common parameters
        self._get_common_params()

    def _get_common_params(self):
        c = self.configs
        self.epoches, self.batch_size, self.batch_size_t, self.lr, self.neg_samples = int(c['epoches']), int(c['batch_size']), \
            int(c['test.batch_size']), float(c['lr']), int(c['test.neg_samples'])
        self.fism_like, self.cml_like = 'fism_like' in c, 'cml_like' in c
        self.is_pairwise = c['is_pairwise']  # the *string* 'True'/'False', compared as such (RankingRecommender.py:35)
        # B200-path keys (all optional; defaults keep the reference's behaviour)
        self.seed = int(c.get('seed', 0))
        self.adam_mode = c.get('adam_mode', 'tf1')          # 'tf1' = tf.train.AdamOptimizer semantics, 'lazy' = LazyAdam
        self.sampler_mode = c.get('sampler', 'philox')        # 'philox' | 'numpy_stream' (the reference's np.random stream, bit for bit)
        if self.sampler_mode not in ('philox', 'numpy_stream'):
            raise ValueError("sampler must be 'philox' or 'numpy_stream'")
        self.score_exact = c.get('score_exact', 'False') == 'True'  # full-rank eval on CUDA cores instead of tcgen05
        self.init_generator = torch.Generator().manual_seed(self.seed)
        self.initializer = get_initializer(c['init_method'], float(c['stddev']), generator=self.init_generator)
        if self.initializer is None:
            raise ValueError('unknown init_method %r' % c['init_method'])
        self.regularizer = None  # the reference attaches an l2_regularizer that never reaches any loss
        self.loss_func = c['loss_func']
        self.optimizer = get_optimizer(c['optimizer'], self.lr, adam_mode=self.adam_mode)
        if self.optimizer is None:
            raise ValueError('unknown optimizer %r' % c['optimizer'])
        self.saved_model_dir = c['saved_dir']
        self.T = int(c['test.interval'])  # Test every T epoches
        self.topk = list(map(int, c['topk'][1:-1].split(',')))
        self.model_params = 'lr=%s, loss_func=%s' % (self.lr, c['loss_func'])
        # one process per GPU (torchrun): rank / world from torch.distributed when it is initialised, else from the launcher's env
        import os
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
        else:
            self.rank, self.world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
        if c.get('dist.mode', 'auto') == 'replica':   # an independent single-GPU model inside a multi-process job (main_tuning)
            self.rank, self.world = 0, 1
        # device handle
        if isinstance(self.sess, Engine):
            self.engine = self.sess
        else:
            # CRB_SHARED_DEVICE=1: every rank of the job uses device 0 (CUDA IPC between processes on ONE GPU, gloo control plane):
            # the multi-GPU code path on a single-GPU box (parity tests)
            shared = os.environ.get('CRB_SHARED_DEVICE', '0') == '1'
            local = os.environ.get('LOCAL_RANK', '0') if int(os.environ.get('WORLD_SIZE', '1')) > 1 and not shared else 0
            self.engine = Engine(int(c.get('engine.device', local)))

    def build_model(self):
        raise NotImplementedError

    # ---- checkpoints (utils/tools.py save_checkpoint): same variable names as the reference's per-model Saver maps ----
    def _variables(self):
        """dict Saver-name -> tensor, e.g. {'BPR_params/P': ..., 'BPR_params/Q': ...}."""
        raise NotImplementedError

    def _flush_before_read(self):
        """Tables under CRB_ADAM_TF1 carry pending decay-only steps until flushed."""
        from ..engine import Table
        seen = set()
        for v in list(vars(self).values()) + list(getattr(self, 'tables', {}).values()) + list(getattr(self, 'tabs', [])):
            if isinstance(v, Table) and v.last is not None and id(v) not in seen:   # only CRB_ADAM_TF1 tables defer decay steps
                seen.add(id(v))
                self.engine.adam_flush(v, self.optimizer)

    def save_model(self, step=None):
        """What `self.saver.save(self.sess, saved_dir/model/model)` (RankingRecommender.py:433, commented out in the reference)
        would write.  Enabled from run_model by the config key save_model=True."""
        import os
        from ..utils.tools import save_checkpoint
        self._flush_before_read()
        torch.cuda.synchronize(self.engine.device)
        return save_checkpoint(os.path.join(self.saved_model_dir, self.model), self.model, self._variables(), step)

    def train_model(self):
        raise NotImplementedError

    def test_model(self):
        raise NotImplementedError

    def run_model(self):
        raise NotImplementedError
