# coding: utf-8
""" Matrix Factorization.  The reference ships no source for model/ranking/MF.py (an empty module and a stale
conf/MF.properties), so MF is specified here: BPR's dot product (BPR.py:39) trained pointwise with
get_loss('square' | 'cross_entropy') (utils/tools.py:66-76) and L2 on the gathered rows, or pairwise with 'bpr' (= BPR).
Parity for MF is unpinned by the reference; it is pinned against oracle/tf1_restatement.mf_loss. """
from ... import _lib
from .GMF import GMF


class MF(GMF):
    score_kind = _lib.SCORE_DOT

    def __init__(self, sess, data, configs, logger):
        configs = dict(configs)
        # conf/MF.properties spells the keys differently (loss_function / reg_mf, quoted values): accept them
        if 'loss_func' not in configs and 'loss_function' in configs:
            configs['loss_func'] = configs['loss_function'].strip("'\"")
        if 'reg' not in configs and 'reg_mf' in configs:
            configs['reg'] = configs['reg_mf']
        super(MF, self).__init__(sess, data, configs, logger)
