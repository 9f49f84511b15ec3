# coding: utf-8
"""Factories of the reference's utils/tools.py:9-87 for the B200 path: re_index, timer, logger, initializer and
optimizer.  (get_loss lives in the CUDA kernels; the social/graph helpers :116-297 are out of scope.)"""
import functools
import logging
import math
import os
import sys
import time

import torch

from ..engine import Optimizer


# Reindex ids (utils/tools.py:9-15)
def re_index(data_set):
    data_map = {}
    id = 0
    for d in data_set:
        data_map[d] = id
        id += 1
    return data_map


# Time cost decorator (utils/tools.py:18-28)
def timer(text):
    def decorator(func):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            t1 = time.time()
            print('Start %s...' % text)
            res = func(*args, **kwargs)
            print('%s done, time: %s' % (text, time.strftime('%H:%M:%S', time.gmtime(time.time() - t1))))
            return res
        return wrapper
    return decorator


# Logger (utils/tools.py:31-48): same format, file + stdout
def get_logger(log_dir, model):
    if not os.path.exists(log_dir):
        os.makedirs(log_dir)
    logger = logging.getLogger()
    logger.setLevel(logging.DEBUG)
    formatter = logging.Formatter('%(asctime)s  %(message)s', datefmt='%Y-%m-%d %H:%M:%S')
    fh = logging.FileHandler(os.path.join(log_dir, model + '.log'))
    fh.setLevel(logging.DEBUG)
    fh.setFormatter(formatter)
    ch = logging.StreamHandler(sys.stdout)
    ch.setLevel(logging.DEBUG)
    ch.setFormatter(formatter)
    logger.addHandler(fh)
    logger.addHandler(ch)
    return logger


# Initializer (utils/tools.py:51-63).  Returns a callable shape -> fp32 CPU tensor (the reference's initializer
# objects are called as `self.initializer(shape_)`).  'xavier_uniform' (used by the shipped FISM/GMF/NeuMF/NAIS
# confs but unknown to the reference factory) is treated as 'xavier'.
def get_initializer(init_method, stddev=None, generator=None):
    init_method = init_method.strip()

    def fans(shape):
        if len(shape) == 1:
            return shape[0], shape[0]
        return shape[0], shape[1]

    def normal(shape):
        return torch.randn(*shape, generator=generator) * stddev

    def tnormal(shape):  # truncated at 2 sigma, re-drawn
        t = torch.empty(*shape)
        torch.nn.init.trunc_normal_(t, mean=0.0, std=stddev, a=-2 * stddev, b=2 * stddev, generator=generator)
        return t

    def uniform(shape):
        return (torch.rand(*shape, generator=generator) * 2 - 1) * stddev

    def xavier(shape):
        fi, fo = fans(shape)
        limit = math.sqrt(6.0 / (fi + fo))
        return (torch.rand(*shape, generator=generator) * 2 - 1) * limit

    def xavier_normal(shape):  # variance_scaling_initializer(factor=1, FAN_AVG, uniform=False): truncated normal
        fi, fo = fans(shape)
        sd = math.sqrt(1.3 * 2.0 / (fi + fo))
        t = torch.empty(*shape)
        torch.nn.init.trunc_normal_(t, mean=0.0, std=sd, a=-2 * sd, b=2 * sd, generator=generator)
        return t

    table = {'normal': normal, 'tnormal': tnormal, 'uniform': uniform, 'xavier': xavier, 'xavier_uniform': xavier,
             'xavier_normal': xavier_normal}
    return table.get(init_method)


# Optimizer (utils/tools.py:79-87)
def get_optimizer(optimizer, lr, adam_mode='tf1'):
    optimizer = optimizer.strip().strip("'\"")
    if optimizer in ('SGD', 'Adam', 'Adagrad'):
        return Optimizer(optimizer, lr, adam_mode=adam_mode)
    return None


# ------------------------------------------------------------------------------------------------ checkpoints
# The reference builds a tf.train.Saver per model with an explicit variable-name map (BPR.py:53-58, GMF.py:59-64,
# NeuMF.py:107-116, FISM.py:72-77, NAIS_single.py:99-106) and restores by those names for pretraining (NeuMF.py:127-139,
# NAIS_single.py:35-38); its `saver.save` call is commented out (RankingRecommender.py:432-433).  Here a checkpoint is one
# `.npz` per save under <saved_dir>/<model>/, keyed by the SAME variable names (typos such as 'FISM_paras/P' included, because
# the restore side spells them the same way), fp32 arrays.
def save_checkpoint(directory, model, variables, step=None):
    """variables: dict name -> torch tensor / ndarray.  Returns the file written."""
    import numpy as np
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, model + ('' if step is None else '-%d' % step) + '.npz')
    arrays = {}
    for name, v in variables.items():
        a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        arrays[name.replace('/', '__')] = np.ascontiguousarray(a, dtype=np.float32)
    tmp = path + '.tmp.npz'
    np.savez(tmp, **arrays)
    os.replace(tmp, path)
    return path


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: newest `.npz` in the directory, or None."""
    if not directory or not os.path.isdir(directory):
        return None
    files = [os.path.join(directory, f) for f in os.listdir(directory) if f.endswith('.npz') and not f.endswith('.tmp.npz')]
    return max(files, key=os.path.getmtime) if files else None


def load_checkpoint(path):
    """-> dict variable name -> fp32 ndarray."""
    import numpy as np
    with np.load(path) as z:
        return {k.replace('__', '/'): z[k] for k in z.files}


# Get SPu for SBPR (reference utils/tools.py:115-127): the items u's friends consumed and u did not, as a list per user.
def get_SPu(data):
    """The list order is the iteration order of the Python set the reference builds, and the sampler indexes the list by position
    (utils/sampler.py:114-115), so the same sequence of set operations is performed here: per friend, union then difference."""
    SPu = {}
    friends_of = data.user_friends
    for u in data.ui_train:
        if u not in friends_of:
            continue
        own = set(data.ui_train[u])
        social = set()
        for friend in friends_of[u]:
            if friend in data.ui_train:
                social = social.union(set(data.ui_train[friend])).difference(own)
        if social:
            SPu[u] = list(social)
    return SPu
