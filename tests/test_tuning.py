"""Grid search (cleverrec_b200/main_tuning.py; reference main_tuning.py:38-66): grid order, dealing of the grid to ranks (gloo,
world_size 2, with the model run stubbed out -- no GPU), and -m gpu: concurrent workers on one GPU give the sequential results."""
import logging
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, synthetic_data


def test_grid_follows_the_reference_loop_order():
    from cleverrec_b200.main_tuning import grid
    g = grid({'embed_size': '[32,64]', 'reg': '[0.01, 0.1,1]', 'neg_ratio': '[4]'})
    # main_tuning.py:41-43: for embed_size: for reg: for neg_ratio
    assert [(c['embed_size'], c['reg'], c['neg_ratio']) for c in g] == [(32, .01, 4), (32, .1, 4), (32, 1., 4), (64, .01, 4), (64, .1, 4), (64, 1., 4)]
    assert all(isinstance(c['embed_size'], int) and isinstance(c['reg'], float) and isinstance(c['neg_ratio'], int) for c in g)
    assert grid({'embed_size': '16', 'reg': '0.5', 'neg_ratio': '[1,2]'}) == [{'embed_size': 16, 'reg': .5, 'neg_ratio': 1}, {'embed_size': 16, 'reg': .5, 'neg_ratio': 2}]


def test_workers_and_numpy_stream_are_exclusive():
    from cleverrec_b200.main_tuning import run_grid
    with pytest.raises(ValueError):
        run_grid({'embed_size': '[8]', 'reg': '[0.1]', 'neg_ratio': '[1]', 'sampler': 'numpy_stream'}, None, logging.getLogger('t'), workers=2, rank=0, world=1)


WORKER = r'''
import os, sys, logging
import torch.distributed as dist
sys.path.insert(0, %r)
from cleverrec_b200 import main_tuning as MT
dist.init_process_group("gloo")
rank = dist.get_rank()
ran = []
def fake(configs, combo, data, logger, device):
    ran.append(combo)
    return {"params": combo, "best_epoch": rank + 1, "best_metrics": {0: (0.0, 0.0, combo["embed_size"] * combo["reg"])}}
MT._run_one = fake
cfg = {"embed_size": "[8,16,32]", "reg": "[0.1,0.2]", "neg_ratio": "[1]", "tuning.workers": "2"}
res = MT.run_grid(cfg, None, logging.getLogger("t"))
g = MT.grid(cfg)
assert [r["params"] for r in res] == g                                   # grid order on every rank
assert [r["best_epoch"] for r in res] == [1 + (k %% 2) for k in range(6)]   # combination k ran on rank k mod world
assert ran == g[rank::2]
assert MT.best_of(res)["params"] == {"embed_size": 32, "reg": 0.2, "neg_ratio": 1}
if rank == 0:
    print("TUNING_OK")
dist.destroy_process_group()
'''


def test_grid_is_dealt_to_ranks_and_gathered_in_order(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % (ROOT,))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           str(script)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0 and "TUNING_OK" in r.stdout, r.stdout[-3000:]


@pytest.mark.gpu
def test_concurrent_workers_match_sequential():
    from cleverrec_b200.main_tuning import run_grid
    cfg = {'model_type': 'ranking', 'recommender': 'BPR', 'saved_dir': './saved_model', 'data.split_way': 'loo', 'test.neg_samples': '49',
           'test.batch_size': '64', 'test.interval': '1', 'topk': '[5,10]', 'epoches': '3', 'batch_size': '256', 'lr': '0.01', 'optimizer': 'Adagrad',
           'init_method': 'normal', 'stddev': '0.05', 'seed': '3', 'is_pairwise': 'True', 'loss_func': 'bpr',
           'embed_size': '[16,32]', 'reg': '[0.01,0.1]', 'neg_ratio': '[2]'}
    data = synthetic_data(120, 300, 12, seed=11, test_per_user=1)
    rs = np.random.RandomState(0)
    for u in data.ui_test:
        cand = np.setdiff1d(np.arange(data.item_nums), data.ui_train[u])
        data.ui_test[u] = rs.choice(cand, 49, replace=False).tolist() + data.ui_test[u]
    log = logging.getLogger('test')
    seq = run_grid(cfg, data, log, workers=1, rank=0, world=1)
    par = run_grid(cfg, data, log, workers=4, rank=0, world=1)
    assert [r['params'] for r in seq] == [r['params'] for r in par] and len(seq) == 4
    for a, b in zip(seq, par):
        assert a['best_epoch'] == b['best_epoch']
        for k in a['best_metrics']:
            np.testing.assert_allclose(a['best_metrics'][k], b['best_metrics'][k], rtol=0, atol=1e-12)
