"""Epoch and evaluation times of every model class behind the reference's interface, at the shapes of BASELINE.json configs[0..3]
(synthetic data of those shapes: the datasets themselves are not redistributable / not all in the reference mount), with the
reference's CPU path beside it (bench.py's cpu_baseline leg: reference-algorithm Python sampler + restated TF-1 step in torch-CPU
fp32, a bounded sample of steps).
    python scripts/bench_models.py > profiles/r01_models.json          (GPU box; ~1-2 minutes)
Not a bench.py line: bench.py measures the headline metric; this is the measurement of the rows widened into."""
import gc
import json
import logging
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Data  # noqa: E402

BASE = {'model_type': 'ranking', 'saved_dir': './saved_model', 'data.split_way': 'loo', 'test.neg_samples': '99', 'test.batch_size': '1024',
        'test.interval': '1', 'topk': '[10,20]', 'epoches': '1', 'batch_size': '6144', 'lr': '0.001', 'neg_ratio': '4', 'optimizer': 'Adam',
        'init_method': 'normal', 'stddev': '0.01', 'seed': '0'}
ML1M = dict(users=6040, items=3706, mean=165)       # configs[0], [1]
CIAO = dict(users=7267, items=11211, mean=20)       # configs[3] (Ciao with item_min=5: 149 147 interactions)
EPIN = dict(users=18098, items=40000, mean=25)      # configs[2]: Epinions' item side is not in the mount; a plausible shape
MODELS = [   # (name, shape, conf as the reference ships it: conf/<name>.properties)
    ('BPR', ML1M, {'embed_size': '64', 'reg': '0.01', 'is_pairwise': 'True', 'loss_func': 'bpr'}),
    ('GMF', ML1M, {'embed_size': '32', 'reg_gmf': '1e-2', 'is_pairwise': 'False', 'loss_func': 'cross_entropy', 'init_method': 'xavier_uniform'}),
    ('MLP', ML1M, {'layers': '[128,64,32]', 'reg_mlp': '1e-2', 'is_pairwise': 'False', 'loss_func': 'cross_entropy', 'init_method': 'xavier_uniform'}),
    ('NeuMF', ML1M, {'embed_size': '32', 'layers': '[128,64,32]', 'reg_gmf': '1e-2', 'reg_mlp': '1e-3', 'is_pairwise': 'False',
                     'loss_func': 'cross_entropy', 'init_method': 'xavier_uniform'}),
    ('CML', EPIN, {'embed_size': '128', 'margin': '1.0', 'reg': '10.0', 'cml_like': 'True', 'is_pairwise': 'False', 'loss_func': 'hinge',
                   'neg_ratio': '20', 'init_method': 'xavier'}),
    ('TransCF', CIAO, {'embed_size': '64', 'margin': '0.5', 'reg1': '0.1', 'reg2': '0.01', 'cml_like': 'True', 'is_pairwise': 'True', 'loss_func': 'hinge'}),
    ('LRML', CIAO, {'embed_size': '128', 'mem_size': '50', 'margin': '0.2', 'reg': '0.001', 'neg_ratio': '1', 'cml_like': 'True', 'is_pairwise': 'True',
                    'loss_func': 'hinge'}),
    ('FISM', CIAO, {'embed_size': '128', 'alpha': '0.4', 'reg': '1e-3', 'reg_bias': '1e-3', 'fism_like': 'True', 'is_pairwise': 'True', 'loss_func': 'bpr',
                    'init_method': 'xavier_uniform'}),
    ('NAIS_single', CIAO, {'embed_size': '128', 'atten_size': '32', 'atten_type': "'prod'", 'beta': '0.5', 'reg': '1e-3', 'optimizer': 'Adagrad', 'lr': '0.01',
                           'nais_like': 'True', 'is_pairwise': 'False', 'loss_func': 'cross_entropy', 'init_method': 'xavier_uniform'}),
    ('SBPR', CIAO, {'embed_size': '128', 'reg': '0.05', 'neg_ratio': '10', 'is_pairwise': 'True', 'loss_func': 'bpr', 'social_file': 'trusts.csv'}),
]


def make_data(shape, seed, friends=False):
    rng = np.random.default_rng(seed)
    U, I = shape['users'], shape['items']
    ui_train, ui_test = {}, {}
    for u in range(U):
        n = int(min(I // 2, max(4, rng.poisson(shape['mean']))))
        items = rng.choice(I, size=n + 1, replace=False)
        ui_train[u] = items[:n].tolist()
        cand = np.setdiff1d(rng.choice(I, size=160, replace=False), items)   # 99 evaluation negatives outside the user's items
        ui_test[u] = cand[:99].tolist() + [int(items[n])]
    d = Data(U, I, ui_train, ui_test)
    if friends:
        d.user_friends = {u: rng.choice(U, size=int(rng.integers(1, 16)), replace=False).tolist() for u in range(U) if u % 10 != 9}
    return d


def rows_per_epoch(name, data, cfg):
    n_pos = sum(len(v) for v in data.ui_train.values())
    R = int(cfg['neg_ratio'])
    if name == 'CML':
        return n_pos
    if cfg['is_pairwise'] == 'True':
        return n_pos * R
    return n_pos * (R + 1)


def main():
    import importlib
    log = logging.getLogger('bench_models')
    out, cache = [], {}
    only = set(sys.argv[1:])
    for name, shape, conf in MODELS:
        if only and name not in only:
            continue
        key = (shape['users'], name == 'SBPR')
        if key not in cache:
            cache[key] = make_data(shape, 1, friends=name == 'SBPR')
        data = cache[key]
        cfg = dict(BASE, recommender=name)
        cfg.update(conf)
        cls = getattr(importlib.import_module('cleverrec_b200.model.ranking.' + name), name)
        t0 = time.perf_counter()
        m = cls(None, data, cfg, log)
        m.build_model()
        torch.cuda.synchronize()
        setup = time.perf_counter() - t0
        m.train_model()                                   # warm epoch (allocations, first-launch costs)
        torch.cuda.synchronize()
        gc.collect()                                      # keep a generation-2 collection of earlier models' objects out of the timed regions
        l0 = m.engine.launches
        t0 = time.perf_counter()
        loss = m.train_model()
        torch.cuda.synchronize()
        epoch_s = time.perf_counter() - t0
        epoch_launches = m.engine.launches - l0
        m.test_model_loo()
        torch.cuda.synchronize()
        gc.collect()
        t0 = time.perf_counter()
        HR, MRR, NDCG = m.test_model_loo()
        torch.cuda.synchronize()
        eval_s = time.perf_counter() - t0
        rows = rows_per_epoch(name, data, cfg) if name != 'SBPR' else m.engine.epoch_rows(int(cfg['neg_ratio']), 'sbpr')
        rec = {"model": name, "shape": shape, "conf": conf, "rows_per_epoch": rows, "steps_per_epoch": (len(data.ui_train) if name == 'NAIS_single' else math.ceil(rows / 6144)),
               "setup_s": setup, "epoch_s": epoch_s, "rows_per_s": rows / epoch_s, "loo_eval_s": eval_s, "loo_eval_users_per_s": len(m.test_users) / eval_s,
               "loss": loss, "hr10": float(np.mean(HR[0]))}
        # what bounds an epoch at these shapes: the step is a chain of dependent microsecond kernels (the tables are L2-resident), so
        # the floor of a step is (library kernels per step) x (one dependent launch + one L2 round trip ~ 2.5 us), not bytes / bandwidth
        steps = rec["steps_per_epoch"]
        rec["us_per_step"] = 1e6 * epoch_s / steps
        rec["library_kernels_per_step"] = epoch_launches / steps
        rec["launch_chain_floor_us_per_step"] = 2.5 * epoch_launches / steps
        out.append(rec)
        sys.stderr.write("%-12s epoch %.3f s (%.3e rows/s, %d steps)  loo eval %.3f s\n" % (name, epoch_s, rec["rows_per_s"], rec["steps_per_epoch"], eval_s))
        m.engine.close()
    # the reference's CPU path for the headline model at this shape: bench.py's cpu_baseline leg (the one place outside tests/ that
    # may execute oracle/) -- reference-algorithm Python sampler + restated TF-1 BPR/Adam step, one 6144-row batch per step
    from bench import cpu_reference_steps
    threads = os.cpu_count() or 1
    times, done, n_users, _ = cpu_reference_steps(dict(users=ML1M['users'], items=ML1M['items'], dim=64, mean_hist=ML1M['mean'], batch=6144, neg_ratio=4),
                                                  6, 6144, threads)
    sec = sum(times[1:]) / len(times[1:])
    cpu = {"what": "reference CPU path restated (bench.py cpu_reference_steps), BPR at the ml-1m shape: Python sampler + row-sparse TF-1 Adam step "
                   "(torch-CPU fp32) per 6144-row batch", "cores": threads, "rows_per_step": done, "step_s": sec, "rows_per_s": done / sec}
    print(json.dumps({"models": out, "cpu_reference_bpr_ml1m": cpu}, indent=1))


if __name__ == "__main__":
    main()
