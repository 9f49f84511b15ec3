"""-m gpu: the whole drop-in flow of the reference's main.py:16-55 through cleverrec_b200.main -- CleverRec.properties +
conf/<recommender>.properties (written here in the reference's format, its shipped values for BPR) -> packaged RankingPreprocess
on a ratings file -> model class resolved by name -> run_model() with the reference's log lines."""
import logging
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

DEFAULTS = """[default]
recommender=BPR
model_type=ranking
config_dir=./conf
saved_dir=./saved_model
data.root_dir=./dataset
data.dataset=toy
data.file_name=ratings.csv
data.sep=,
data.format=UI
data.split_way=%(split)s
data.split_ratio=[0.7,0.2,0.1]
data.split_by_time=False
data.user_min=0
data.item_min=5
gpu.is_gpu=True
gpu.id=0
gpu.mem_frac=0.90
test.neg_samples=%(neg)s
test.batch_size=1024
test.interval=1
metrics=['precision', 'recall', 'ndcg', 'mrr', 'map']
topk=[10,20]
log.dir=./logs
"""
BPR_CONF = """[parameters]
epoches=3
batch_size=6144
embed_size=64
reg=0.01
lr=0.01
neg_ratio=4
optimizer=Adam
is_pairwise=True
loss_func=bpr
init_method=normal 
stddev=0.01
"""


def write_tree(root, split, neg):
    rs = np.random.RandomState(0)
    os.makedirs(root / "conf"); os.makedirs(root / "dataset" / "toy")
    (root / "CleverRec.properties").write_text(DEFAULTS % {"split": split, "neg": neg})
    (root / "conf" / "BPR.properties").write_text(BPR_CONF)
    rows = ["user,item"]
    for u in range(400):   # clustered preferences so that three epochs learn something
        base = (u % 8) * 60
        for i in np.unique(np.r_[rs.randint(base, base + 60, 25), rs.randint(0, 480, 3)]):
            rows.append("%d,%d" % (1000 + u, 5000 + i))
    (root / "dataset" / "toy" / "ratings.csv").write_text("\n".join(rows) + "\n")


@pytest.mark.parametrize("split,neg", [("loo", 99), ("rs", 0)])
def test_main_flow_in_process(tmp_path, split, neg, caplog):
    from cleverrec_b200 import main
    write_tree(tmp_path, split, neg)
    cfg = main.load_configs(str(tmp_path))
    cfg["data.root_dir"], cfg["saved_dir"] = str(tmp_path / "dataset"), str(tmp_path / "saved_model")
    assert cfg["recommender"] == "BPR" and cfg["embed_size"] == "64" and cfg["init_method"] == "normal"   # configparser strips the trailing space
    from cleverrec_b200.model.RankingPreprocess import RankingPreprocess
    np.random.seed(1)
    data = RankingPreprocess(cfg, logging.getLogger("flow"))
    assert data.user_nums == 400 and data.train_rows[0].dtype == np.int32
    with caplog.at_level(logging.INFO):
        best_epoch, best = main.run(cfg, data, logging.getLogger("flow"))
    assert "Training loss:" in caplog.text and "best_epoch:" in caplog.text
    hr10 = best[0][0]
    assert best_epoch >= 1 and hr10 > (0.18 if split == "loo" else 0.02), best   # chance is 0.10 (loo, 99 negatives) and ~0.01 (rs)


def test_main_module_as_a_script(tmp_path):
    write_tree(tmp_path, "loo", 99)
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "cleverrec_b200.main", "."], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "Current model: BPR" in r.stdout and " epoch 3" in r.stdout and "best_epoch:" in r.stdout
    assert os.path.exists(tmp_path / "logs" / "BPR.log")


SBPR_CONF = """[parameters]
social_file=trusts.csv
epoches=3
batch_size=6144
embed_size=64
reg=0.05
lr=0.01
neg_ratio=4
optimizer=Adam
is_pairwise=True
loss_func=bpr
init_method=normal 
stddev=0.01
"""
TUNING_CONF = """[parameters]
epoches=2
batch_size=6144
embed_size=[16,32]
reg=[0.01,0.1]
neg_ratio=[4]
lr=0.01
optimizer=Adam
is_pairwise=True
loss_func=bpr
init_method=normal
stddev=0.01
tuning.workers=2
"""


def test_sbpr_through_the_entry_point_with_a_trust_file(tmp_path):
    """recommender=SBPR with conf/SBPR.properties in the reference's format (social_file=trusts.csv): the packaged RankingPreprocess
    reads the trust pairs (RankingPreprocess.py:49-58), SBPR builds SPu and trains on the social sampler, run_model logs as usual."""
    write_tree(tmp_path, "loo", 99)
    (tmp_path / "conf" / "SBPR.properties").write_text(SBPR_CONF)
    (tmp_path / "CleverRec.properties").write_text((DEFAULTS % {"split": "loo", "neg": 99}).replace("recommender=BPR", "recommender=SBPR"))
    rs = np.random.RandomState(3)
    rows = ["trustor,trustee"]
    for u in range(400):   # friends mostly from the user's own preference cluster, plus ids that never rated (dropped by the filter)
        for v in np.unique(np.r_[rs.randint(0, 50, 4) * 8 + (u % 8), rs.randint(0, 400, 1)]):
            if v != u:
                rows.append("%d,%d" % (1000 + u, 1000 + v))
        rows.append("%d,%d" % (1000 + u, 99999))
    (tmp_path / "dataset" / "toy" / "trusts.csv").write_text("\n".join(rows) + "\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "cleverrec_b200.main", "."], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "Current model: SBPR" in r.stdout and " epoch 3" in r.stdout and "best_epoch:" in r.stdout
    import re
    hr = [float(x) for x in re.findall(r"\(k=10\) HR=([0-9.]+)", r.stdout)]
    assert hr and max(hr) > 0.15, r.stdout[-1500:]      # chance is 0.10 with 99 negatives


def test_main_tuning_as_a_script(tmp_path):
    """python -m cleverrec_b200.main_tuning: the reference's grid (main_tuning.py:38-45) over bracketed lists in the model's conf,
    two combinations at a time on the GPU; one run_model log per combination and the summary lines."""
    write_tree(tmp_path, "loo", 99)
    (tmp_path / "conf" / "BPR.properties").write_text(TUNING_CONF)
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "cleverrec_b200.main_tuning", "."], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("best_epoch:") == 4 and r.stdout.count("[tuning ") == 4
    assert "best by NDCG@topk[0]:" in r.stdout and "'embed_size': 32" in r.stdout
