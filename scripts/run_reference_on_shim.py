#!/usr/bin/env python
"""Run the GENUINE reference (its own preprocessing, sampler, model class, epoch loop, evaluation and log lines) on the CPU without
TensorFlow, with oracle/tf1_shim.py standing in for `tensorflow` -- test infrastructure, fp64, slow (Python sampler, lazy graph):

    python scripts/run_reference_on_shim.py /path/to/CleverRec [key=value ...]
    python scripts/run_reference_on_shim.py /root/reference recommender=BPR epoches=2 data.dataset=ml-100k data.file_name=u.data \\
        "data.sep=\t" data.format=UIRT data.split_way=loo test.neg_samples=99 init_method=normal embed_size=16

Overrides are the reference's own config keys (CleverRec.properties / conf/<Model>.properties).  What the shim does and does not pin:
oracle/tf1_shim.py's header."""
import logging
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refimport as R  # noqa: E402
from oracle import tf1_shim as tf  # noqa: E402


def main(argv):
    if not argv:
        raise SystemExit(__doc__)
    R.REFERENCE_ROOT = os.path.abspath(argv[0])
    over = dict(a.split("=", 1) for a in argv[1:])
    over = {k: v.encode().decode("unicode_escape") for k, v in over.items()}
    name = over.get("recommender") or R.default_configs()["recommender"]
    cfg = R.default_configs(**dict(over, recommender=name))
    cfg["init_method"] = cfg["init_method"].strip()
    if cfg["init_method"] == "xavier_uniform":      # unknown to utils/tools.py:51-63
        cfg["init_method"] = "xavier"
    for alias, key in (("reg_gmf", "reg"), ("reg_mlp", "reg"), ("reg_gmf", "reg1"), ("reg_mlp", "reg2")):   # GMF / MLP / NeuMF read these
        if alias in cfg and key not in cfg:
            cfg[key] = cfg[alias]
    cfg["data.root_dir"] = os.path.join(R.REFERENCE_ROOT, "dataset") if not os.path.isabs(cfg["data.root_dir"]) else cfg["data.root_dir"]
    logging.basicConfig(level=logging.INFO, format="%(asctime)s  %(message)s", datefmt="%Y-%m-%d %H:%M:%S", stream=sys.stdout)
    logger = logging.getLogger("reference-on-shim")
    np.random.seed(int(over.get("seed", 0)))
    data = R.load().RankingPreprocess(cfg, logger)
    classes = tf.load_reference_models(R.REFERENCE_ROOT, sorted({name, "MLP"}))
    classes[name].__init__.__globals__.setdefault("get_loss", classes["MLP"].__init__.__globals__["get_loss"])   # GMF.py never imports it
    tf.reset_default_graph()
    tf.seed_initializers(int(over.get("seed", 0)))
    model = classes[name](tf.Session(), data, cfg, logger)
    if name == "NAIS_single":                        # NAIS_single.py:87 calls the config string
        model.loss_func = tf.nn.sigmoid_cross_entropy_with_logits
    model.run_model()


if __name__ == "__main__":
    main(sys.argv[1:])
