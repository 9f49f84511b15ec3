"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *genuine* reference host code (pure NumPy parts) from
``/root/reference`` with ``tensorflow`` and ``gensim`` replaced by empty stub
modules.  Only usable inside the build container: the GPU box has
no ``/root/reference``, so everything this module produces that a GPU test needs
is committed as a fixture under ``tests/golden/`` by ``oracle/make_golden.py``.

What is exposed (all unmodified reference code):
  * utils/sampler.py:10-99   pointwise_ranking_sampler / pairwise_ranking_sampler / ranking_sampler_cml
  * utils/sampler.py:102-141 ranking_sampler_sbpr;  utils/tools.py:115-127 get_SPu
  * utils/metrics.py:9-19    cal_ranking_metrics
  * model/RankingPreprocess.py:13-134  RankingPreprocess
  * model/RankingRecommender.py:33-100,198-348  train/test loops, driven by a FakeSession
"""
import importlib
import logging
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CLEVERREC_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "utils"))


class _Anything(types.ModuleType):
    """Stub module: attribute access returns another stub so `tf.contrib.layers...` parses."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return None


_loaded = {}


def load():
    """Return a namespace with the reference's own functions/classes."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("tensorflow", "gensim", "gensim.models", "gensim.models.word2vec"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    # the reference packages are called `utils` and `model`; import them under a
    # private sys.path entry and make sure we do not shadow/receive a foreign `utils`
    for name in list(sys.modules):
        if name in ("utils", "model") or name.startswith(("utils.", "model.")):
            del sys.modules[name]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        sampler = importlib.import_module("utils.sampler")
        metrics = importlib.import_module("utils.metrics")
        tools = importlib.import_module("utils.tools")
        prep = importlib.import_module("model.RankingPreprocess")
        rr = importlib.import_module("model.RankingRecommender")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    _loaded.update(
        pairwise_ranking_sampler=sampler.pairwise_ranking_sampler,
        pointwise_ranking_sampler=sampler.pointwise_ranking_sampler,
        ranking_sampler_cml=sampler.ranking_sampler_cml,
        ranking_sampler_sbpr=sampler.ranking_sampler_sbpr,
        get_SPu=tools.get_SPu,
        cal_ranking_metrics=metrics.cal_ranking_metrics,
        cal_rmse_mae=metrics.cal_rmse_mae,
        re_index=tools.re_index,
        RankingPreprocess=prep.RankingPreprocess,
        RankingRecommender=rr.RankingRecommender,
    )
    return types.SimpleNamespace(**_loaded)


class Data(object):
    """Same attribute surface as the reference data object (RankingPreprocess.py:17,43)."""

    def __init__(self, user_nums, item_nums, ui_train, ui_test):
        self.user_nums, self.item_nums = user_nums, item_nums
        self.ui_train, self.ui_test = ui_train, ui_test


def default_configs(**over):
    """CleverRec.properties [default] + conf/BPR.properties [parameters] as the flat str dict main.py:18-25 builds."""
    import configparser as cp
    conf = cp.ConfigParser()
    conf.read(os.path.join(REFERENCE_ROOT, "CleverRec.properties"), encoding="utf-8")
    configs = dict(conf.items("default"))
    rec = over.get("recommender", configs["recommender"])
    conf.read(os.path.join(REFERENCE_ROOT, "conf", rec + ".properties"), encoding="utf-8")
    configs.update(dict(conf.items("parameters")))
    configs.update({k: str(v) for k, v in over.items()})
    return configs


def preprocess(configs):
    ref = load()
    cfg = dict(configs)
    cfg["data.root_dir"] = os.path.join(REFERENCE_ROOT, "dataset")
    logger = logging.getLogger("oracle.refimport")
    return ref.RankingPreprocess(cfg, logger)


def make_driver(configs, data, fake_sess):
    """Build the reference's RankingRecommender without TF (SURVEY.md section 4 'Oracle-loop').

    ``fake_sess.run(fetches, feed_dict)`` receives the placeholder *names* below as keys."""
    ref = load()
    import math
    drv = object.__new__(ref.RankingRecommender)
    drv.sess, drv.data, drv.configs = fake_sess, data, configs
    drv.logger = logging.getLogger("oracle.refimport")
    drv.model = configs["recommender"]
    drv.epoches, drv.batch_size, drv.batch_size_t = int(configs["epoches"]), int(configs["batch_size"]), int(configs["test.batch_size"])
    drv.lr, drv.neg_samples = float(configs["lr"]), int(configs["test.neg_samples"])
    drv.fism_like, drv.cml_like = "fism_like" in configs, "cml_like" in configs
    drv.is_pairwise = configs["is_pairwise"]
    drv.T = int(configs["test.interval"])
    drv.topk = list(map(int, configs["topk"][1:-1].split(",")))
    drv.neg_ratio = int(configs["neg_ratio"])
    drv.test_users = list(data.ui_test.keys())
    drv.test_batches = math.ceil(len(drv.test_users) / drv.batch_size_t)
    for name in ("u_idx", "i_idx", "j_idx", "y", "u_neighbors_num", "neg_items", "batch_size_t_",
                 "u_nbrs_num", "i_nums", "train", "loss", "pre_scores"):
        setattr(drv, name, name)
    return drv
