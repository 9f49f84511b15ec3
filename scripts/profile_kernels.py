"""Short workload that launches EVERY kernel behind the model classes a few times, for ncu (run under gpurun):
    python scripts/profile_kernels.py                                      # plain run first (must exit 0)
    ncu --profile-from-start off --metrics <list> --csv ... python scripts/profile_kernels.py
Every model class of scripts/bench_models.py at the item / batch / dimension shapes of BASELINE.json configs[0..3], with the user
count cut so that an epoch is a handful of steps (a step's kernels see the same batch, tables and dimensions as in the full-size
run; the tables of these configurations are L2-resident either way).  One warm epoch + evaluation outside the profiled range, then
one epoch + one evaluation inside cudaProfilerStart/Stop."""
import importlib
import logging
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench_models as BM  # noqa: E402


def main():
    log = logging.getLogger("profile_kernels")
    only = set(sys.argv[1:])
    for name, shape, conf in BM.MODELS:
        if only and name not in only:
            continue
        small = dict(shape, users=max(48, shape["users"] // (60 if name != "NAIS_single" else 150)))
        data = BM.make_data(small, 1, friends=name == "SBPR")
        for split in (("loo", "99"), ("rs", "0")):
            cfg = dict(BM.BASE, recommender=name)
            cfg.update(conf)
            cfg.update({"data.split_way": split[0], "test.neg_samples": split[1]})
            cls = getattr(importlib.import_module("cleverrec_b200.model.ranking." + name), name)
            m = cls(None, data, cfg, log)
            m.build_model()
            ev = m.test_model_loo if split[0] == "loo" else m.test_model_rs
            m.train_model()
            ev()
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            if split[0] == "loo":
                m.train_model()
            ev()
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            m.engine.close()
        sys.stderr.write("%s done\n" % name)
    print("PROFILE_KERNELS_OK")


if __name__ == "__main__":
    main()
