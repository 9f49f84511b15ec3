"""Worker for tests/test_gpu2_sharded.py::test_bpr_model_class_under_torchrun (one process per GPU): the drop-in BPR class behind
Recommender / RankingRecommender runs the multi-GPU path when WORLD_SIZE > 1 -- same methods, same returns on every rank."""
import logging
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_split  # noqa: E402
from oracle import c_oracle as O  # noqa: E402
from oracle import ref_host as H  # noqa: E402

CFG = {'recommender': 'BPR', 'model_type': 'ranking', 'saved_dir': './saved_model', 'data.split_way': 'loo', 'test.neg_samples': '99',
       'test.batch_size': '256', 'test.interval': '1', 'topk': '[10,20]', 'epoches': '2', 'batch_size': '6144', 'embed_size': '64',
       'reg': '0.01', 'lr': '0.01', 'neg_ratio': '4', 'optimizer': 'Adam', 'is_pairwise': 'True', 'loss_func': 'bpr',
       'init_method': 'normal ', 'stddev': '0.01', 'seed': '5'}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if os.environ.get("CRB_SHARED_DEVICE", "0") == "1":   # both processes on device 0 (gloo control plane, CUDA IPC data path)
        local = 0
    torch.cuda.set_device(local)
    from cleverrec_b200.model.ranking.BPR import BPR
    ok = True
    for split, name in (("loo", "split_ml100k_loo.npz"), ("rs", "split_ml100k_rs.npz")):
        data = load_split(name)
        cfg = dict(CFG, **{'data.split_way': split, 'test.neg_samples': '99' if split == 'loo' else '0'})
        m = BPR(None, data, cfg, logging.getLogger("w%d" % rank))
        assert m.sharded and m.world == world
        m.build_model()
        l0 = m.train_model()
        hr_before = None
        for _ in range(3):
            l1 = m.train_model()
        both = [None] * world
        dist.all_gather_object(both, (l0, l1))
        if not (np.isfinite(l1) and l1 < l0 and all(b == both[0] for b in both)):
            print("LOSS", rank, l0, l1, both); ok = False
        HR, MRR, NDCG = m.test_model_loo() if split == "loo" else m.test_model_rs()
        if len(HR[0]) != len(m.test_users) or len(NDCG[1]) != len(m.test_users):
            print("LEN", rank, len(HR[0]), len(m.test_users)); ok = False
        # oracle: the reference's own evaluation loop on the gathered tables
        P, Q = m._shm.gather_P().cpu().numpy(), m._shm.gather_Q().cpu().numpy()
        if rank == 0:
            if split == "loo":
                scores = {u: O.score_pairs(0, P, Q, np.full(len(data.ui_test[u]), u), np.asarray(data.ui_test[u])) for u in m.test_users}
                want = H.eval_loo(m.test_users, data.ui_test, scores, 99, m.topk)
            else:
                users = np.asarray(m.test_users, dtype=np.int32)
                rows = O.score_pairs(0, P, Q, np.repeat(users, data.item_nums), np.tile(np.arange(data.item_nums, dtype=np.int32), users.shape[0]))
                want = H.eval_rs(m.test_users, data.ui_train, data.ui_test, rows.reshape(users.shape[0], data.item_nums), m.topk)
            for k in range(len(m.topk)):
                if not (HR[k] == want[0][k] and MRR[k] == want[1][k] and NDCG[k] == want[2][k]):
                    print("EVAL MISMATCH", split, k); ok = False
            if split == "loo" and not np.mean(HR[0]) > 0.3:
                print("QUALITY", np.mean(HR[0])); ok = False
        m._shm.close()
    # ---- GMF and MF behind the same interface under WORLD_SIZE > 1 (pointwise epochs over the sharded tables)
    import importlib
    for name, extra in (("GMF", {'loss_func': 'cross_entropy', 'init_method': 'xavier_uniform', 'embed_size': '32', 'reg': '0.001'}),
                        ("MF", {'loss_func': 'cross_entropy', 'embed_size': '32', 'reg': '0.001'})):
        data = load_split("split_ml100k_loo.npz")
        cfg = dict(CFG, recommender=name, is_pairwise='False', **extra)
        cls = getattr(importlib.import_module('cleverrec_b200.model.ranking.' + name), name)
        m = cls(None, data, cfg, logging.getLogger("w%d" % rank))
        assert m.sharded
        m.build_model()
        l0 = m.train_model()
        for _ in range(3):
            l1 = m.train_model()
        both = [None] * world
        dist.all_gather_object(both, (l0, l1))
        if not (np.isfinite(l1) and l1 < l0 and all(b == both[0] for b in both)):
            print("PW LOSS", name, rank, l0, l1, both); ok = False
        HR, MRR, NDCG = m.test_model_loo()
        P, Q = m._shm.gather_P().cpu().numpy(), m._shm.gather_Q().cpu().numpy()
        hv = m.h_gmf.cpu().numpy() if m.h_gmf is not None else None
        if rank == 0:
            kind = 1 if name == "GMF" else 0
            scores = {u: O.score_pairs(kind, P, Q, np.full(len(data.ui_test[u]), u), np.asarray(data.ui_test[u]), hv) for u in m.test_users}
            want = H.eval_loo(m.test_users, data.ui_test, scores, 99, m.topk)
            for k in range(len(m.topk)):
                if not (HR[k] == want[0][k] and MRR[k] == want[1][k] and NDCG[k] == want[2][k]):
                    print("PW EVAL MISMATCH", name, k); ok = False
            if not np.mean(HR[0]) > 0.2:
                print("PW QUALITY", name, np.mean(HR[0])); ok = False
        m._shm.close()
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    from cleverrec_b200.dist import all_reduce_dev
    all_reduce_dev(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_MODEL_OK" if flag.item() == 1.0 else "SHARDED_MODEL_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
