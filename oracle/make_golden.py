"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the GENUINE reference host code
(imported from /root/reference through oracle/refimport.py, TensorFlow/gensim stubbed) under fixed NumPy seeds.
Run in the build container only:  python -m oracle.make_golden      (the GPU box has no /root/reference)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refimport as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def flat(d):
    keys = np.asarray(list(d.keys()), dtype=np.int32)
    lens = np.asarray([len(d[k]) for k in d], dtype=np.int32)
    items = np.asarray([i for k in d for i in d[k]], dtype=np.int32)
    return keys, lens, items


def unflat(keys, lens, items):
    out, p = {}, 0
    for k, n in zip(keys.tolist(), lens.tolist()):
        out[k] = items[p:p + n].tolist()
        p += n
    return out


class FakeSession(object):
    """Returns seeded random scores for whatever the reference eval loop asks (shape taken from the feed)."""

    def __init__(self, seed, item_nums):
        self.rs = np.random.RandomState(seed)
        self.item_nums = item_nums

    def run(self, fetches, feed):
        if "i_idx" in feed:
            return self.rs.rand(len(feed["i_idx"])).astype(np.float32)
        return self.rs.rand(len(feed["u_idx"]), self.item_nums).astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = R.load()
    ml = {"data.dataset": "ml-100k", "data.file_name": "u.data", "data.sep": "\t", "data.format": "UIRT", "data.item_min": 0}

    # A/E: the reference's own preprocessing + split under a seed
    cfg_loo = R.default_configs(**dict(ml, **{"data.split_way": "loo", "test.neg_samples": 99}))
    np.random.seed(7)
    d_loo = R.preprocess(cfg_loo)
    np.savez_compressed(os.path.join(OUT, "split_ml100k_loo.npz"), user_nums=d_loo.user_nums, item_nums=d_loo.item_nums,
                        **{"train_" + k: v for k, v in zip(("keys", "lens", "items"), flat(d_loo.ui_train))},
                        **{"test_" + k: v for k, v in zip(("keys", "lens", "items"), flat(d_loo.ui_test))})
    cfg_rs = R.default_configs(**dict(ml, **{"data.split_way": "rs", "test.neg_samples": 0}))
    np.random.seed(11)
    d_rs = R.preprocess(cfg_rs)
    np.savez_compressed(os.path.join(OUT, "split_ml100k_rs.npz"), user_nums=d_rs.user_nums, item_nums=d_rs.item_nums,
                        **{"train_" + k: v for k, v in zip(("keys", "lens", "items"), flat(d_rs.ui_train))},
                        **{"test_" + k: v for k, v in zip(("keys", "lens", "items"), flat(d_rs.ui_test))})

    # B: sampler outputs on a 120-user slice of A (dict order preserved)
    sub_users = list(d_loo.ui_train.keys())[:120]
    sub = R.Data(d_loo.user_nums, d_loo.item_nums, {u: d_loo.ui_train[u] for u in sub_users}, {})
    np.random.seed(3)
    pw = ref.pairwise_ranking_sampler(sub, 4, 6144, fism_like=True)
    pt = ref.pointwise_ranking_sampler(sub, 2, 1000)
    cm = ref.ranking_sampler_cml(sub, 5, 512)
    np.savez_compressed(os.path.join(OUT, "sampler_seed3.npz"), n_sub_users=120,
                        pw_batches=pw[0], pw_u=pw[1].astype(np.int32), pw_i=pw[2].astype(np.int32), pw_j=pw[3].astype(np.int32),
                        pw_nbr=pw[4].astype(np.int32),
                        pt_batches=pt[0], pt_u=pt[1].astype(np.int32), pt_i=pt[2].astype(np.int32), pt_y=pt[3].astype(np.float32),
                        cm_batches=cm[0], cm_u=cm[1].astype(np.int32), cm_i=cm[2].astype(np.int32), cm_neg=cm[3].astype(np.int32))

    # C: metric known answers
    rs = np.random.RandomState(5)
    cases = []
    for _ in range(200):
        K = int(rs.choice([1, 5, 10, 20]))
        rec = rs.permutation(60)[:K].astype(np.int64)
        real = rs.permutation(60)[:rs.randint(1, 8)].tolist()
        h, m, n = ref.cal_ranking_metrics(real, rec, K)
        cases.append((K, rec, real, h, m, n))
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), K=np.asarray([c[0] for c in cases]),
                        rec=np.asarray([np.pad(c[1], (0, 20 - len(c[1])), constant_values=-1) for c in cases]),
                        real=np.asarray([np.pad(c[2], (0, 8 - len(c[2])), constant_values=-1) for c in cases]),
                        out=np.asarray([[c[3], c[4], c[5]] for c in cases], dtype=np.float64))

    # D: the reference's eval loops driven by seeded random scores
    cfg = dict(cfg_loo)
    drv = R.make_driver(cfg, d_loo, FakeSession(21, d_loo.item_nums))
    HR, MRR, NDCG = drv.test_model_loo()
    loo = np.asarray([[HR[k], MRR[k], NDCG[k]] for k in range(len(drv.topk))], dtype=np.float64)
    drv = R.make_driver(dict(cfg_rs), d_rs, FakeSession(22, d_rs.item_nums))
    HR, MRR, NDCG = drv.test_model_rs()
    rs_out = np.asarray([[HR[k], MRR[k], NDCG[k]] for k in range(len(drv.topk))], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "eval_loops.npz"), loo=loo, rs=rs_out, topk=np.asarray(drv.topk),
                        batch_size_t=drv.batch_size_t)
    # F: SBPR on Ciao (the dataset with a trust file): the genuine preprocessing with social_file, get_SPu and ranking_sampler_sbpr
    # on the first 150 training users (dict order; friends outside the slice are skipped by the reference's own membership tests)
    cfg_sb = R.default_configs(recommender="SBPR", **{"data.dataset": "Ciao", "data.file_name": "ratings.csv", "data.sep": ",", "data.format": "UI",
                                                       "data.split_way": "rs", "test.neg_samples": 0})
    np.random.seed(5)
    d_sb = R.preprocess(cfg_sb)
    sub_users = list(d_sb.ui_train.keys())[:150]
    sub = R.Data(d_sb.user_nums, d_sb.item_nums, {u: d_sb.ui_train[u] for u in sub_users}, {})
    sub.user_friends = {u: d_sb.user_friends[u] for u in sub_users if u in d_sb.user_friends}
    SPu = ref.get_SPu(sub)
    np.random.seed(3)
    sb = ref.ranking_sampler_sbpr(sub, SPu, 2, 4096)
    np.savez_compressed(os.path.join(OUT, "sbpr_ciao.npz"), user_nums=sub.user_nums, item_nums=sub.item_nums,
                        n_friend_users=len(d_sb.user_friends), n_friend_pairs=sum(len(v) for v in d_sb.user_friends.values()),
                        **{"train_" + k: v for k, v in zip(("keys", "lens", "items"), flat(sub.ui_train))},
                        **{"friends_" + k: v for k, v in zip(("keys", "lens", "items"), flat(sub.user_friends))},
                        **{"spu_" + k: v for k, v in zip(("keys", "lens", "items"), flat(SPu))},
                        sb_batches=sb[0], sb_u=sb[1].astype(np.int32), sb_i=sb[2].astype(np.int32), sb_k=sb[3].astype(np.int32),
                        sb_j=sb[4].astype(np.int32), sb_suk=sb[5].astype(np.int32))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
