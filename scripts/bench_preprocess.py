"""Device preprocessing (csrc/preprocess.cu + csrc/history.cu) on a synthetic interaction log resident on the device:
filter (user_min = item_min = 5) + re-index + leave-one-out split by time + native history build + 100 evaluation negatives per
test user.  python scripts/bench_preprocess.py [rows]   (default 1e8; run under gpurun)"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleverrec_b200.engine import Engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
eng = Engine(0)
dev = eng.device
g = torch.Generator(device=dev).manual_seed(0)
users, items = max(1000, n // 100), max(1000, n // 500)
raw_u = torch.randint(0, users, (n,), device=dev, generator=g, dtype=torch.int64) * 3 + 7
raw_i = torch.randint(0, items, (n,), device=dev, generator=g, dtype=torch.int64) * 5 + 1
t = torch.randint(0, 1 << 30, (n,), device=dev, generator=g, dtype=torch.int64)
out = {"rows": n}
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = eng.prep_filter_reindex(raw_u, raw_i, 5, 5)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    perm, is_test = eng.prep_split_loo(res["u"], res["n_users"], t[res["row"]])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    tr = perm[~is_test]
    eng.build_history(res["u"][tr], res["i"][tr], res["n_users"], res["n_items"])
    torch.cuda.synchronize(); t3 = time.perf_counter()
    te_users = res["u"][perm[is_test]]
    negs = eng.prep_eval_negatives(0, te_users, 100)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    out = {"rows": n, "kept": int(res["u"].numel()), "users": res["n_users"], "items": res["n_items"], "test_users": int(te_users.numel()),
           "filter_reindex_s": t1 - t0, "split_loo_s": t2 - t1, "build_history_s": t3 - t2, "eval_negatives_s": t4 - t3, "total_s": t4 - t0,
           "rows_per_s": n / (t4 - t0)}
print(json.dumps(out))
