"""Debug helper: tensor-core vs exact full-rank top-K on the test_gpu_eval input; prints where they differ."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import synthetic_data
from cleverrec_b200.engine import Engine
kind = 0
d = synthetic_data(150, 3000, 60, seed=kind)
eng = Engine(0)
eng.set_history(d.ui_train, d.user_nums, d.item_nums)
rs = np.random.RandomState(kind)
dim = 64
P, Q = (rs.randn(d.user_nums, dim) * 0.1).astype(np.float32), (rs.randn(d.item_nums, dim) * 0.1).astype(np.float32)
Q[100] = Q[200]
users = np.arange(d.user_nums, dtype=np.int32)
Pd, Qd = torch.tensor(P).cuda(), torch.tensor(Q).cuda()
a = eng.score_topk(kind, Pd, Qd, users, 20, exact=True)
b = eng.score_topk(kind, Pd, Qd, users, 20, exact=False)
print("stats", eng.score_topk_stats())
for u in range(d.user_nums):
    if not np.array_equal(a[u], b[u]):
        extra = [int(x) for x in b[u] if x not in a[u]]
        miss = [int(x) for x in a[u] if x not in b[u]]
        seen = sorted(d.ui_train.get(u, []))
        print("user", u, "extra", extra, "in history:", [x in seen for x in extra], "missing", miss, "col%128", [x % 128 for x in extra])
        print("   seen near:", [s for s in seen if any(abs(s - x) < 200 for x in extra)])
