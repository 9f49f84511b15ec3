"""Times score_pairs and topk_segments separately at the bench's loo-evaluation shape (experiments)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleverrec_b200.engine import Engine
eng = Engine(0)
dev = torch.device("cuda", 0)
U, I, d, n_users, n_cand = 1_000_000, 2_000_000, 128, 65536, 1001
g = torch.Generator(device=dev).manual_seed(0)
P = torch.randn(U, d, device=dev, generator=g) * 0.1
Q = torch.randn(I, d, device=dev, generator=g) * 0.1
lu = torch.arange(n_users, device=dev, dtype=torch.int32).repeat_interleave(n_cand)
li = torch.randint(0, I, (n_users * n_cand,), device=dev, generator=g, dtype=torch.int32)
seg = torch.arange(n_users + 1, device=dev, dtype=torch.int64) * n_cand
out = torch.empty(n_users * n_cand, dtype=torch.float32, device=dev)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("score_pairs ms", t(lambda: eng.score_pairs(0, P, Q, lu, li, out=out)))
print("topk_segments ms", t(lambda: eng.topk_segments(out, seg, 20)))
