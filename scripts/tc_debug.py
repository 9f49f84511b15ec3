import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleverrec_b200.engine import Engine
eng = Engine(0)
U, I, d = 37888, 2_000_000, 128
g = torch.Generator(device="cuda").manual_seed(0)
P = torch.randn(U, d, device="cuda", generator=g) * 0.01
Q = torch.randn(I, d, device="cuda", generator=g) * 0.01
rp = torch.zeros(U + 1, dtype=torch.int64, device="cuda")
z = torch.zeros(1, dtype=torch.int32, device="cuda")
eng.set_history_arrays(U, I, z[:0], z[:0], rp, z)
users = torch.arange(U, dtype=torch.int32, device="cuda")
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.score_topk(0, P, Q, users, 20)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("CRB_TC_DEBUG=%s  %.2f ms  (%.0f TFLOP/s)" % (os.environ.get("CRB_TC_DEBUG", "0"), dt * 1e3, 2.0 * U * I * d / dt / 1e12))
