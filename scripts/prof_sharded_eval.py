"""Section timing of ShardedEval.topk (debugging aid; run under torch.distributed.run)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cleverrec_b200.engine import Engine
from cleverrec_b200 import dist as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(local)
users, items, dim = 2_000_000, 2_000_000, 128
lo, hi = D.user_range(users, rank, world)
pu, pi, rowptr = bench.build_history_device(torch, dev, hi - lo, items, 50, seed=rank)
eng.set_history_arrays(hi - lo, items, pu, pi, rowptr, pi)
m = D.ShardedBPR(eng, users, items, dim, "Adam", 1e-3, "tf1", 1 << 18, seed=rank)
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0
t0 = time.perf_counter(); ev = D.ShardedEval(m, rowptr, pi); tick("setup", t0)
orig = ev.eng.score_topk
def timed(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = orig(*a, **k); tick("score_topk", t0); return r
ev.eng.score_topk = timed
for it in range(2):
    T.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev.topk(20, batch_users=65536, limit=65536)
    tick("total", t0)
    if rank == 0:
        print(it, {k: round(v * 1e3, 1) for k, v in T.items()}, ev.eng.score_topk_stats())
dist.destroy_process_group()
