"""One launch of every large-shape kernel inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off --set full`
(run under gpurun, after a plain run exited 0):  K1 sample_kernel, K2 assign_kernel, K3 bpr_step_kernel, K4 dup_reduce_kernel,
the multi-GPU step's kernels through the world = 1 path (shard_resolve / item_fetch / shard_step / dup_reduce<SHARD> / inbox_apply),
loo_topk_kernel, prep / score_tc / rescore.  Shape: bench.py's `medium` workload by default (2M users x 500K items, d=128,
B=2^18: large against the 126 MB L2, small enough for ncu's save/restore replays); WORKLOAD=s_large for the headline shape."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from cleverrec_b200.dist import ShardedBPR  # noqa: E402
from cleverrec_b200.engine import Engine, Optimizer, Table  # noqa: E402


def main():
    w = dict(B.WORKLOADS[os.environ.get("WORKLOAD", "medium")])
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29547")
    dist.init_process_group("gloo", rank=0, world_size=1)
    eng = Engine(0)
    users, items, dim, batch, R = w["users"], w["items"], w["dim"], w["batch"], w["neg_ratio"]
    pu, pi, rowptr = B.build_history_device(torch, dev, users, items, w["mean_hist"], seed=1234)
    eng.set_history_arrays(users, items, pu, pi, rowptr, pi)
    g = torch.Generator(device=dev).manual_seed(0)
    P = Table(torch.randn(users, dim, device=dev, generator=g) * 0.01, "Adam", "tf1")
    Q = Table(torch.randn(items, dim, device=dev, generator=g) * 0.01, "Adam", "tf1")
    opt = Optimizer("Adam", 1e-3)
    sh = ShardedBPR(eng, users, items, dim, "Adam", 1e-3, "tf1", batch, seed=0)
    losses = torch.zeros(8, dtype=torch.float64, device=dev)
    n_loo, n_cand = 16384, 1001
    lu = torch.arange(n_loo, device=dev, dtype=torch.int32)
    li = torch.randint(0, items, (n_loo * n_cand,), device=dev, dtype=torch.int32)
    seg = torch.arange(n_loo + 1, device=dev, dtype=torch.int64) * n_cand
    eu = torch.arange(32768, device=dev, dtype=torch.int32)

    def once(k):
        eng.train_epoch_bpr(P, Q, opt, 7, 0, k * batch, batch, 1, R, 0.01, losses[k:k + 1])
        sh.run_steps(1, 0.01, neg_ratio=R, seed=7, epoch=0, first=k * batch, batch=batch, loss_out=losses[k:k + 1])
        eng.score_pairs_topk(0, P.w, Q.w, lu, li, seg, 20)
        eng.score_topk(0, P.w, Q.w, eu, 20)
    for k in range(3):
        once(k)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    once(3)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sh.check()
    print("PROFILE_BIG_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
