"""The 4096-user full-rank top-20 call at 2M items, d = 128 (the reference's test.batch_size loop at the S-large catalogue), alone:
    python scripts/eval_small_call.py [n_users] [cold]     # prints ms per call (CUDA events, 8 calls after 2 warm ones); cold = the bf16
                                                           # item table is dropped before every call (evaluation after a training epoch)
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python scripts/eval_small_call.py
gives the per-kernel split of one call (prep / score_tc / rescore)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_history_device  # noqa: E402
from cleverrec_b200.engine import Engine  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    cold = len(sys.argv) > 2 and sys.argv[2] == "cold"
    items, dim = 2_000_000, 128
    dev = torch.device("cuda", 0)
    eng = Engine(0)
    pu, pi, rowptr = build_history_device(torch, dev, n, items, 100, seed=3)
    eng.set_history_arrays(n, items, pu, pi, rowptr, pi)
    g = torch.Generator(device=dev).manual_seed(0)
    P = torch.randn(n, dim, device=dev, generator=g) * 0.01
    Q = torch.randn(items, dim, device=dev, generator=g) * 0.01
    users = torch.arange(n, device=dev, dtype=torch.int32)
    for _ in range(2):
        eng.score_topk(0, P, Q, users, 20)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(8):
        if cold:
            eng.invalidate_eval_cache()
        eng.score_topk(0, P, Q, users, 20)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 8
    print(("cold " if cold else "") + "users %d ms/call %.3f  users/s %.3e  frac of 1375.1 TF/s %.3f  stats %s" % (n, ms, n / ms * 1e3, 2.0 * n * items * dim / (ms * 1e-3) / 1375.1e12,
                                                                                     eng.score_topk_stats()))


if __name__ == "__main__":
    main()
