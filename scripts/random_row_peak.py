"""What does HBM3e give for RANDOM 512-byte rows?  The roofline denominator in MEASURED_PEAKS.json is a contiguous copy; the training
step gathers and scatters whole embedding rows at random addresses.  This measures library gathers / scatters of the same shape
(torch.index_select / index_copy_ on a [10M, 128] fp32 table, 2^22 distinct random rows) as context for the K3 / K4 fractions."""
import json
import torch

dev = torch.device("cuda", 0)
rows, dim, n = 10_000_000, 128, 1 << 22
g = torch.Generator(device=dev).manual_seed(0)
table = torch.randn(rows, dim, device=dev, generator=g)
idx = torch.randperm(rows, device=dev, generator=g)[:n].contiguous()
src = torch.randn(n, dim, device=dev, generator=g)
out = torch.empty(n, dim, device=dev)
big = torch.empty(1 << 28, device=dev)   # 1 GiB: flushes the 126 MB L2 between repetitions


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        big.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


res = {}
ms = timed(lambda: torch.index_select(table, 0, idx, out=out))
res["gather_random_rows_write_contiguous"] = {"ms": ms, "GBs": 2 * n * dim * 4 / ms / 1e6}
ms = timed(lambda: table.index_copy_(0, idx, src))
res["read_contiguous_scatter_random_rows"] = {"ms": ms, "GBs": 2 * n * dim * 4 / ms / 1e6}
ms = timed(lambda: out.copy_(src))
res["contiguous_copy_same_bytes"] = {"ms": ms, "GBs": 2 * n * dim * 4 / ms / 1e6}
print(json.dumps(res))
