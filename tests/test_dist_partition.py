"""CPU (gloo, world_size 2): the partition logic of the multi-GPU path (cleverrec_b200/dist.py) -- user ranges, item ownership,
history sharding -- is a consistent cover, checked across two real processes."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT, synthetic_data

WORKER = r'''
import os, sys, pickle
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
from conftest import synthetic_data
from cleverrec_b200.dist import user_range, shard_history, shard_rows, item_owner, item_local_row
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
d = synthetic_data(101, 57, 6, seed=1)
mine, n_local = shard_history(d.ui_train, d.user_nums, rank, world)
lo, hi = user_range(d.user_nums, rank, world)
assert n_local == hi - lo and all(0 <= u < n_local for u in mine)
rows = shard_rows(d.item_nums, rank, world)
owned = [i for i in range(d.item_nums) if item_owner(i, world) == rank]
assert rows == len(owned) and sorted(item_local_row(i, world) for i in owned) == list(range(rows))
out = [None] * world
dist.all_gather_object(out, {"users": sorted(u + lo for u in mine), "n_pos": sum(len(v) for v in mine.values()), "rows": rows, "range": (lo, hi)})
if rank == 0:
    users = sum((o["users"] for o in out), [])
    assert users == sorted(d.ui_train.keys())                      # every user exactly once
    assert sum(o["n_pos"] for o in out) == sum(len(v) for v in d.ui_train.values())
    assert sum(o["rows"] for o in out) == d.item_nums
    assert out[0]["range"][0] == 0 and out[-1]["range"][1] == d.user_nums and all(out[k]["range"][1] == out[k + 1]["range"][0] for k in range(world - 1))
    print("PARTITION_OK")
dist.destroy_process_group()
'''


def test_partition_cover_two_processes(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % (ROOT, ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29519",
           str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "PARTITION_OK" in r.stdout, r.stdout[-3000:]


def test_partition_helpers_single_process():
    from cleverrec_b200.dist import item_local_row, item_owner, shard_rows, user_range
    for n, w in ((10, 3), (7, 8), (1_000_003, 8)):
        assert sum(shard_rows(n, r, w) for r in range(w)) == n
        assert [user_range(n, r, w)[0] for r in range(w)] + [n] == [0] + [user_range(n, r, w)[1] for r in range(w)]
    items = np.arange(1000)
    assert np.array_equal(item_owner(items, 8) + 8 * item_local_row(items, 8), items)


EVAL_WORKER = r'''
import os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
from conftest import synthetic_data
from cleverrec_b200.dist import user_range, shard_history, transpose_history, merge_topk, shard_rows
from cleverrec_b200.engine import history_from_dict
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
d = synthetic_data(64, 45, 7, seed=2)
mine, n_local = shard_history(d.ui_train, d.user_nums, rank, world)
lo, hi = user_range(d.user_nums, rank, world)
_, _, rowptr, cols = history_from_dict(mine, n_local)
tp, tc = transpose_history(torch.from_numpy(rowptr), torch.from_numpy(cols), lo, d.user_nums, world, rank)
# for every user: the local rows of its seen items that live on this shard
for u in range(d.user_nums):
    want = sorted(i // world for i in set(d.ui_train.get(u, [])) if i %% world == rank)
    assert tc[tp[u]:tp[u + 1]].tolist() == want, (u, want)
# merge of per-shard exact top-K == global top-K (ties by id)
rs = np.random.RandomState(0)
scores = np.round(rs.randn(10, d.item_nums) * 2).astype(np.float32) / 2   # heavy ties
K = 6
mine_items = np.arange(rank, d.item_nums, world)
loc = torch.from_numpy(scores[:, mine_items])
order = torch.argsort(-loc, dim=1, stable=True)[:, :K]
ids = torch.from_numpy(mine_items)[order].to(torch.int32)
sc = torch.gather(loc, 1, order)
all_ids, all_sc = [torch.empty_like(ids) for _ in range(world)], [torch.empty_like(sc) for _ in range(world)]
dist.all_gather(all_ids, ids); dist.all_gather(all_sc, sc)
got, _ = merge_topk(all_ids, all_sc, K)
want = np.argsort(-scores, axis=1, kind="stable")[:, :K]
assert np.array_equal(got.numpy(), want), (got, want)
if rank == 0:
    print("EVAL_PARTITION_OK")
dist.destroy_process_group()
'''


def test_transposed_history_and_topk_merge_two_processes(tmp_path):
    script = tmp_path / "w2.py"
    script.write_text(EVAL_WORKER % (ROOT, ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29521",
           str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "EVAL_PARTITION_OK" in r.stdout, r.stdout[-3000:]
