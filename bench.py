#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on synthetic data of the named shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload s_large|ml1m|tiny]

metric  : BPR train triplets/sec (one triplet = sampled on the device, scored, back-propagated and applied:
          K1 sampler+count, K2 assign, K3 fused step, K4 duplicate reduce, K5 loss), whole job over all N GPUs.
workload: s_large = BASELINE configs[4], the configuration the metric is quoted on: 10M users x 2M items, d=128,
          1e9 interactions (mean history 100, uniform item popularity), B = 2^20 triplets / step / GPU, neg_ratio 4,
          Adam with tf.train.AdamOptimizer semantics (CRB_ADAM_TF1).  It fits one GPU (about 31 GB).
value   : device-timed throughput, inputs resident in HBM.   e2e: the same step driven through the C ABI with HOST
          index buffers (pinned) in and the HOST loss out every step -- the reference's feed_dict / fetch boundary.
One JSON line on stdout (rank 0).  `--impl reference` times the oracle port of the reference's CPU path."""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: users, items, dim, mean history, batch per GPU, neg_ratio
    "s_large": dict(users=10_000_000, items=2_000_000, dim=128, mean_hist=100, batch=1 << 20, neg_ratio=4),
    "medium": dict(users=2_000_000, items=500_000, dim=128, mean_hist=50, batch=1 << 18, neg_ratio=4),  # profiling size (ncu replays)
    "ml1m": dict(users=6040, items=3706, dim=64, mean_hist=165, batch=6144, neg_ratio=4),
    "tiny": dict(users=20000, items=5000, dim=64, mean_hist=30, batch=1 << 14, neg_ratio=4),
}
METRIC, UNIT = "bpr_train_triplets_per_sec", "triplets/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def nvlink_counters(index):
    """Sum of the per-link NVLink data counters of one GPU (`nvidia-smi nvlink -gt d`), bytes: (tx, rx) or None.  Read before and
    after the timed steps at N > 1: measured wire payload per step, next to the formula."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                             timeout=10).stdout
        tx = rx = 0
        seen = False
        for line in out.splitlines():
            if "Data Tx:" in line or "Data Rx:" in line:
                val = int(line.split(":")[-1].strip().split()[0]) * 1024
                seen = True
                if "Tx" in line:
                    tx += val
                else:
                    rx += val
        return (tx, rx) if seen else None
    except Exception:
        return None


# --------------------------------------------------------------------------------------------- synthetic data (device)
def build_history_device(torch, dev, users, items, mean_hist, seed, user_lo=0, user_hi=None, zipf=False):
    """Per-user sorted unique histories for users [user_lo, user_hi) drawn on the device, chunk by chunk.
    Returns (pos_user int32 [global ids], pos_item int32, rowptr int64 over ALL `users` rows)."""
    user_hi = users if user_hi is None else user_hi
    g = torch.Generator(device=dev).manual_seed(seed)
    pu, pi, counts = [], [], []
    chunk = max(1, min(user_hi - user_lo, (1 << 27) // max(1, mean_hist)))
    for a in range(user_lo, user_hi, chunk):
        b = min(user_hi, a + chunk)
        # history length: 1 + Poisson-like (sum of uniforms) around mean_hist, clipped to the catalogue
        lens = torch.clamp((torch.rand(b - a, device=dev, generator=g) * 2 * (mean_hist - 1)).long() + 1, max=items // 2)
        tot = int(lens.sum().item())
        uid = torch.repeat_interleave(torch.arange(a, b, device=dev, dtype=torch.int64), lens)
        if zipf:   # Zipf(1.0) item popularity (SURVEY 8d 'hub contention' variant): rank = floor(I^U), item 0 is the hub
            it = torch.clamp(torch.exp(torch.rand(tot, device=dev, generator=g, dtype=torch.float64) * math.log(items)).long() - 1, 0, items - 1)
        else:
            it = torch.randint(0, items, (tot,), device=dev, generator=g, dtype=torch.int64)
        key = torch.unique(uid * items + it)  # sorted, duplicates dropped
        del uid, it
        ku = torch.div(key, items, rounding_mode="floor")
        pu.append(ku.to(torch.int32))
        pi.append((key - ku * items).to(torch.int32))
        counts.append(torch.bincount(ku - a, minlength=b - a))
        del key, ku
    cnt = torch.zeros(users, dtype=torch.int64, device=dev)
    cnt[user_lo:user_hi] = torch.cat(counts)
    rowptr = torch.zeros(users + 1, dtype=torch.int64, device=dev)
    torch.cumsum(cnt, 0, out=rowptr[1:])
    return torch.cat(pu), torch.cat(pi), rowptr


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def workload_config(args, w, world):
    """The `config` object of the JSON line: the workload and nothing else, a pure function of the command line and the world size, so
    that the GPU arm and `--impl reference` print the SAME object for the same launch.  What a run measures about itself (setup time,
    the exact number of synthetic interactions, what the host could hold) goes into `run_info`."""
    users, items, dim = w["users"], w["items"], w["dim"]
    return {"workload": args.workload, "users": users, "items": items, "dim": dim, "mean_history": w["mean_hist"], "batch_per_gpu": w["batch"],
            "neg_ratio": w["neg_ratio"], "item_popularity": args.item_popularity, "optimizer": args.optimizer,
            "adam_mode": args.adam_mode if args.optimizer == "Adam" else None, "reg": 0.01,
            "l2_policy": "inputs larger than L2 (tables %.1f GB vs 126 MB L2)" % ((users + items) * dim * 4 / 1e9),
            "parallelism": ("single GPU" if world == 1 else "%d ranks: users partitioned, item table row-sharded (item %% N), rows and gradients "
                            "over NVLink peer memory (per-rank de-duplicated), flag barriers in peer memory, no data-path collective" % world)}


def cpu_reference_steps(w, n_steps, batch, threads, seed=0):
    """The reference's CPU path restated (oracle/): Python-loop sampler with np.random (utils/sampler.py:46-74) +
    one TF-1 graph step per batch in torch-CPU fp32 (BPR.py:31-44, Adam).  Same configuration as the GPU arm: full-size tables
    (users x d and items x d, when host memory allows) and the GPU arm's batch; the bounded part is the epoch -- the sampler walks
    only as many users (random rows of the user table) as fill ONE batch per step, not all of them.
    Returns (seconds per step list, triplets per step, users sampled per step, user-table rows)."""
    import torch
    from oracle import ref_host as H
    from oracle import tf1_restatement as T
    torch.set_num_threads(threads)
    rs = np.random.RandomState(seed)
    np.random.seed(seed)
    items, dim, R = w["items"], w["dim"], w["neg_ratio"]
    n_sample = max(8, int(math.ceil(1.08 * batch / (R * w["mean_hist"]))))  # users whose full epoch is about one batch
    table_rows = w["users"]
    try:
        import psutil
        need = 3.2 * 4.0 * dim * (w["users"] + items)     # tables + Adam m, v (+ slack)
        if psutil.virtual_memory().available < need:
            table_rows = n_sample
    except Exception:
        table_rows = n_sample
    n_sample = min(n_sample, table_rows)

    class D(object):
        pass
    data = D()
    data.item_nums, data.user_nums = items, table_rows
    params = {"P": torch.empty(table_rows, dim).normal_(0, 0.01, generator=torch.Generator().manual_seed(seed)),
              "Q": torch.empty(items, dim).normal_(0, 0.01, generator=torch.Generator().manual_seed(seed + 1))}
    opt = T.TF1Optimizer("Adam", 1e-3, adam_mode="lazy")  # row-sparse apply: generous to the CPU baseline
    times, done = [], 0
    for _ in range(n_steps):
        rows = np.sort(rs.choice(table_rows, n_sample, replace=False)) if table_rows > n_sample else np.arange(n_sample)
        ui_train = {}
        for u in rows.tolist():   # synthetic histories of the workload's law (not timed: the reference gets them from its preprocessing)
            n = int(min(items // 2, 1 + rs.randint(0, 2 * (w["mean_hist"] - 1) + 1)))
            ui_train[u] = np.unique(rs.randint(0, items, n)).tolist()
        data.ui_train = ui_train
        t0 = time.perf_counter()
        out = H.pairwise_ranking_sampler(data, R, batch)
        n = min(batch, out[1].shape[0])
        b = {"u": torch.from_numpy(out[1][:n]), "i": torch.from_numpy(out[2][:n]), "j": torch.from_numpy(out[3][:n])}
        T.bpr_step_rowsparse(params, b, 0.01, opt)
        times.append(time.perf_counter() - t0)
        done = n
    return times, done, n_sample, table_rows


def cpu_sample_text(done, n_sample, table_rows, w):
    return ("%d triplets/step: reference-algorithm Python sampler (np.random, utils/sampler.py:46-74) over %d users per step (random rows of a "
            "%d-row user table) + restated TF-1 BPR/Adam step in torch-CPU fp32 (row-sparse apply), %d-item x d=%d item table"
            % (done, n_sample, table_rows, w["items"], w["dim"]))


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = args.ref_batch or w["batch"]   # the GPU arm's per-GPU batch (2^20 at s_large: ~6 s per step on the host cores)
    times, done, n_sample, table_rows = cpu_reference_steps(w, args.warmup + args.steps, batch, threads)
    t = times[args.warmup:]
    ms = 1000.0 * sum(t) / len(t)
    value = done / (ms / 1000.0)
    sample = cpu_sample_text(done, n_sample, table_rows, w)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, w, int(os.environ.get("WORLD_SIZE", "1"))),
            "run_info": {"triplets_per_step": done, "user_table_rows_on_host": table_rows, "sampler_in_timed_region": True,
                         "apply": "row-sparse Adam (touched rows only): generous to the CPU, TF-1's dense moment update would touch every row",
                         "histories": "uniform items (the sampler's cost does not depend on the popularity law)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args, w):
    import torch
    import torch.distributed as dist
    from cleverrec_b200.engine import Engine, Optimizer, Table

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    elif args.sharded:  # experiment: the multi-GPU code path on one GPU (every "peer" is local memory)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    eng = Engine(local)
    users, items, dim, B, R = w["users"], w["items"], w["dim"], w["batch"], w["neg_ratio"]
    reg, opt_kind, adam_mode = 0.01, args.optimizer, args.adam_mode

    # ---- data + tables.  Users (rows of P, histories, sampling) are partitioned across ranks; at N > 1 the item table is
    # row-sharded (owner = item % N) and read / updated over NVLink peer memory (cleverrec_b200/dist.py) ----
    u_lo, u_hi = (users * rank) // world, (users * (rank + 1)) // world
    t_setup = time.time()
    pu, pi, rowptr = build_history_device(torch, dev, u_hi - u_lo, items, w["mean_hist"], seed=1234 + rank, zipf=args.item_popularity == "zipf")
    eng.set_history_arrays(u_hi - u_lo, items, pu, pi, rowptr, pi)
    n_pos = int(pu.numel())
    g = torch.Generator(device=dev).manual_seed(rank)
    sharded = None
    if world == 1 and not args.sharded:
        P = Table(torch.randn(users, dim, device=dev, generator=g) * 0.01, opt_kind, adam_mode)
        Q = Table(torch.randn(items, dim, device=dev, generator=g) * 0.01, opt_kind, adam_mode)
        opt = Optimizer(opt_kind, 1e-3, adam_mode=adam_mode)
    else:
        from cleverrec_b200.dist import ShardedBPR
        sharded = ShardedBPR(eng, users, items, dim, opt_kind, 1e-3, adam_mode, B, seed=rank)
        P, opt = sharded.P, sharded.opt
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup
    rows = eng.epoch_rows(R, "pairwise")
    steps_total = args.warmup + args.steps
    assert rows >= B, "workload smaller than one batch"
    if steps_total * B > rows:
        raise SystemExit("steps*batch exceeds one epoch of the synthetic workload")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    phase_ms = {} if args.phases else None

    def run_steps(first_step, n, losses, phases=None):
        if sharded is None:
            eng.train_epoch_bpr(P, Q, opt, 7, 0, first_step * B, B, n, R, reg, losses)
        else:
            sharded.run_steps(n, reg, neg_ratio=R, seed=7, epoch=0, first=first_step * B, batch=B, loss_out=losses, phase_ms=phases)

    # ---- device-resident run: value ----
    losses = torch.zeros(steps_total, dtype=torch.float64, device=dev)
    run_steps(0, args.warmup, losses)  # warm-up steps
    clocks = ClockSampler(local)
    clocks.start()
    eng.profile(True)
    eng.profile_read()
    l0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nv0 = nvlink_counters(local) if world > 1 and rank == 0 else None   # before the barrier: the other ranks must not wait for it inside the timed region
    barrier()
    ev0.record()
    run_steps(args.warmup, args.steps, losses[args.warmup:], phase_ms)
    ev1.record()
    barrier()
    nv1 = nvlink_counters(local) if world > 1 and rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    k3_ms, k3_n = eng.profile_read()
    other_kernels = {name: eng.profile_read(tag) for tag, name in ((1, "item_fetch_kernel"), (2, "dup_reduce+dup_final"), (3, "inbox_apply_kernel"))}
    eng.profile(False)
    launches = eng.launches - l0
    clk = clocks.summary()
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1000.0)
    loss_last = float(losses[-1].item())

    # ---- e2e: host feed in (pinned), host loss out, every step, through the C ABI ----
    n_e2e = args.steps
    first = steps_total * B
    if first + (n_e2e + 1) * B > rows:
        first = 0
    feeds = []
    for k in range(n_e2e + 1):
        u, i, j = eng.sample_pairwise(7, 0, first + k * B, B, R)
        feeds.append(tuple(x.cpu().pin_memory() for x in (u, i, j)))

    def e2e_step(f):
        if sharded is None:
            return eng.train_step_bpr(P, Q, opt, f[0], f[1], f[2], reg=reg)  # returns the host loss (sync)
        return sharded.step(reg, feed=f)
    e2e_step(feeds[n_e2e])  # warm
    barrier()
    t0 = time.perf_counter()
    for k in range(n_e2e):
        e2e_step(feeds[k])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    def max_over_ranks(sec):
        if world > 1:
            t = torch.tensor([sec], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return sec
    e2e_sync = world * B * n_e2e / max_over_ranks(e2e_s)   # one blocking call per step: feed in, loss out, host waits (the reference's sess.run)
    e2e_path = "crb_train_step_bpr per step (host int32 feeds, host loss, blocking)"
    if sharded is None:
        # the reference's epoch loop (RankingRecommender.py:39-46) as ONE call over the caller-sampled epoch arrays: every step still
        # copies its own 12*B bytes of feed from pinned host memory and returns its own loss to pinned host memory, but step k+1's
        # feed is staged on the copy stream while step k computes
        eu, ei, ej = (torch.cat([f[c] for f in feeds[:n_e2e]]).pin_memory() for c in range(3))
        host_losses = torch.zeros(n_e2e, dtype=torch.float64).pin_memory()
        eng.train_epoch_bpr_feeds(P, Q, opt, feeds[n_e2e][0], feeds[n_e2e][1], feeds[n_e2e][2], B, reg, host_losses[:1])  # warm
        barrier()
        t0 = time.perf_counter()
        eng.train_epoch_bpr_feeds(P, Q, opt, eu, ei, ej, B, reg, host_losses)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        assert bool(torch.isfinite(host_losses).all()) and float(host_losses.min()) > 0
        e2e_path = "crb_train_epoch_bpr_feeds: the reference's epoch loop over host feed arrays, per-step H2D feed + per-step D2H loss, feeds staged one step ahead"
    if sharded is not None:
        # the same epoch loop on the multi-GPU path: every rank's feeds are staged one step ahead (crb_shard_step_prepare with feeds),
        # every step's loss goes to pinned host memory with its own copy
        host_losses = torch.zeros(n_e2e, dtype=torch.float64).pin_memory()
        sharded.run_steps(1, reg, neg_ratio=R, seed=0, epoch=0, feeds=[feeds[n_e2e]], host_losses=host_losses[:1])  # warm
        barrier()
        t0 = time.perf_counter()
        sharded.run_steps(n_e2e, reg, neg_ratio=R, seed=0, epoch=0, feeds=feeds[:n_e2e], host_losses=host_losses)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        assert bool(torch.isfinite(host_losses).all()) and float(host_losses.min()) > 0
        e2e_path = "ShardedBPR.run_steps(feeds=...): the epoch loop over host feed arrays, per-step H2D feed + per-step D2H loss, feeds staged one step ahead"
    e2e_value = world * B * n_e2e / max_over_ranks(e2e_s)
    # the same loop with the SAMPLER inside the timed region (what the reference arm times): triplets drawn on the device from the
    # resident history (no host input exists for such a step), every step's loss returned to the host
    n_ws = args.steps
    first_ws = 0 if (steps_total + 2 * n_ws) * B > rows else steps_total * B
    if sharded is None:
        hl = np.zeros(n_ws, dtype=np.float64)
        eng.train_epoch_bpr(P, Q, opt, 9, 1, first_ws, B, 1, R, reg, hl[:1])   # warm
        barrier()
        t0 = time.perf_counter()
        eng.train_epoch_bpr(P, Q, opt, 9, 1, first_ws, B, n_ws, R, reg, hl)     # returns when the losses are on the host
        ws_s = time.perf_counter() - t0
    else:
        hl = torch.zeros(n_ws, dtype=torch.float64).pin_memory()
        sharded.run_steps(1, reg, neg_ratio=R, seed=9, epoch=1, first=first_ws, batch=B, host_losses=hl[:1])   # warm
        barrier()
        t0 = time.perf_counter()
        sharded.run_steps(n_ws, reg, neg_ratio=R, seed=9, epoch=1, first=first_ws, batch=B, host_losses=hl)
        torch.cuda.synchronize()
        ws_s = time.perf_counter() - t0
    e2e_with_sampling = world * B * n_ws / max_over_ranks(ws_s)
    if sharded is not None:
        sharded.check()   # raises if a cross-rank barrier timed out or the sampler gave up on a row

    # ---- roofline of the dominant kernel (K3 fused step) ----
    hbm, tf, which = peaks()
    per_triplet = {"SGD": 24, "Adagrad": 48, "Adam": 72}[opt_kind] * dim  # algorithmic bytes / triplet
    k3_avg_ms = k3_ms / max(1, k3_n)
    achieved = per_triplet * B / (k3_avg_ms / 1000.0) / 1e9 if k3_n else None
    roofline = {"bound": "hbm", "kernel": "bpr_step_kernel" if world == 1 else "shard_step_kernel", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": (achieved / hbm) if achieved else None, "traffic": None, "peak_source": which,
                "algorithmic_bytes_per_launch": per_triplet * B, "kernel_ms": k3_avg_ms, "kernel_share_of_step": k3_avg_ms / ms_step,
                "step_frac": per_triplet * B / (ms_step / 1000.0) / 1e9 / hbm}
    roofline["other_kernels_ms"] = {k: v[0] / v[1] for k, v in other_kernels.items() if v[1]}
    if phase_ms:
        roofline["phase_ms_per_step"] = {k: v / phase_ms["steps"] for k, v in phase_ms.items() if k != "steps"}
    if world > 1 and k3_n:
        # per triplet two item rows are fetched and two gradients returned, (N-1)/N of them over NVLink.  Every GPU's
        # ingress carries the rows it fetches plus the gradients its peers send it (and its egress the mirror image): 16*d bytes per
        # direction per triplet.  Peak: NVLink 5, 900 GB/s per direction per GPU (raw link rate; not measured on this pool).
        nv_bytes = 16 * dim * B * (world - 1) / world
        roofline["nvlink"] = {"undeduplicated_bytes_per_direction_per_step": nv_bytes, "peak": 770.0, "unit": "GB/s per direction per GPU",
                              "peak_source": "measured peer copy on this pool (B200_PROFILING.md); 900 nominal",
                              "note": "16*d*B*(N-1)/N is what the step would move without the per-rank de-duplication (round 1); the measured "
                                      "counters below are what it does move"}
        # exact payload of one step, counted from the step's own indices (nvidia-smi's NVLink counters read N/A on this pool and
        # ncu may not wrap a multi-rank command): every DISTINCT remote item row is fetched once and its summed gradient sent once
        try:
            allit = torch.cat([feeds[0][1], feeds[0][2]]).to(dev).long()
            remote = allit[allit % world != rank]
            n_remote_rows = int(torch.unique(remote).numel())
            t = torch.tensor([float(n_remote_rows)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wire = 2.0 * float(t.item()) * dim * 4
            roofline["nvlink"].update({"counted_bytes_per_direction_per_step": wire, "counted_from": "2 x distinct remote item rows of one step's batch (max over ranks) x d x 4: "
                                       "ingress = the rows this rank fetches + the gradients its peers send it (the mirror image by symmetry), egress likewise",
                                       "remote_occurrences": int(remote.numel()), "distinct_remote_rows": n_remote_rows,
                                       "rate_over_step": wire / (ms_step / 1000.0) / 1e9, "frac_of_peak_over_step": wire / (ms_step / 1000.0) / 1e9 / 770.0})
        except Exception as e:
            roofline["nvlink"]["counted_error"] = str(e)
        if nv0 and nv1:
            tx, rx = (nv1[0] - nv0[0]) / args.steps, (nv1[1] - nv0[1]) / args.steps
            roofline["nvlink"].update({"measured_tx_bytes_per_step": tx, "measured_rx_bytes_per_step": rx,
                                       "measured_rate_over_step": max(tx, rx) / (ms_step / 1000.0) / 1e9,
                                       "counter": "nvidia-smi nvlink -gt d on rank 0's GPU, read around the timed steps (includes the sampler's and barriers' few bytes)"})
        roofline["traffic"] = None
    traffic_file = os.path.join(ROOT, "profiles", "r02_bpr_step_traffic.json")   # dram__bytes_read + write of one K3 launch, this round's ncu capture
    if os.path.exists(traffic_file):
        try:
            # the committed capture is of the default configuration's kernel (Adam, tf1 semantics, uniform popularity, default shape)
            default_cfg = opt_kind == "Adam" and adam_mode == "tf1" and args.item_popularity == "uniform" and not (args.users or args.items or args.batch or args.dim)
            if world == 1 and not args.sharded and default_cfg:
                roofline["traffic"] = json.load(open(traffic_file)).get(args.workload)
        except Exception:
            pass

    # ---- secondary metric: full-rank top-20 evaluation users/sec ----
    ev = None
    if args.eval_users > 0 and sharded is None:
        eng.adam_flush(P, opt)
        eng.adam_flush(Q, opt)
        n_eval = min(args.eval_users, u_hi - u_lo)
        eu = torch.arange(u_lo, u_lo + n_eval, device=dev, dtype=torch.int32)
        try:
            eng.score_topk(0, P.w, Q.w, eu, 20, exact=args.eval_exact)  # warm: same size, so the workspace is allocated outside the timed call
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            eng.score_topk(0, P.w, Q.w, eu, 20, exact=args.eval_exact)
            e1.record()
            barrier()
            ems = e0.elapsed_time(e1)
            flops = 2.0 * n_eval * items * dim
            ev = {"metric": "fullrank_top20_eval_users_per_sec", "value": world * n_eval / (ems / 1000.0), "unit": "users/s", "users": n_eval,
                  "items": items, "ms": ems, "path": "fp32 CUDA cores (exact)" if args.eval_exact else "bf16 tcgen05 + certified fp32 rescoring",
                  "roofline": {"bound": "tensor", "achieved": flops / (ems / 1000.0) / 1e12, "peak": tf, "unit": "TFLOP/s",
                               "frac": flops / (ems / 1000.0) / 1e12 / tf}}
            if not args.eval_exact:
                ev["stats"] = eng.score_topk_stats()
        except Exception as e:  # reported, never hidden
            ev = {"error": str(e)}

    # ---- the reference's test.batch_size regime: 4096 users per call, item table already converted by the previous call ----
    if args.eval_users > 0 and sharded is None and ev is not None and "error" not in ev:
        try:
            su = torch.arange(u_lo, u_lo + min(4096, u_hi - u_lo), device=dev, dtype=torch.int32)
            eng.score_topk(0, P.w, Q.w, su, 20, exact=args.eval_exact)   # warm (same shape: workspace and bf16 item table in place)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for _ in range(4):
                eng.score_topk(0, P.w, Q.w, su, 20, exact=args.eval_exact)
            e1.record()
            barrier()
            sms = e0.elapsed_time(e1) / 4
            ev["small_call"] = {"users": int(su.numel()), "ms": sms, "users_per_sec": su.numel() / (sms / 1000.0),
                                "frac_of_tensor_peak": 2.0 * su.numel() * items * dim / (sms / 1000.0) / 1e12 / tf,
                                "note": "4 back-to-back calls of 4096 users; the bf16 item table is kept between calls"}
        except Exception as e:
            ev["small_call"] = {"error": str(e)}

    # ---- the whole user table through one call (north_star: "full-rank top-20 evaluation of 10M users x 2M items in seconds") ----
    if args.eval_full and sharded is None and ev is not None and "error" not in ev:
        try:
            allu = torch.arange(u_lo, u_hi, device=dev, dtype=torch.int32)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            full_ids = eng.score_topk(0, P.w, Q.w, allu, 20, exact=args.eval_exact)
            e1.record()
            barrier()
            fms = e0.elapsed_time(e1)
            ev["full_sweep"] = {"users": int(allu.numel()), "items": items, "seconds": fms / 1000.0, "users_per_sec": allu.numel() / (fms / 1000.0),
                                "frac_of_tensor_peak": 2.0 * allu.numel() * items * dim / (fms / 1000.0) / 1e12 / tf,
                                "stats": eng.score_topk_stats(), "measured": True}
            del full_ids, allu
        except Exception as e:
            ev["full_sweep"] = {"error": str(e)}

    # ---- secondary metric: sampled-candidate evaluation (test_model_loo with 1000 negatives, SURVEY 8(d) "Eval loo") ----
    ev_loo = None
    if args.eval_users > 0 and sharded is None and ev is not None and "error" not in ev:
        try:
            n_loo, n_cand = min(65536, args.eval_users, u_hi - u_lo), 1001
            g2 = torch.Generator(device=dev).manual_seed(99)
            lu = torch.arange(n_loo, device=dev, dtype=torch.int32)
            li = torch.randint(0, items, (n_loo * n_cand,), device=dev, generator=g2, dtype=torch.int32)
            seg = torch.arange(n_loo + 1, device=dev, dtype=torch.int64) * n_cand

            def loo_once():
                # pre_scores of every (user, candidate) pair + np.argsort(-scores_u)[:20] per user, one kernel (crb_score_pairs_topk)
                return eng.score_pairs_topk(0, P.w, Q.w, lu, li, seg, 20)
            loo_once()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            loo_once()
            e1.record()
            barrier()
            lms = e0.elapsed_time(e1)
            lbytes = float(n_loo) * n_cand * 4 * dim                    # (1 + neg_samples) * 4 * d bytes per user: the gathered item rows
            ev_loo = {"metric": "loo_1000neg_top20_eval_users_per_sec", "value": n_loo / (lms / 1000.0), "unit": "users/s", "users": n_loo,
                      "candidates_per_user": n_cand, "ms": lms,
                      "roofline": {"bound": "hbm", "achieved": lbytes / (lms / 1000.0) / 1e9, "peak": hbm, "unit": "GB/s",
                                   "frac": lbytes / (lms / 1000.0) / 1e9 / hbm, "algorithmic_bytes": lbytes}}
            ev_loo["path"] = "loo_topk_kernel: scores and top-20 in one kernel, the score vector never reaches HBM"
            del lu, li
        except Exception as e:
            ev_loo = {"error": str(e)}

    if args.eval_users > 0 and sharded is not None:
        from cleverrec_b200.dist import ShardedEval
        try:
            n_eval = min(args.eval_users, u_hi - u_lo)
            sev = ShardedEval(sharded, rowptr, pi)
            sev.topk(20, batch_users=n_eval, limit=n_eval)  # warm (also flushes pending Adam decay and sizes the workspace)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            sev.topk(20, batch_users=n_eval, limit=n_eval)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
            flops = 2.0 * world * n_eval * items * dim
            ev = {"metric": "fullrank_top20_eval_users_per_sec", "value": world * n_eval / (ems / 1000.0), "unit": "users/s", "users": world * n_eval,
                  "users_per_gpu": n_eval, "items": items, "ms": ems,
                  "path": "item shards all-gathered over NVLink once, then every rank ranks its own users: bf16 tcgen05 + certified fp32 rescoring",
                  "roofline": {"bound": "tensor", "achieved": flops / (ems / 1000.0) / 1e12 / world, "peak": tf, "unit": "TFLOP/s per GPU",
                               "frac": flops / (ems / 1000.0) / 1e12 / world / tf}}
            if args.eval_full:
                # every user of the job: each rank ranks ALL its own users in one call (the all-gather of the item shards is inside)
                n_all = u_hi - u_lo
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                e0.record()
                sev.topk(20, batch_users=n_all)
                e1.record()
                barrier()
                t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                fms = float(t.item())
                ev["full_sweep"] = {"users": users, "items": items, "seconds": fms / 1000.0, "users_per_sec": users / (fms / 1000.0),
                                    "frac_of_tensor_peak": 2.0 * users * items * dim / (fms / 1000.0) / 1e12 / tf / world, "measured": True}
        except Exception as e:
            ev = {"error": str(e)}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        times, done, n_sample, table_rows = cpu_reference_steps(w, 3, args.ref_batch or B, threads)
        sec = sum(times[1:]) / len(times[1:])
        cpu = {"value": done / sec, "unit": UNIT, "cores": threads, "kind": "port", "sample": "2 timed steps (1 warm-up) of " + cpu_sample_text(done, n_sample, table_rows, w)}
    if world > 1:
        dist.barrier()   # the other ranks wait for rank 0's host baseline before the process group goes away

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, w, world),
                "run_info": {"interactions_per_gpu": n_pos, "setup_s": round(t_setup, 1)},
                "clocks": clk, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 12 * B, "d2h_bytes_per_step": 8, "steps": n_e2e, "path": e2e_path,
                        "blocking_per_step_value": e2e_sync,
                        "with_device_sampling": {"value": e2e_with_sampling, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8,
                                                 "note": "sampler inside the timed region like the reference arm: triplets drawn on the device, per-step loss to the host"}},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "eval": ev, "eval_loo": ev_loo, "final_loss": loss_last}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CRB_WORKLOAD", "s_large"), choices=sorted(WORKLOADS))
    ap.add_argument("--optimizer", default="Adam", choices=["SGD", "Adagrad", "Adam"])
    ap.add_argument("--adam-mode", dest="adam_mode", default="tf1", choices=["tf1", "lazy"])
    ap.add_argument("--eval-users", dest="eval_users", type=int, default=262144)
    ap.add_argument("--eval-exact", dest="eval_exact", action="store_true")
    ap.add_argument("--no-eval-full", dest="eval_full", action="store_false", help="skip the sweep over ALL users (about 5 s at 10M users on one GPU)")
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--ref-batch", dest="ref_batch", type=int, default=0, help="triplets per step of the CPU baseline / reference arm (default: the GPU arm's batch)")
    ap.add_argument("--item-popularity", dest="item_popularity", default="uniform", choices=["uniform", "zipf"],
                    help="popularity of the positives' items in the synthetic history (zipf = Zipf(1.0): hub rows repeat ~10^4 times per batch)")
    ap.add_argument("--sharded", action="store_true", help="use the multi-GPU code path even at N=1 (experiments)")
    ap.add_argument("--phases", action="store_true", help="multi-GPU path: split every step into compute / barrier / apply with CUDA events (rank 0)")
    ap.add_argument("--users", type=int, default=0, help="override the workload's user count (experiments)")
    ap.add_argument("--items", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--dim", type=int, default=0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3
    w = dict(WORKLOADS[args.workload])
    for k in ("users", "items", "batch", "dim"):
        if getattr(args, k):
            w[k] = getattr(args, k)
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
