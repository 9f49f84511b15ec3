"""bench.py's contract where no GPU is needed: the reference arm (the oracle port of the reference's CPU path -- the one place
bench.py may execute oracle/) prints ONE JSON line with the agreed keys, under torchrun only rank 0 works and prints; the GPU arm
has no CPU fallback (it fails loudly without a device instead of measuring something else)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

ENV = dict(os.environ, CUDA_VISIBLE_DEVICES="")


def _json_lines(text):
    out = []
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("{") and line.endswith("}"):
            try:
                out.append(json.loads(line))
            except ValueError:
                pass
    return out


def test_reference_arm_prints_one_contract_line_under_torchrun():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3",
           "--workload", "ml1m"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=ENV, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1, r.stdout[-3000:]          # rank 1 exits 0 without work
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "bpr_train_triplets_per_sec" and d["unit"] == "triplets/s"
    assert d["n_gpus"] == 2 and d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["config"]["workload"] == "ml1m" and d["run_info"]["sampler_in_timed_region"] is True
    assert d["value"] > 0 and abs(d["value"] - d["run_info"]["triplets_per_step"] / (d["ms_per_step"] / 1000.0)) <= 1e-6 * d["value"]
    # `config` is the workload and nothing else: the object the GPU arm prints for the same launch (both call workload_config)
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    args = argparse.Namespace(workload="ml1m", item_popularity="uniform", optimizer="Adam", adam_mode="tf1")
    assert d["config"] == bench.workload_config(args, bench.WORKLOADS["ml1m"], 2)
    assert d["config"]["batch_per_gpu"] == 6144 and d["config"]["parallelism"].startswith("2 ranks")
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": workload_config(args, w, ') == 2        # the two arms, no third way to build it
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"] and "sampler" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "tiny", "--steps", "1", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=ENV, cwd=ROOT)
    assert r.returncode != 0 and not _json_lines(r.stdout)
