"""-m gpu, needs >= 2 GPUs on one box (run with `gpurun --gpus 2`; skipped on one GPU): multi-GPU BPR over NVLink peer memory
reproduces the single-GPU step on the union batch (loss to 1e-5, tables to the single-GPU tolerances)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2])
def test_sharded_bpr_matches_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "_sharded_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-4000:]


@pytest.mark.parametrize("world", [2])
def test_bpr_model_class_under_torchrun(world):
    """The reference-facing class (cleverrec_b200.model.ranking.BPR) on the golden ml-100k splits with one process per GPU: losses
    agree across ranks and fall, and the returned HR / MRR / NDCG lists are bit-identical to the reference's own evaluation loop
    run on the gathered tables."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(ROOT, "tests", "_sharded_model_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "SHARDED_MODEL_OK" in r.stdout, r.stdout[-4000:]


@pytest.mark.parametrize("world", [2])
def test_main_tuning_under_torchrun(world, tmp_path):
    """torchrun -m cleverrec_b200.main_tuning: the grid of main_tuning.py:38-45 dealt round-robin to the ranks (one GPU each, every
    combination a single-GPU replica), results gathered on rank 0 in grid order."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from test_gpu_main_flow import TUNING_CONF, write_tree
    write_tree(tmp_path, "loo", 99)
    (tmp_path / "conf" / "BPR.properties").write_text(TUNING_CONF.replace("tuning.workers=2", "tuning.workers=1"))
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29523", "-m", "cleverrec_b200.main_tuning", "."]
    r = subprocess.run(cmd, cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:]
    assert r.stdout.count("[tuning ") == 4                     # two combinations per rank
    assert r.stdout.count("best by NDCG@topk[0]:") == 1        # the summary is rank 0's
    for combo in ("'embed_size': 16, 'reg': 0.01", "'embed_size': 16, 'reg': 0.1", "'embed_size': 32, 'reg': 0.01", "'embed_size': 32, 'reg': 0.1"):
        assert combo in r.stdout
