"""GPU, 2 ranks: replicated-table data-parallel training == single-GPU training on the concatenated batch.
Launched by the test through torch.distributed.run on one box (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["replay", "dense"])
def test_two_rank_training_equals_single_gpu(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "dist_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout


@pytest.mark.parametrize("shard", ["candidates", "queries"])
def test_two_rank_sharded_allpairs_topk_equals_oracle(shard):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29543", os.path.join(ROOT, "tests", "dist_worker.py"), "sim_" + shard]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout


@pytest.mark.parametrize("mode", ["replay", "dense"])
def test_two_rank_row_sharded_training_equals_single_gpu(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29545", os.path.join(ROOT, "tests", "dist_worker.py"), "shard_" + mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout


def _peer_env(mode):
    """replay_staged: the per-step stage kernels of csrc/peer.inl instead of the persistent kernel;
    replay_strict: the persistent kernel with system-scope fences in front of the row words (AR_PEER_STRICT=1)."""
    env = dict(os.environ)
    env.pop("AR_PEER_STAGED", None)
    env.pop("AR_PEER_STRICT", None)
    if mode.endswith("_staged"):
        env["AR_PEER_STAGED"] = "1"
    if mode.endswith("_strict"):
        env["AR_PEER_STRICT"] = "1"
    return mode.replace("_staged", "").replace("_strict", ""), env


@pytest.mark.parametrize("mode", ["replay", "replay_staged", "replay_strict", "dense", "touched"])
def test_two_rank_peer_memory_training_equals_single_gpu(mode):
    """csrc/peer.inl + the peer mode of csrc/chunk.inl: owners pull the other table's rows over NVLink, flags
    instead of collectives."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mode, env = _peer_env(mode)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "tests", "dist_worker.py"), "peer_" + mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout


@pytest.mark.parametrize("mode", ["replay", "replay_staged", "replay_strict"])
def test_one_rank_peer_memory_training_equals_single_gpu(mode):
    """The peer-memory path with a world of ONE rank (every pull is local): runs on a single-GPU box, so the
    persistent peer kernel, peer_select / peer_fwd / peer_pull and the chunked planning are covered there too."""
    mode, env = _peer_env(mode)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "1", "--master-addr",
           "127.0.0.1", "--master-port", "29549", os.path.join(ROOT, "tests", "dist_worker.py"), "peer_" + mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout


def test_two_rank_sharded_scoring():
    """cfg4 (model_recs over many users): users sharded over 2 GPUs, gathered result == single-GPU result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29551", os.path.join(ROOT, "tests", "dist_worker.py"), "score"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout


def test_two_rank_distributed_fit_equals_single_gpu_fit():
    """dist_fit.DistributedEmbeddingDotModel.fit (epochs, seeded shuffle, validation, callbacks, sharded tables)
    against EmbeddingDotModel.fit with the global batch on one GPU: histories to 1e-4."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29553", os.path.join(ROOT, "tests", "dist_worker.py"), "fit"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout


def test_one_rank_distributed_fit_equals_single_gpu_fit():
    """The same with a world of one rank: runs on a single-GPU box."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "1", "--master-addr",
           "127.0.0.1", "--master-port", "29555", os.path.join(ROOT, "tests", "dist_worker.py"), "fit"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout
