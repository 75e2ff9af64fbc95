"""The gene-sharded path under torchrun, inside the GPU test-suite: two ranks run GeneClust-fast on their gene shards
and the result is compared with the single-GPU run of the same matrix (scripts/check_sharded_parity.py: PCA scores
within 1e-3 relative, labels and selected-gene list identical).  With >= 2 GPUs the ranks talk NCCL; on a single-GPU box
both ranks share cuda:0 and gloo carries the CUDA tensors, so the sharded code path is always exercised."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(extra, nproc=2, timeout=900, script="check_sharded_parity.py"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "scripts", script), *extra]
    res = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    print(res.stdout[-3000:])
    return res


def _need_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    return torch


def test_two_ranks_one_gpu_gloo_sharded_equals_single():
    _need_cuda()
    res = _torchrun(["12000", "5000", "--backend", "gloo", "--same-device"])
    assert res.returncode == 0, res.stdout[-2000:]
    assert "tail exact True" in res.stdout


def test_two_ranks_host_csr_shards_gloo():
    _need_cuda()
    res = _torchrun(["6000", "3000", "--backend", "gloo", "--same-device", "--csr"])
    assert res.returncode == 0, res.stdout[-2000:]


def test_two_ranks_nccl_sharded_equals_single():
    torch = _need_cuda()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL refuses two ranks on one device); the gloo variant above covers the code path")
    res = _torchrun(["20000", "6000", "--backend", "nccl"])
    assert res.returncode == 0, res.stdout[-2000:]


def test_two_ranks_geneclust_ps_sharded_equals_single():
    """BASELINE.json config c4's entry: `scGeneClust(version='ps')` on gene-sharded host counts under torchrun - relevance
    per gene shard, pair ranges streamed into per-rank forests and merged, complementarity split by edges - selects the
    same genes as the single-GPU run."""
    torch = _need_cuda()
    if torch.cuda.device_count() >= 2:
        res = _torchrun(["3000", "2500", "--backend", "nccl"], script="check_sharded_ps.py")
    else:
        res = _torchrun(["3000", "2500", "--backend", "gloo", "--same-device"], script="check_sharded_ps.py")
    assert res.returncode == 0, res.stdout[-3000:]
    assert "identical to single GPU: True" in res.stdout


def test_pca_is_run_to_run_deterministic():
    """Every raw pass and panel normalisation of the PCA must return the same bits in every repetition
    (scripts/check_pca_determinism.py checksums them on the device).  Regression test of a stale-tile hazard in the pass
    kernel's raw-tile ring that only showed under memory-system contention (2 of 15 gene-sharded runs over NCCL)."""
    torch = _need_cuda()
    script = os.path.join(ROOT, "scripts", "check_pca_determinism.py")
    res = subprocess.run([sys.executable, script, "300000", "3776", "6"], cwd=ROOT, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=600)
    print(res.stdout[-1500:])
    assert res.returncode == 0 and "DIFFERS" not in res.stdout and res.stdout.count("identical (") == 5
    if torch.cuda.device_count() >= 2:
        res = _torchrun(["600000", "7552", "8"], script="check_pca_determinism.py")
        assert res.returncode == 0 and "DIFFERS" not in res.stdout and res.stdout.count("identical (") == 7, res.stdout[-2000:]
