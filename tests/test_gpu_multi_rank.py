"""-m gpu test of the N > 1 path on real GPUs: tools/multi_gpu_check.py under torchrun with 2 ranks (one proof sharded by LDE coset
ranges over 2 GPUs, NCCL all-gathers at the commitment points) must return, on every rank, proofs byte-identical to the CPU
oracle's for the whole small-program set.  Skipped on a box with one GPU (the CPU-side logic of the N > 1 path is covered by
tests/test_multi_rank_cpu.py with gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_proofs_equal_the_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_CHECK PASS world 2" in out.stdout


def test_single_process_two_devices():
    """dg_init_devices(2): one process, one dg_prove call per proof, two GPUs (host threads + ncclCommInitAll inside the library)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "single_process_check.py"), "2", "14"], cwd=ROOT, capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SINGLE_PROCESS_CHECK PASS devices 2" in out.stdout
