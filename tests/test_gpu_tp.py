"""Multi-GPU tensor parallelism (needs >= 2 GPUs on the box): fused all-reduce path vs TP=1 and the oracle."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [2, 4, 8])
def test_tp_matches_single_gpu(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29610 + n), os.path.join(ROOT, "tools", "tp_check.py"), "tiny-llama-tp8" if n == 8 else "tiny-llama31"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert "TP_CHECK PASS" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("mode", ["tp", "ep"])
def test_moe_tensor_and_expert_parallel(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29633", os.path.join(ROOT, "tools", "tp_check.py"), "tiny-qwen3-moe", mode]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert "TP_CHECK PASS" in r.stdout and f"moe_mode={mode}" in r.stdout, r.stdout[-3000:]
