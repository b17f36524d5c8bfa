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
