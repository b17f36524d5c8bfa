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


@pytest.mark.parametrize("n,model,mega", [(4, "tiny-llama-kvrep", "0"), (4, "tiny-llama-kvrep", "1"), (8, "tiny-llama-kvrep8", "1")])
def test_kv_head_replication_more_ranks_than_kv_heads(n, model, mega):
    """More ranks than KV heads (4 ranks / 2 KV heads; 8 ranks / 2 KV heads): groups of ranks share a KV head, each with its own
    query heads (the reference cannot run this configuration). Both decode paths at 4 ranks, the persistent kernel at 8."""
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29644 + n), os.path.join(ROOT, "tools", "tp_check.py"), model]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=dict(os.environ, DL_MEGA=mega))
    assert "TP_CHECK PASS" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("model,moe_mode,wtype", [("tiny-llama31", "auto", "q40"), ("tiny-qwen3-moe", "tp", "q40"),
                                                 ("tiny-llama31", "auto", "f32")])
def test_library_collectives_path(model, moe_mode, wtype):
    """NCCL all-reduce between kernel groups: the path taken across nodes (no shared peer-memory domain) and by dense
    f32/f16 weight files; must agree with TP=1 and the oracle like the fused path does."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(ROOT, "tools", "tp_check.py"), model, moe_mode, wtype]
    env = dict(os.environ, DL_COLLECTIVES="nccl")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert "TP_CHECK PASS" in r.stdout and "collectives=nccl" in r.stdout, r.stdout[-3000:]


def test_dllama_cli_spawns_ranks():
    """`dllama inference --gpus 2` (root + one worker process) prints the same continuation as the single-GPU run."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import tempfile
    from distributed_llama_b200.models.config import get_config
    from distributed_llama_b200.models.synthetic import write_synthetic_model, write_synthetic_tokenizer
    with tempfile.TemporaryDirectory() as d:
        m, t = os.path.join(d, "m.m"), os.path.join(d, "t.t")
        write_synthetic_model(m, get_config("tiny-llama31"), seed=5)
        write_synthetic_tokenizer(t, 512)
        outs = []
        for extra in ([], ["--gpus", "2"]):
            r = subprocess.run([os.path.join(ROOT, "dllama"), "inference", "--model", m, "--tokenizer", t, "--buffer-float-type", "q80",
                                "--prompt", "Hello world, the model", "--steps", "20", "--temperature", "0"] + extra,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
            assert r.returncode == 0, r.stdout[-2000:]
            outs.append([ln.split("|")[-1] for ln in r.stdout.splitlines() if "🔶 Pred" in ln])
        assert outs[0] == outs[1] and len(outs[0]) >= 10
