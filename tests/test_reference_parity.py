"""Cross-check against the unmodified reference binary (CPU): the reference's `dllama perplexity` on our synthetic `.m`/`.t`
files must report the same per-token probabilities as the PyTorch oracle (which the CUDA engine is tested against).
This pins down file formats, tokenizer encoding, RoPE/QK-norm conventions and the MoE routing for all three families."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref", "distributed-llama")


@pytest.fixture(scope="module")
def ref_exe():
    if not os.path.isdir(REF_DIR):
        if not os.path.isdir("/root/reference"):
            pytest.skip("reference sources not available")
        os.makedirs(os.path.dirname(REF_DIR), exist_ok=True)
        shutil.copytree("/root/reference", REF_DIR)
        subprocess.run(["chmod", "-R", "u+w", REF_DIR])
    exe = os.path.join(REF_DIR, "dllama")
    probe = subprocess.run([exe, "--help"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) if os.path.exists(exe) else None
    if probe is None or probe.returncode not in (0, 1):
        if not shutil.which("make") or not shutil.which("g++"):
            pytest.skip("cannot build the reference here")
        subprocess.run(["make", "clean"], cwd=REF_DIR, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        r = subprocess.run(["make", "dllama", "-j8"], cwd=REF_DIR, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            pytest.skip("reference build failed: " + r.stdout[-300:])
    return exe


@pytest.mark.parametrize("name", ["tiny-llama", "tiny-llama31", "tiny-qwen3", "tiny-qwen3-moe"])
def test_reference_binary_agrees_with_oracle(ref_exe, tmp_models, name):
    from distributed_llama_b200 import host
    from distributed_llama_b200.formats import ModelFile
    from distributed_llama_b200.models.reference import OracleModel
    m, t = tmp_models[name]
    prompt = "Hello world, the model is a llama and the token"
    r = subprocess.run([ref_exe, "perplexity", "--model", m, "--tokenizer", t, "--buffer-float-type", "q80", "--prompt", prompt,
                        "--nthreads", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    ref_probs = [float(x) for x in re.findall(r"prob=([0-9.eE+-]+)", r.stdout)]
    tok = host().Tokenizer(t)
    ids = tok.encode(prompt, True, True)
    assert len(ref_probs) == len(ids) - 1          # same tokenisation length as the reference
    oracle = OracleModel(ModelFile(m), act_quant="q80")
    logits = oracle.forward(ids[:-1], 0)
    probs = torch.softmax(logits, dim=-1)
    ours = [float(probs[i, ids[i + 1]]) for i in range(len(ids) - 1)]
    np.testing.assert_allclose(ours, ref_probs, rtol=0.08, atol=2e-5)
    m_ppl = re.search(r"perplexity: ([0-9.]+)", r.stdout)
    ppl = float(np.exp(-np.mean(np.log(np.maximum(ours, 1e-30)))))
    assert abs(ppl - float(m_ppl.group(1))) / float(m_ppl.group(1)) < 0.05
