"""`dllama-native` (C++ CLI + native engine driver, no interpreter) against the Python front end on the same files."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pred(out: str):
    return [ln.split("|")[-1] for ln in out.splitlines() if "🔶 Pred" in ln]


@pytest.mark.parametrize("name", ["tiny-llama31", "tiny-qwen3", "tiny-qwen3-moe"])
def test_native_binary_matches_python_cli(tmp_models, name):
    m, t = tmp_models[name]
    prompt = "Hello world, the model " * 6          # long enough for the tensor-core prefill path on dense models
    common = ["inference", "--model", m, "--tokenizer", t, "--buffer-float-type", "q80", "--prompt", prompt, "--steps", "64",
              "--temperature", "0"]
    outs = []
    for exe in ("dllama-native", "dllama"):
        r = subprocess.run([os.path.join(ROOT, exe)] + common, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:]
        assert "Prediction" in r.stdout and "tokens/s" in r.stdout
        outs.append(_pred(r.stdout))
    assert len(outs[0]) >= 8 and outs[0] == outs[1]


def test_native_perplexity_and_sampling(tmp_models):
    m, t = tmp_models["tiny-llama31"]
    r = subprocess.run([os.path.join(ROOT, "dllama-native"), "perplexity", "--model", m, "--tokenizer", t, "--prompt",
                        "The quick brown fox jumps over the lazy dog"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "perplexity:" in r.stdout, r.stdout[-2000:]
    p = subprocess.run([os.path.join(ROOT, "dllama"), "perplexity", "--model", m, "--tokenizer", t, "--buffer-float-type", "q80", "--prompt",
                        "The quick brown fox jumps over the lazy dog"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    get = lambda s: float([ln for ln in s.splitlines() if "perplexity:" in ln][0].split(":")[1].split()[0])
    assert abs(get(r.stdout) - get(p.stdout)) / get(p.stdout) < 1e-3
    # seeded sampling is reproducible
    runs = []
    for _ in range(2):
        s = subprocess.run([os.path.join(ROOT, "dllama-native"), "inference", "--model", m, "--tokenizer", t, "--prompt", "Hello", "--steps", "24",
                            "--temperature", "0.8", "--topp", "0.9", "--seed", "12345"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=600)
        assert s.returncode == 0, s.stdout[-2000:]
        runs.append(_pred(s.stdout))
    assert runs[0] == runs[1] and len(runs[0]) >= 8
