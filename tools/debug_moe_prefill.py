"""MoE prefill: grouped tensor-core path vs the token-by-token GEMV path vs the oracle (tiny-qwen3-moe), and run-to-run determinism."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_llama_b200.formats import ModelFile
from distributed_llama_b200.models.config import get_config
from distributed_llama_b200.models.loader import load_device_weights
from distributed_llama_b200.models.reference import OracleModel
from distributed_llama_b200.models.synthetic import write_synthetic_model
from distributed_llama_b200.runtime import Engine

d = tempfile.mkdtemp()
path = os.path.join(d, "moe.m")
write_synthetic_model(path, get_config("tiny-qwen3-moe"), seed=3)
mf = ModelFile(path)
toks = [(7 * i + 3) % 500 + 1 for i in range(45)]
oracle = OracleModel(mf, act_quant="q80", device="cuda")
ref = oracle.forward(toks, 0)[-1]
outs = []
for rep in range(3):
    eng = Engine(load_device_weights(mf))
    lg = eng.prefill(toks, 0).clone()
    outs.append(lg)
    print(f"grouped run {rep}: max|lg - oracle| = {(lg - ref).abs().max().item():.4f}  argmax {int(lg.argmax())}")
print("grouped deterministic:", all(torch.equal(outs[0], o) for o in outs[1:]))
eng = Engine(load_device_weights(mf))
eng.use_tc_prefill = False
lg_b = eng.prefill(toks, 0).clone()
print(f"token path: max|lg - oracle| = {(lg_b - ref).abs().max().item():.4f}  max|grouped - token| = {(outs[0] - lg_b).abs().max().item():.4f}")
# greedy continuation after each kind of prefill
for tc in (True, False):
    eng = Engine(load_device_weights(mf))
    eng.use_tc_prefill = tc
    eng.prefill(toks[:-1], 0, want_logits=False)
    print("tc" if tc else "tok", eng.decode_greedy(toks[-1], len(toks) - 1, 16))
