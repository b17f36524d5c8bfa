import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_llama_b200.formats import ModelFile
from distributed_llama_b200.models.config import get_config
from distributed_llama_b200.models.loader import load_device_weights
from distributed_llama_b200.models.synthetic import write_synthetic_model
from distributed_llama_b200.runtime import Engine

name = sys.argv[1] if len(sys.argv) > 1 else "tiny-llama"
path = f"/tmp/dbg_{name}.m"
write_synthetic_model(path, get_config(name), seed=7)
mf = ModelFile(path)
W = load_device_weights(mf)
prompt = [3, 17, 250, 9]
res = {}
for pdl in (False, True):
    for graph in (False, True):
        for rep in range(3):
            eng = Engine(W, use_pdl=pdl)
            eng.prefill(prompt[:-1], 0, want_logits=False)
            out = eng.decode_greedy(prompt[-1], len(prompt) - 1, 40, use_graph=graph)
            res[(pdl, graph, rep)] = out
            torch.cuda.synchronize()
gold = res[(False, False, 0)]
for k, v in res.items():
    first = next((i for i, (a, b) in enumerate(zip(gold, v)) if a != b), None)
    print(k, "OK" if first is None else f"diverges at {first}: {v[max(0,first-1):first+3]} vs {gold[max(0,first-1):first+3]}")
