"""A handful of eager decode steps of the persistent kernel on Llama-3.1-8B shapes (random-init weights) — the target of
`ncu -k regex:megaDecode` captures (one launch = one token)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_llama_b200.models.config import get_config
from distributed_llama_b200.models.loader import synthetic_device_weights
from distributed_llama_b200.runtime import Engine
eng = Engine(synthetic_device_weights(get_config("llama-3.1-8b"), 0, 1, "cuda:0", max_seq_len=2048))
eng.enable_mega()
eng.prefill([(7 * i + 3) % 1000 + 1 for i in range(16)], 0, want_logits=False)
print(eng.decode_greedy(5, 16, 8, use_graph=False), "mega" if eng.mega_active else "multi-kernel")
torch.cuda.synchronize()
