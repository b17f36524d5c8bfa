#!/usr/bin/env python
"""Meta `consolidated.*.pth` checkpoint -> `.m` (reference converter/convert-llama.py:14-98).

    python tools/convert_llama.py <modelFolder> <q40|q80|f16|f32>

Shards are concatenated along the axis Meta's model-parallel split used (column-parallel tensors along dim 0,
row-parallel ones — wo, w2, tok_embeddings — along dim 1). The Meta layout already uses interleaved rotary pairs, so no
q/k re-ordering is needed. Layers are processed one at a time to bound host memory.
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from distributed_llama_b200.formats import quants, write_model_header, write_tensor
from distributed_llama_b200.models.config import ARCH_LLAMA

CAT_DIM1 = ("wo.weight", "w2.weight", "tok_embeddings.weight")


def convert(folder: str, float_type_name: str, out_path: str) -> str:
    wt = quants.parse_float_type(float_type_name)
    with open(os.path.join(folder, "params.json")) as f:
        p = json.load(f)
    if p.get("vocab_size", 0) < 1:
        raise ValueError("vocab_size is invalid, please update params.json file")
    if p.get("max_seq_len") is None:
        raise ValueError("max_seq_len is required, please update params.json file")
    shards = sorted(os.path.join(folder, f) for f in os.listdir(folder) if f.startswith("consolidated.") and f.endswith(".pth"))
    if not shards:
        raise FileNotFoundError("no consolidated.*.pth files found")
    models = [torch.load(s, map_location="cpu", mmap=True, weights_only=True) for s in shards]

    def tensor(name: str) -> torch.Tensor:
        parts = [m[name] for m in models]
        if len(parts) == 1 or parts[0].dim() == 1:
            return parts[0]
        return torch.cat(parts, dim=1 if name.endswith(CAT_DIM1) else 0)

    first = tensor("layers.0.feed_forward.w1.weight")
    params = {"version": 0, "arch_type": ARCH_LLAMA, "dim": p["dim"], "hidden_dim": first.shape[0], "n_layers": p["n_layers"],
              "n_heads": p["n_heads"], "n_kv_heads": p.get("n_kv_heads") or p["n_heads"], "n_experts": 0, "n_active_experts": 0,
              "vocab_size": p["vocab_size"], "max_seq_len": p["max_seq_len"], "hidden_act": 1, "weights_float_type": wt}
    if "rope_theta" in p:
        params["rope_theta"] = int(p["rope_theta"])
    if p.get("norm_eps") in (1e-05, 1e-06):
        params["norm_epsilon"] = 5 if p["norm_eps"] == 1e-05 else 6
    with open(out_path, "wb") as out:
        write_model_header(out, params)
        write_tensor(out, tensor("tok_embeddings.weight"), quants.F_32)
        for l in range(p["n_layers"]):
            pre = f"layers.{l}."
            for name in ("attention.wq.weight", "attention.wk.weight", "attention.wv.weight", "attention.wo.weight",
                         "feed_forward.w1.weight", "feed_forward.w2.weight", "feed_forward.w3.weight"):
                write_tensor(out, tensor(pre + name), wt)
            write_tensor(out, tensor(pre + "attention_norm.weight"), quants.F_32)
            write_tensor(out, tensor(pre + "ffn_norm.weight"), quants.F_32)
            print(f"🔶 layer {l + 1}/{p['n_layers']}")
        write_tensor(out, tensor("norm.weight"), quants.F_32)
        write_tensor(out, tensor("output.weight"), wt)
    return out_path


if __name__ == "__main__":
    if len(sys.argv) < 3:
        print("Usage: python convert_llama.py <modelPath> <targetFloatType>")
        sys.exit(1)
    name = os.path.basename(os.path.normpath(sys.argv[1])).lower().replace("-", "_")
    out = f"dllama_model_{name}_{sys.argv[2]}.m"
    convert(sys.argv[1], sys.argv[2], out)
    print(f"✅ {out} created successfully")
