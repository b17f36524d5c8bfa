#!/usr/bin/env python
"""HuggingFace checkpoint -> `.m` model file (Llama / Mistral / Qwen3 / Qwen3-MoE).

    python tools/convert_hf.py <hfFolder> <q40|q80|f16|f32> <name>      ->  dllama_model_<name>_<type>.m

Behavioural parity with the reference converter (converter/convert-hf.py:59-236): same header keys, same tensor order,
Llama q/k projections re-ordered from HF's half-split rotary layout to interleaved pairs, lm_head falling back to the tied
embedding. Implementation differences: safetensors shards are indexed once (tensor name -> shard) and opened lazily, and
tensors are quantised with the vectorised numpy codecs of this repo.
"""
from __future__ import annotations

import json
import os
import sys
from typing import Dict, List

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from distributed_llama_b200.formats import quants, write_model_header, write_tensor
from distributed_llama_b200.models.config import ARCH_LLAMA, ARCH_QWEN3, ARCH_QWEN3_MOE

ARCH_BY_MODEL_TYPE = {"llama": ARCH_LLAMA, "mistral": ARCH_LLAMA, "qwen3": ARCH_QWEN3, "qwen3_moe": ARCH_QWEN3_MOE}
ROPE_TYPE_LLAMA3_1 = 2


def to_interleaved(w: torch.Tensor, n_heads: int) -> torch.Tensor:
    """Rows of each head go from [first half | second half] to (x0, x_{h/2}, x1, x_{h/2+1}, ...)."""
    rows = w.shape[0]
    hd = rows // n_heads
    return w.reshape(n_heads, 2, hd // 2, *w.shape[1:]).transpose(1, 2).reshape(w.shape)


def header_from_config(cfg: dict, weights_float_type: int) -> Dict[str, int]:
    mt = cfg["model_type"]
    if mt not in ARCH_BY_MODEL_TYPE:
        raise ValueError(f"Unsupported arch type: {mt}")
    act = {"gelu": 0, "silu": 1}.get(cfg["hidden_act"])
    if act is None:
        raise ValueError(f"Unsupported hidden act: {cfg['hidden_act']}")
    # the expert count is saved as `num_experts` by the published Qwen3-MoE checkpoints and as `num_local_experts` by transformers >= 5
    p = {"version": 0, "arch_type": ARCH_BY_MODEL_TYPE[mt], "hidden_act": act, "dim": cfg["hidden_size"],
         "hidden_dim": cfg["intermediate_size"], "n_layers": cfg["num_hidden_layers"], "n_heads": cfg["num_attention_heads"],
         "n_kv_heads": cfg["num_key_value_heads"], "weights_float_type": weights_float_type,
         "max_seq_len": cfg["max_position_embeddings"], "vocab_size": cfg["vocab_size"],
         "n_experts": int(cfg.get("num_experts") or cfg.get("num_local_experts") or 0), "n_active_experts": int(cfg.get("num_experts_per_tok") or 0)}
    rope_params = cfg.get("rope_parameters") or {}
    theta = cfg.get("rope_theta", rope_params.get("rope_theta"))
    if theta is not None:
        p["rope_theta"] = int(theta)
    scaling = cfg.get("rope_scaling") or (rope_params if rope_params.get("rope_type") not in (None, "default") else None)
    if scaling:
        if scaling.get("rope_type", scaling.get("type")) != "llama3":
            raise ValueError(f"Unsupported rope type: {scaling.get('rope_type')}")
        p.update(rope_scaling_factor=int(scaling["factor"]), rope_scaling_low_freq_factor=int(scaling["low_freq_factor"]),
                 rope_scaling_high_freq_factory=int(scaling["high_freq_factor"]),
                 rope_scaling_orig_max_seq_len=int(scaling["original_max_position_embeddings"]), rope_type=ROPE_TYPE_LLAMA3_1)
    if cfg.get("head_dim") is not None:
        p["head_dim"] = cfg["head_dim"]
    eps = cfg.get("rms_norm_eps")
    if eps is not None:
        if eps == 1e-05:
            p["norm_epsilon"] = 5
        elif eps == 1e-06:
            p["norm_epsilon"] = 6
        else:
            raise ValueError(f"Unsupported epsilon: {eps}")
    if cfg.get("moe_intermediate_size") is not None:
        p["moe_hidden_dim"] = int(cfg["moe_intermediate_size"])
    return p


class ShardIndex:
    def __init__(self, folder: str):
        from safetensors import safe_open
        self._open = safe_open
        self.files = sorted(os.path.join(folder, f) for f in os.listdir(folder) if f.endswith(".safetensors") and not f.startswith("."))
        if not self.files:
            raise FileNotFoundError("Not found any model file")
        self.where: Dict[str, str] = {}
        for f in self.files:
            with safe_open(f, framework="pt", device="cpu") as h:
                for k in h.keys():
                    self.where[k] = f
        self._cur, self._handle = None, None

    def get(self, *names: str) -> torch.Tensor:
        for n in names:
            if n in self.where:
                f = self.where[n]
                if f != self._cur:
                    self._handle = self._open(f, framework="pt", device="cpu")
                    self._cur = f
                return self._handle.get_tensor(n)
        raise KeyError(f"Layer {names[0]} not found")


def convert(folder: str, float_type_name: str, out_path: str) -> str:
    wt = quants.parse_float_type(float_type_name)
    with open(os.path.join(folder, "config.json")) as f:
        cfg = json.load(f)
    params = header_from_config(cfg, wt)
    arch = params["arch_type"]
    idx = ShardIndex(folder)
    n_heads, n_kv = params["n_heads"], params["n_kv_heads"]
    with open(out_path, "wb") as out:
        write_model_header(out, params)

        def emit(t, ftype):
            print(f"🔶 Writing tensor {tuple(t.shape)} as {quants.float_type_name(ftype)}")
            write_tensor(out, t, ftype)

        emit(idx.get("model.embed_tokens.weight"), quants.F_32)
        for l in range(params["n_layers"]):
            pre = f"model.layers.{l}."
            q, k = idx.get(pre + "self_attn.q_proj.weight"), idx.get(pre + "self_attn.k_proj.weight")
            if arch == ARCH_LLAMA:
                q, k = to_interleaved(q, n_heads), to_interleaved(k, n_kv)
            emit(q, wt); emit(k, wt)
            emit(idx.get(pre + "self_attn.v_proj.weight"), wt)
            emit(idx.get(pre + "self_attn.o_proj.weight"), wt)
            if params["n_experts"] > 0:
                emit(idx.get(pre + "mlp.gate.weight"), quants.F_32)
                for e in range(params["n_experts"]):
                    for part in ("gate_proj", "down_proj", "up_proj"):
                        emit(idx.get(f"{pre}mlp.experts.{e}.{part}.weight"), wt)
            else:
                for part in ("gate_proj", "down_proj", "up_proj"):
                    emit(idx.get(f"{pre}mlp.{part}.weight"), wt)
            if arch in (ARCH_QWEN3, ARCH_QWEN3_MOE):
                emit(idx.get(pre + "self_attn.q_norm.weight"), quants.F_32)
                emit(idx.get(pre + "self_attn.k_norm.weight"), quants.F_32)
            emit(idx.get(pre + "input_layernorm.weight"), quants.F_32)
            emit(idx.get(pre + "post_attention_layernorm.weight"), quants.F_32)
        emit(idx.get("model.norm.weight"), quants.F_32)
        emit(idx.get("lm_head.weight", "model.embed_tokens.weight"), wt)
    return out_path


def main(argv: List[str]) -> int:
    if len(argv) < 3:
        print("Usage: python convert_hf.py <sourceFolderPath> <weightsFloatType> <name>\n\n"
              "  <sourceFolderPath> folder with config.json and *.safetensors\n  <weightsFloatType> q40 | q80 | f16 | f32\n"
              "  <name>             model name used in the output file name")
        return 1
    out = f"dllama_model_{argv[2]}_{argv[1]}.m"
    print(f"Output file: {out}")
    convert(argv[0], argv[1], out)
    print(f"✅ {out} created successfully")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
