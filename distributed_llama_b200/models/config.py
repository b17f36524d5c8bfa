"""Model configurations (the `.m` header fields) for the families the reference supports.

Presets follow the published HF configs of the models in the reference's launch.py model zoo
(reference launch.py:17-73) and SURVEY §3.4's shape table. `tiny-*` configs are for tests.
"""
from __future__ import annotations

from dataclasses import dataclass, replace, field, asdict
from typing import Dict, Optional

ARCH_LLAMA = 0xABCD00
ARCH_QWEN3 = 0xABCD01
ARCH_QWEN3_MOE = 0xABCD02
ROPE_LLAMA, ROPE_FALCON, ROPE_LLAMA3_1 = 0, 1, 2
ACT_GELU, ACT_SILU = 0, 1


@dataclass
class ModelConfig:
    name: str
    arch_type: int
    dim: int
    hidden_dim: int
    n_layers: int
    n_heads: int
    n_kv_heads: int
    vocab_size: int
    max_seq_len: int
    head_dim: Optional[int] = None
    n_experts: int = 0
    n_active_experts: int = 0
    moe_hidden_dim: int = 0
    rope_theta: int = 10000
    rope_type: Optional[int] = None            # only written when rope scaling is present (as convert-hf does)
    rope_scaling_factor: int = 0
    rope_scaling_low_freq_factor: int = 0
    rope_scaling_high_freq_factory: int = 0
    rope_scaling_orig_max_seq_len: int = 0
    norm_epsilon: int = 5                      # 5 -> 1e-5, 6 -> 1e-6
    hidden_act: int = ACT_SILU

    def header_params(self, weights_float_type: int) -> Dict[str, int]:
        p: Dict[str, int] = {
            "version": 0, "arch_type": self.arch_type, "hidden_act": self.hidden_act, "dim": self.dim,
            "hidden_dim": self.hidden_dim, "n_layers": self.n_layers, "n_heads": self.n_heads,
            "n_kv_heads": self.n_kv_heads, "weights_float_type": weights_float_type,
            "max_seq_len": self.max_seq_len, "vocab_size": self.vocab_size,
            "n_experts": self.n_experts, "n_active_experts": self.n_active_experts,
            "rope_theta": self.rope_theta,
        }
        if self.rope_type is not None:
            p.update(rope_scaling_factor=self.rope_scaling_factor,
                     rope_scaling_low_freq_factor=self.rope_scaling_low_freq_factor,
                     rope_scaling_high_freq_factory=self.rope_scaling_high_freq_factory,
                     rope_scaling_orig_max_seq_len=self.rope_scaling_orig_max_seq_len,
                     rope_type=self.rope_type)
        if self.head_dim is not None:
            p["head_dim"] = self.head_dim
        p["norm_epsilon"] = self.norm_epsilon
        if self.moe_hidden_dim:
            p["moe_hidden_dim"] = self.moe_hidden_dim
        return p


def _llama3(name, dim, hidden, layers, heads, kv, seq=131072, scaling=8, vocab=128256):
    return ModelConfig(name, ARCH_LLAMA, dim, hidden, layers, heads, kv, vocab, seq, rope_theta=500000,
                       rope_type=ROPE_LLAMA3_1, rope_scaling_factor=scaling, rope_scaling_low_freq_factor=1,
                       rope_scaling_high_freq_factory=4, rope_scaling_orig_max_seq_len=8192, norm_epsilon=5)


PRESETS: Dict[str, ModelConfig] = {
    "llama-3.2-1b": _llama3("llama-3.2-1b", 2048, 8192, 16, 32, 8, scaling=32),
    "llama-3.2-3b": _llama3("llama-3.2-3b", 3072, 8192, 28, 24, 8, scaling=32),
    "llama-3.1-8b": _llama3("llama-3.1-8b", 4096, 14336, 32, 32, 8),
    "llama-3.3-70b": _llama3("llama-3.3-70b", 8192, 28672, 80, 64, 8),
    # the per-rank slice of Llama-3.1-8B at 8 ranks as a single-GPU model (tools/bench_config.py: kernel tuning for the N = 8 regime)
    "llama-3.1-8b-slice8": replace(_llama3("llama-3.1-8b-slice8", 4096, 1792, 32, 4, 1, vocab=16032), head_dim=128),
    "llama-3.1-405b": _llama3("llama-3.1-405b", 16384, 53248, 126, 128, 8),
    "qwen3-0.6b": ModelConfig("qwen3-0.6b", ARCH_QWEN3, 1024, 3072, 28, 16, 8, 151936, 40960, head_dim=128,
                              rope_theta=1000000, norm_epsilon=6),
    "qwen3-8b": ModelConfig("qwen3-8b", ARCH_QWEN3, 4096, 12288, 36, 32, 8, 151936, 40960, head_dim=128,
                            rope_theta=1000000, norm_epsilon=6),
    "qwen3-14b": ModelConfig("qwen3-14b", ARCH_QWEN3, 5120, 17408, 40, 40, 8, 151936, 40960, head_dim=128,
                             rope_theta=1000000, norm_epsilon=6),
    "qwen3-30b-a3b": ModelConfig("qwen3-30b-a3b", ARCH_QWEN3_MOE, 2048, 6144, 48, 32, 4, 151936, 40960,
                                 head_dim=128, n_experts=128, n_active_experts=8, moe_hidden_dim=768,
                                 rope_theta=1000000, norm_epsilon=6),
    # test-sized
    "tiny-llama": ModelConfig("tiny-llama", ARCH_LLAMA, 256, 512, 2, 4, 2, 512, 256, rope_theta=10000),
    "tiny-llama31": _llama3("tiny-llama31", 512, 1024, 3, 8, 4, seq=512, vocab=512),
    "tiny-llama-tp8": _llama3("tiny-llama-tp8", 1024, 2048, 2, 16, 8, seq=512, vocab=1024),
    "tiny-llama-kvrep": _llama3("tiny-llama-kvrep", 512, 1024, 2, 8, 2, seq=512, vocab=512),
    "tiny-llama-kvrep8": _llama3("tiny-llama-kvrep8", 1024, 2048, 2, 16, 2, seq=512, vocab=1024),   # 8 ranks: 2 heads each, every KV head on 4 ranks
    "tiny-qwen3": ModelConfig("tiny-qwen3", ARCH_QWEN3, 256, 512, 2, 4, 2, 512, 256, head_dim=128,
                              rope_theta=1000000, norm_epsilon=6),
    "tiny-qwen3-moe": ModelConfig("tiny-qwen3-moe", ARCH_QWEN3_MOE, 256, 512, 2, 4, 2, 512, 256, head_dim=128,
                                  n_experts=8, n_active_experts=2, moe_hidden_dim=256, rope_theta=1000000,
                                  norm_epsilon=6),
}


def get_config(name: str) -> ModelConfig:
    try:
        return PRESETS[name]
    except KeyError:
        raise ValueError(f"unknown model config {name!r}; known: {sorted(PRESETS)}") from None
