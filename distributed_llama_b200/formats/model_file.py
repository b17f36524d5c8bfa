"""`.m` model file access.

Layout (reference converter/writer.py:109-145, src/llm.cpp:52-97,624-666):
  int32 magic 0xA00ABCD | int32 headerSize (= 8 + 8*nPairs) | nPairs x (int32 key, int32 value) | tensors
Tensor order: embedding f32; per layer q,k,v,wo, (w1,w2,w3 | moe gate f32 + per expert w1,w2,w3),
[q_norm,k_norm], norm_0, norm_1; final_norm; wcls.  Matrices are row-major [out][in], quant blocks along `in`.

The header parser / tensor directory / slicers live in the native host library (csrc/host/model_format.cpp);
this module adds mmap access and the writers used by converters and the synthetic-model generator.
"""
from __future__ import annotations

import struct
from typing import BinaryIO, Dict, Iterable, List

import numpy as np

from .. import host
from . import quants

HEADER_KEYS: Dict[str, int] = {
    "version": 0, "arch_type": 1, "dim": 2, "hidden_dim": 3, "n_layers": 4, "n_heads": 5, "n_kv_heads": 6,
    "n_experts": 7, "n_active_experts": 8, "vocab_size": 9, "max_seq_len": 10, "hidden_act": 11,
    "rope_theta": 12, "weights_float_type": 13, "rope_scaling_factor": 14, "rope_scaling_low_freq_factor": 15,
    "rope_scaling_high_freq_factory": 16, "rope_scaling_orig_max_seq_len": 17, "rope_type": 18, "head_dim": 19,
    "norm_epsilon": 20, "moe_hidden_dim": 21,
}


def write_model_header(f: BinaryIO, params: Dict[str, int]) -> int:
    """Writes magic + size + (key,value) pairs for every known key in `params` (unknown keys are skipped)."""
    pairs = [(HEADER_KEYS[k], int(v)) for k, v in params.items() if k in HEADER_KEYS]
    data = host().build_model_header(pairs)
    f.write(data)
    return len(data)


def write_tensor(f: BinaryIO, x, float_type: int) -> int:
    """Serialises a tensor (numpy array or torch tensor, any shape) in row-major order."""
    if hasattr(x, "detach"):
        x = x.detach().to("cpu").float().numpy()
    raw = quants.quantize(float_type, np.asarray(x, dtype=np.float32))
    f.write(raw.tobytes())
    return raw.size


class ModelFile:
    """Memory-mapped `.m` file + tensor directory."""

    def __init__(self, path: str, max_seq_len: int = 0):
        self.path = path
        self.header = host().load_model_header(path, max_seq_len)
        self.directory = host().build_tensor_directory(self.header, True)
        self.data = np.memmap(path, dtype=np.uint8, mode="r")
        self._index = {(t.name, t.layer, t.expert): t for t in self.directory}

    def entry(self, name: str, layer: int = 0, expert: int = 0):
        return self._index[(name, layer, expert)]

    def raw(self, entry) -> np.ndarray:
        return self.data[entry.offset: entry.offset + entry.n_bytes]

    def tensor_f32(self, entry) -> np.ndarray:
        """Full tensor dequantised to f32, shape [d, n] (or [n] for vectors)."""
        x = quants.dequantize(entry.type, self.raw(entry), entry.d * entry.n)
        return x.reshape(entry.d, entry.n) if entry.d > 1 else x

    def slice_bytes(self, entry, rank: int, n_ranks: int) -> np.ndarray:
        """This rank's share of the tensor as tightly packed bytes (native extractSlice)."""
        return host().extract_slice(entry, self.data, rank, n_ranks)

    def slice_f32(self, entry, rank: int, n_ranks: int) -> np.ndarray:
        s = host().slice_tensor(entry, rank, n_ranks)
        raw = self.slice_bytes(entry, rank, n_ranks)
        return quants.dequantize(entry.type, raw, s.n_rows * s.n_cols).reshape(s.n_rows, s.n_cols)

    def close(self):
        del self.data
