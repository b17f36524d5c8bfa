"""Vectorised numpy codecs for the `.m` quant formats (bulk/tooling twin of csrc/host/quants.cpp).

Format semantics: reference src/nn/nn-quants.hpp:64-72, converter/writer.py:29-78.
  Q40: 18-byte blocks  [fp16 d | 16 bytes], byte j = elem j (low nibble) | elem j+16 (high nibble), value=(nib-8)*d
  Q80: 34-byte blocks  [fp16 d | 32 int8],  value = q*d
"""
from __future__ import annotations

import numpy as np

F_32, F_16, F_Q40, F_Q80 = 0, 1, 2, 3
_NAMES = {"f32": F_32, "f16": F_16, "q40": F_Q40, "q80": F_Q80}
QBLOCK = 32
Q40_BYTES = 18
Q80_BYTES = 34


def parse_float_type(name: str) -> int:
    try:
        return _NAMES[name]
    except KeyError:
        raise ValueError(f"{name} is not supported") from None


def float_type_name(t: int) -> str:
    return {v: k for k, v in _NAMES.items()}[t]


def tensor_bytes(t: int, n: int) -> int:
    if t == F_32:
        return n * 4
    if t == F_16:
        return n * 2
    if n % QBLOCK:
        raise ValueError("quantised tensors need a multiple of 32 elements")
    return n // QBLOCK * (Q40_BYTES if t == F_Q40 else Q80_BYTES)


def quantize_q40(x: np.ndarray) -> np.ndarray:
    """f32 array (size % 32 == 0) -> uint8 array [nBlocks, 18]."""
    g = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, QBLOCK)
    gmax = g.max(axis=1)
    gmin = g.min(axis=1)
    extreme = np.where(-gmin > gmax, gmin, gmax)
    d = (extreme / -8.0).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
    q = np.clip((g * inv[:, None] + np.float32(8.5)), 0, 15).astype(np.uint8)  # trunc toward zero == floor for >= 0
    out = np.empty((g.shape[0], Q40_BYTES), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:] = q[:, :16] | (q[:, 16:] << 4)
    return out


def dequantize_q40(raw: np.ndarray, n: int | None = None) -> np.ndarray:
    b = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1, Q40_BYTES)
    d = b[:, 0:2].copy().view(np.float16).astype(np.float32)  # [nb, 1]
    qs = b[:, 2:]
    lo = (qs & 0x0F).astype(np.int8) - 8
    hi = (qs >> 4).astype(np.int8) - 8
    out = np.concatenate([lo, hi], axis=1).astype(np.float32) * d
    out = out.reshape(-1)
    return out if n is None else out[:n]


def quantize_q80(x: np.ndarray) -> np.ndarray:
    g = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, QBLOCK)
    amax = np.abs(g).max(axis=1)
    d = (amax / np.float32(127.0)).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
    # round half away from zero (C roundf), not numpy's half-to-even
    v = g * inv[:, None]
    q = (np.sign(v) * np.floor(np.abs(v) + np.float32(0.5))).astype(np.int8)
    out = np.empty((g.shape[0], Q80_BYTES), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:] = q.view(np.uint8)
    return out


def dequantize_q80(raw: np.ndarray, n: int | None = None) -> np.ndarray:
    b = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1, Q80_BYTES)
    d = b[:, 0:2].copy().view(np.float16).astype(np.float32)
    out = (b[:, 2:].view(np.int8).astype(np.float32) * d).reshape(-1)
    return out if n is None else out[:n]


def quantize(t: int, x: np.ndarray) -> np.ndarray:
    """f32 values -> flat uint8 byte stream of float type `t`."""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    if t == F_32:
        return x.view(np.uint8)
    if t == F_16:
        return x.astype(np.float16).view(np.uint8)
    if t == F_Q40:
        return quantize_q40(x).reshape(-1)
    if t == F_Q80:
        return quantize_q80(x).reshape(-1)
    raise ValueError("Unknown float type")


def dequantize(t: int, raw: np.ndarray, n: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1)
    if t == F_32:
        return raw[: n * 4].view(np.float32).copy()
    if t == F_16:
        return raw[: n * 2].view(np.float16).astype(np.float32)
    if t == F_Q40:
        return dequantize_q40(raw[: tensor_bytes(t, n)], n)
    if t == F_Q80:
        return dequantize_q80(raw[: tensor_bytes(t, n)], n)
    raise ValueError("Unknown float type")
