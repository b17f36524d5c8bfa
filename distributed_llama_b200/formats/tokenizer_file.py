"""`.t` tokenizer file (reference converter/tokenizer-writer.py:3-57, src/tokenizer.cpp:57-136).

  int32 magic 0x567124 | int32 headerSize | kv pairs {bos_id, version=1, vocab_size, max_token_length,
  [chat_template len], n_eos_tokens, add_bos} | chat template bytes | int32 eos ids | vocab x (f32 score, u32 len, bytes)
"""
from __future__ import annotations

from typing import List, Optional, Sequence

from .. import host


def write_tokenizer(path: str, tokens: Sequence[bytes], scores: Sequence[float], chat_template: Optional[bytes],
                    bos_id: int, add_bos: bool, eos_tokens: Sequence[int]) -> None:
    h = host()
    d = h.TokenizerData()
    d.vocab = [bytes(t) for t in tokens]
    d.scores = [float(s) for s in scores]
    d.bos_id = int(bos_id)
    d.add_bos = bool(add_bos)
    d.eos_ids = [int(e) for e in eos_tokens]
    d.chat_template = bytes(chat_template) if chat_template else b""
    h.write_tokenizer_file(path, d)


def read_tokenizer(path: str):
    return host().read_tokenizer_file(path)
