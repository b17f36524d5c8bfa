#!/usr/bin/env python
"""HF tokenizer folder -> `.t` (reference converter/convert-tokenizer-hf.py:14-137).

Fast (byte-level BPE) tokenizers: the GPT-2 printable-unicode alphabet is mapped back to raw bytes and scores are -id, so
the merge loop of the runtime prefers earlier (more frequent) merges. SentencePiece tokenizers keep their own scores,
`▁` becomes a space and `<0xNN>` pieces become single bytes.
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from distributed_llama_b200.formats import write_tokenizer


def gpt2_unicode_to_byte():
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    table = {chr(b): b for b in keep}
    extra = 0
    for b in range(256):
        if b not in keep:
            table[chr(256 + extra)] = b
            extra += 1
    return table


def from_fast_tokenizer(folder: str):
    from transformers import PreTrainedTokenizerFast
    tk = PreTrainedTokenizerFast(tokenizer_file=os.path.join(folder, "tokenizer.json"))
    u2b = gpt2_unicode_to_byte()
    n = len(tk.get_vocab())
    tokens, scores = [], []
    for i in range(n):
        piece = tk.convert_ids_to_tokens([i])[0]
        raw = bytearray()
        for ch in piece:
            raw += bytes([u2b[ch]]) if ch in u2b else ch.encode("utf-8")
        tokens.append(bytes(raw))
        scores.append(-float(i))
    bos, eos = tk.bos_token_id, ([tk.eos_token_id] if tk.eos_token_id is not None else None)
    if bos is None or eos is None:
        with open(os.path.join(folder, "config.json")) as f:
            cfg = json.load(f)
        bos = cfg["bos_token_id"] if bos is None else bos
        if eos is None:
            e = cfg["eos_token_id"]
            eos = e if isinstance(e, list) else [e]
    return tokens, scores, bos, eos


def from_sentencepiece(folder: str):
    from sentencepiece import SentencePieceProcessor
    sp = SentencePieceProcessor(model_file=os.path.join(folder, "tokenizer.model"))
    tokens, scores = [], []
    for i in range(sp.vocab_size()):
        t = sp.id_to_piece(i).replace("▁", " ")
        tokens.append(bytes.fromhex(t[3:-1]) if len(t) == 6 and t.startswith("<0x") and t.endswith(">") else t.encode("utf-8"))
        scores.append(sp.get_score(i))
    return tokens, scores, sp.bos_id(), [sp.eos_id()]


def convert(folder: str, out_path: str) -> str:
    with open(os.path.join(folder, "tokenizer_config.json"), encoding="utf-8") as f:
        cfg = json.load(f)
    cls = cfg.get("tokenizer_class")
    if cls in ("PreTrainedTokenizerFast", "LlamaTokenizerFast", "Qwen2Tokenizer", "TokenizersBackend"):
        tokens, scores, bos, eos = from_fast_tokenizer(folder)
    elif cls == "LlamaTokenizer":
        tokens, scores, bos, eos = from_sentencepiece(folder)
    else:
        raise ValueError(f"Tokenizer {cls} is not supported")
    if bos is None or eos is None:
        raise ValueError("Cannot resolve bosId or eosIds")
    template = cfg["chat_template"].encode("utf-8") if isinstance(cfg.get("chat_template"), str) else None
    write_tokenizer(out_path, tokens, scores, template, bos, bool(cfg.get("add_bos_token", True)), eos)
    return out_path


if __name__ == "__main__":
    if len(sys.argv) < 3:
        print("Usage: python convert_tokenizer_hf.py <tokenizerFolderPath> <name>")
        sys.exit(1)
    out = f"dllama_tokenizer_{sys.argv[2]}.t"
    convert(sys.argv[1], out)
    print(f"✅ Created {out}")
