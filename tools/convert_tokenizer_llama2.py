#!/usr/bin/env python
"""Llama-2 `tokenizer.model` (sentencepiece) -> `.t` (reference converter/convert-tokenizer-llama2.py, with the 7-argument
writer signature the reference script forgot to update)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from distributed_llama_b200.formats import write_tokenizer

LLAMA2_TEMPLATE = b"{% if messages[0]['role'] == 'system' %}[INST] <<SYS>>\n{{ messages[0]['content'] }}\n<</SYS>>\n\n{% endif %}"


def convert(folder: str, out_path: str) -> str:
    from sentencepiece import SentencePieceProcessor
    sp = SentencePieceProcessor(model_file=os.path.join(folder, "tokenizer.model"))
    tokens, scores = [], []
    for i in range(sp.vocab_size()):
        t = sp.id_to_piece(i)
        if i == sp.bos_id():
            t = "\n<s>\n"
        elif i == sp.eos_id():
            t = "\n</s>\n"
        t = t.replace("▁", " ")
        tokens.append(bytes.fromhex(t[3:-1]) if len(t) == 6 and t.startswith("<0x") and t.endswith(">") else t.encode("utf-8"))
        scores.append(sp.get_score(i))
    write_tokenizer(out_path, tokens, scores, LLAMA2_TEMPLATE, sp.bos_id(), True, [sp.eos_id()])
    return out_path


if __name__ == "__main__":
    if len(sys.argv) < 2:
        print("Usage: python convert_tokenizer_llama2.py <llama2FolderPath>")
        sys.exit(1)
    print("✅ Created " + convert(sys.argv[1], "dllama_tokenizer_llama2.t"))
