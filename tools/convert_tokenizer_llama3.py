#!/usr/bin/env python
"""Llama-3 tiktoken `tokenizer.model` (base64 token + rank per line) -> `.t` (reference converter/convert-tokenizer-llama3.py)."""
import base64
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from distributed_llama_b200.formats import write_tokenizer

N_RESERVED = 256
SPECIALS = ["<|begin_of_text|>", "<|end_of_text|>", "<|reserved_special_token_0|>", "<|reserved_special_token_1|>",
            "<|reserved_special_token_2|>", "<|reserved_special_token_3|>", "<|start_header_id|>", "<|end_header_id|>",
            "<|reserved_special_token_4|>", "<|eot_id|>"]
SPECIALS += [f"<|reserved_special_token_{i}|>" for i in range(5, N_RESERVED - 5)]
TEMPLATE = ("{% set loop_messages = messages %}{% for message in loop_messages %}{% set content = '<|start_header_id|>' + message['role'] + "
            "'<|end_header_id|>\n\n'+ message['content'] | trim + '<|eot_id|>' %}{% if loop.index0 == 0 %}{% set content = bos_token + content %}"
            "{% endif %}{{ content }}{% endfor %}{% if add_generation_prompt %}{{ '<|start_header_id|>assistant<|end_header_id|>\n\n' }}{% endif %}")


def convert(model_path: str, out_path: str) -> str:
    tokens, scores = [], []
    with open(model_path, "r") as f:
        for line in f:
            if not line.strip():
                continue
            tok, rank = line.split()
            tokens.append(base64.b64decode(tok))
            scores.append(-float(rank))
    bos = len(tokens)
    for i, s in enumerate(SPECIALS):
        tokens.append(s.encode("utf-8"))
        scores.append(-float(bos + i))
    write_tokenizer(out_path, tokens, scores, TEMPLATE.encode("utf-8"), bos, True, [bos + 1, bos + 9])
    return out_path


if __name__ == "__main__":
    if len(sys.argv) < 2:
        print("Usage: python convert_tokenizer_llama3.py <tokenizerPath>")
        sys.exit(1)
    print("✅ Created " + convert(sys.argv[1], "dllama_tokenizer_llama3.t"))
