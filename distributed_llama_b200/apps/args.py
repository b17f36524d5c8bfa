"""Command-line flags of `dllama` / `dllama-api` — same surface as the reference parser (src/app.cpp:24-135):
`--key value` pairs, first positional = mode, `--workers` variadic, `--help/-h/--usage` short-circuit, unknown flags
are an error. Flags that only make sense for the CPU/TCP design (`--nthreads`, `--net-turbo`, `--gpu-index`,
`--gpu-segments`, `--workers host:port`) are accepted for drop-in compatibility; on the B200 build the worker set is
the torch.distributed world (one process per GPU), so `--workers` only has to name as many entries as there are ranks-1.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class AppArgs:
    mode: Optional[str] = None
    help: bool = False
    info: bool = True
    model: Optional[str] = None
    tokenizer: Optional[str] = None
    prompt: Optional[str] = None
    buffer_float_type: str = "q80"
    workers: List[str] = field(default_factory=list)
    host: str = "0.0.0.0"
    port: int = 9990
    nthreads: int = 1
    n_batches: int = 32
    steps: int = 0
    temperature: float = 0.8
    topp: float = 0.9
    seed: int = field(default_factory=lambda: int(time.time()))
    chat_template: Optional[str] = None
    max_seq_len: int = 0
    net_turbo: bool = True
    gpu_index: int = -1
    gpu_segments: Optional[str] = None
    benchmark: bool = False
    # B200 build extras
    gpus: int = 0            # spawn this many local ranks (0 = use the current torch.distributed world / single GPU)
    moe_mode: str = "auto"   # tp | ep | auto


FLOAT_TYPES = ("f32", "f16", "q40", "q80")
TEMPLATES = ("llama2", "llama3", "deepSeek3", "chatml")


def parse_args(argv: List[str], require_mode: bool) -> AppArgs:
    a = AppArgs()
    i = 0
    if require_mode and len(argv) > 0 and not argv[0].startswith("-"):
        a.mode = argv[0]
        i = 1
    if any(x in ("--usage", "--help", "-h") for x in argv):
        a.help = True
        return a
    while i < len(argv):
        name = argv[i]
        if i + 1 >= len(argv):
            raise ValueError(f"Missing value for option: {name}")
        value = argv[i + 1]
        if name == "--model":
            a.model = value
        elif name == "--tokenizer":
            a.tokenizer = value
        elif name == "--prompt":
            a.prompt = value
        elif name == "--buffer-float-type":
            if value not in FLOAT_TYPES:
                raise ValueError(f"Invalid float type: {value}")
            a.buffer_float_type = value
        elif name == "--workers":
            j = i + 1
            while j < len(argv) and not argv[j].startswith("-"):
                if ":" not in argv[j]:
                    raise ValueError(f"Invalid worker address: {argv[j]}")
                a.workers.append(argv[j])
                j += 1
            i = j
            continue
        elif name == "--port":
            a.port = int(value)
        elif name == "--host":
            a.host = value
        elif name == "--nthreads":
            a.nthreads = int(value)
            if a.nthreads < 1:
                raise ValueError("Number of threads must be at least 1")
        elif name == "--steps":
            a.steps = int(value)
        elif name == "--temperature":
            a.temperature = float(value)
        elif name == "--topp":
            a.topp = float(value)
        elif name == "--seed":
            a.seed = int(value)
        elif name == "--chat-template":
            if value not in TEMPLATES:
                raise ValueError(f"Invalid chat template type: {value}")
            a.chat_template = value
        elif name == "--max-seq-len":
            a.max_seq_len = int(value)
        elif name == "--gpu-index":
            a.gpu_index = int(value)
        elif name == "--gpu-segments":
            if ":" not in value:
                raise ValueError("GPU segments expected in the format <from>:<to>")
            a.gpu_segments = value
        elif name == "--net-turbo":
            a.net_turbo = int(value) == 1
        elif name == "--gpus":
            a.gpus = int(value)
        elif name == "--moe-mode":
            a.moe_mode = value
        else:
            raise ValueError(f"Unknown option: {name}")
        i += 2
    return a


USAGE = """Usage: dllama {inference|chat|perplexity|worker} {--model <path>} {--tokenizer <path>}
        [--prompt <text>] [--steps <n>] [--buffer-float-type {f32|f16|q40|q80}] [--max-seq-len <max>]
        [--temperature <temp>] [--topp <t>] [--seed <s>] [--chat-template {llama2|llama3|deepSeek3|chatml}]
        [--gpus <n>]                      run tensor-parallel on n local B200s (spawns one process per GPU)
        [--workers <ip:port> ...] [--nthreads <n>] [--net-turbo {0|1}] [--gpu-index <i>] [--gpu-segments <a:b>]
                                          accepted for compatibility with the reference CLI
"""
