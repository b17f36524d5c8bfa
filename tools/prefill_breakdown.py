"""Per-kernel time of one 64-token prefill (Llama-3.1-8B shapes, random-init weights). Run under
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prefill_launches.csv python tools/prefill_breakdown.py
and summarise with `python tools/prefill_breakdown.py --summarise gpurun_out/prefill_launches.csv` (durations under ncu are
serialised per kernel: shares, not bench values)."""
import csv, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
    rows = [r for r in csv.reader(l for l in open(sys.argv[2]) if l.startswith('"'))]
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    names = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:] if len(r) > vi]
    # the last prefill = everything after the last embedding kernel
    last = max(i for i, (n, _) in enumerate(names) if "mbedding" in n)
    agg = collections.OrderedDict()
    for n, v in names[last:]:
        k = n.split("(")[0]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"last prefill: {sum(a[0] for a in agg.values())} kernels, {tot / 1e3:.1f} us of kernel time")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v / 1e3:9.1f} us {100 * v / tot:5.1f}%  x{c:<4d} {v / c / 1e3:7.2f} us each  {k[:90]}")
    sys.exit(0)

import torch
from distributed_llama_b200.models.config import get_config
from distributed_llama_b200.models.loader import synthetic_device_weights
from distributed_llama_b200.runtime import Engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = Engine(synthetic_device_weights(get_config("llama-3.1-8b"), 0, 1, "cuda:0", max_seq_len=2048))
prompt = [(7 * i + 3) % 1000 + 1 for i in range(T)]
for _ in range(2):
    eng.prefill(prompt, 0)
    torch.cuda.synchronize()
